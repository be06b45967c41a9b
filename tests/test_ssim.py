"""SSIM regulariser (reference utils/ssim.py:19-39): the oracle restatement against golden vectors
written from the unmodified reference file (CPU), the fused CUDA kernels against the same vectors (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import ssim_oracle
from tests.conftest import ROOT

GOLDEN = np.load(os.path.join(ROOT, "tests", "golden", "ssim.npz"))
CASES = sorted({k.split(".")[0] for k in GOLDEN.files})


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    for tag, dt, tol in (("f32", torch.float32, 2e-6), ("f64", torch.float64, 1e-12)):
        a = torch.from_numpy(GOLDEN[f"{name}.img1"]).to(dt).requires_grad_(True)
        b = torch.from_numpy(GOLDEN[f"{name}.img2"]).to(dt).requires_grad_(True)
        m = ssim_oracle.ssim(a, b)
        m.backward()
        assert abs(m.item() - float(GOLDEN[f"{name}.{tag}.mean"])) <= tol
        assert np.abs(a.grad.numpy() - GOLDEN[f"{name}.{tag}.grad1"]).max() <= tol
        assert np.abs(b.grad.numpy() - GOLDEN[f"{name}.{tag}.grad2"]).max() <= tol
        per = ssim_oracle.ssim(a.detach(), b.detach(), size_average=False)
        assert np.abs(per.numpy() - GOLDEN[f"{name}.{tag}.per_image"]).max() <= tol


def test_module_surface_matches_reference():
    """Same names and defaults as utils/ssim.py (no GPU needed: construction and the argument rule only)."""
    from utils.ssim import SSIM, ssim  # noqa: F401  (module path of the reference)
    m = SSIM()
    assert (m.window_size, m.size_average) == (11, True)
    with pytest.raises(NotImplementedError):
        ssim(torch.zeros(1, 3, 16, 16), torch.zeros(1, 3, 16, 16), window_size=7)
    with pytest.raises(RuntimeError):
        ssim(torch.zeros(1, 3, 16, 16), torch.zeros(1, 3, 16, 16))            # CUDA only, no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_fused_ssim_matches_reference_golden(name):
    """Value within 1e-5 of the fp32 reference, gradients within 1e-5 relative to their scale; the fp64 truth is no
    further away than twice the reference's own fp32 error."""
    from raw2logit_b200.ssim import SSIM, ssim
    dev = torch.device("cuda:0")
    a = torch.from_numpy(GOLDEN[f"{name}.img1"]).to(dev).requires_grad_(True)
    b = torch.from_numpy(GOLDEN[f"{name}.img2"]).to(dev).requires_grad_(True)
    m = SSIM(window_size=11)(a, b)
    (3.0 * m).backward()
    assert m.shape == () and m.dtype == torch.float32
    assert abs(m.item() - float(GOLDEN[f"{name}.f32.mean"])) <= 1e-5
    for got, key in ((a.grad, "grad1"), (b.grad, "grad2")):
        want32, want64 = GOLDEN[f"{name}.f32.{key}"], GOLDEN[f"{name}.f64.{key}"]
        scale = max(float(np.abs(want64).max()), 1e-12)
        err = np.abs(got.cpu().numpy() / 3.0 - want64).max()
        ref_err = np.abs(want32 - want64).max()
        assert err <= max(1e-5 * scale, 2.0 * ref_err, 5e-8), (key, err, ref_err, scale)   # (5e-8: identical images, true gradient 0)
    per = ssim(a.detach(), b.detach(), size_average=False)
    assert np.abs(per.cpu().numpy() - GOLDEN[f"{name}.f32.per_image"]).max() <= 1e-5
    # one-sided gradient (the training case: the reference image comes from a no_grad forward, utils/base.py:356)
    b2 = b.detach().clone().requires_grad_(True)
    ssim(a.detach(), b2).backward()
    assert a.grad is not None and torch.allclose(b2.grad * 3.0, b.grad, rtol=1e-6, atol=1e-12)
    # per-image means with per-image upstream gradients
    b3 = b.detach().clone().requires_grad_(True)
    wts = torch.arange(1, b3.shape[0] + 1, device=dev, dtype=torch.float32)
    (ssim(a.detach(), b3, size_average=False) * wts).sum().backward()
    bo = torch.from_numpy(GOLDEN[f"{name}.img2"]).double().requires_grad_(True)            # fp64 truth
    (ssim_oracle.ssim(torch.from_numpy(GOLDEN[f"{name}.img1"]).double(), bo, size_average=False) * wts.cpu().double()).sum().backward()
    sc = max(float(bo.grad.abs().max()), 1e-12)
    assert (b3.grad.cpu().double() - bo.grad).abs().max().item() <= max(5e-5 * sc, 5e-8)
