"""Staged-mode kernels (csrc/isp_stages.cu) against stock torch ops on the same GPU inputs (full fp32, TF32 off), and the
staged forward as a whole: no stock convolution runs inside it.  The reference's stage tensors / stage gradients are
checked in tests/test_gpu_parity.py::test_track_stages_mode_matches_reference_stages_and_stage_gradients."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _full_fp32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def _ref_conv(x, w, reflect):
    k = w.shape[-1]
    return F.conv2d(F.pad(x, (k // 2,) * 4, mode='reflect' if reflect else 'constant'), w)


@pytest.mark.parametrize("shape", [(2, 3, 40, 52), (1, 3, 3, 3), (3, 3, 7, 5), (2, 3, 64, 64)])
@pytest.mark.parametrize("k,reflect", [(3, True), (3, False), (5, True), (5, False)])
def test_stage_conv_forward_and_both_gradients_match_torch(shape, k, reflect):
    from raw2logit_b200.staged import _StageConv
    if reflect and min(shape[2], shape[3]) <= k // 2:
        pytest.skip("reflect padding needs a frame larger than the pad")
    g = torch.Generator().manual_seed(7)
    x = torch.randn(shape, generator=g).cuda().requires_grad_()
    w = (0.3 * torch.randn(3, 3, k, k, generator=g)).cuda().requires_grad_()
    cot = torch.randn(shape, generator=g).cuda()
    y = _StageConv.apply(x, w, reflect)
    y.backward(cot)
    x2, w2 = x.detach().clone().requires_grad_(), w.detach().clone().requires_grad_()
    y2 = _ref_conv(x2, w2, reflect)
    y2.backward(cot)
    assert torch.allclose(y, y2, atol=1e-5, rtol=1e-5)
    assert torch.allclose(x.grad, x2.grad, atol=1e-5, rtol=1e-5)
    assert torch.allclose(w.grad, w2.grad, atol=1e-4 * max(1.0, float(w2.grad.abs().max())), rtol=1e-5)


def test_stage_conv_weight_gradient_is_bit_reproducible():
    from raw2logit_b200.staged import _StageConv
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 3, 96, 128, generator=g).cuda()
    cot = torch.randn(4, 3, 96, 128, generator=g).cuda()
    grads = []
    for _ in range(3):
        w = torch.randn(3, 3, 5, 5, generator=torch.Generator().manual_seed(5)).cuda().requires_grad_()
        _StageConv.apply(x, w, True).backward(cot)
        grads.append(w.grad.cpu().numpy().copy())
    assert np.array_equal(grads[0], grads[1]) and np.array_equal(grads[0], grads[2])


def test_stage_clip_and_gamma_match_torch():
    from raw2logit_b200.staged import _StageClip, _StageGamma
    g = torch.Generator().manual_seed(11)
    x = (1.4 * torch.rand(3, 3, 33, 47, generator=g) - 0.2).cuda().requires_grad_()
    x.data[0, 0, 0, :4] = torch.tensor([1e-5, 1.0, 0.0, 2.0])            # the inclusive ends and both clipped sides
    gamma = torch.tensor([2.2]).cuda().requires_grad_()
    cot = torch.randn(3, 3, 33, 47, generator=g).cuda()
    c = _StageClip.apply(x, 1e-5, 1.0)
    y = _StageGamma.apply(c, gamma)
    y.backward(cot)
    x2, g2 = x.detach().clone().requires_grad_(), gamma.detach().clone().requires_grad_()
    c2 = torch.clip(x2, 1e-5, 1)
    y2 = torch.exp((1 / g2) * torch.log(c2))
    y2.backward(cot)
    assert torch.equal(c, c2)
    assert torch.allclose(y, y2, atol=1e-6, rtol=1e-6)
    assert torch.allclose(x.grad, x2.grad, atol=1e-4, rtol=1e-4)
    assert abs(float(gamma.grad) - float(g2.grad)) <= 1e-4 * max(1.0, abs(float(g2.grad)))


def test_staged_forward_runs_no_stock_convolution(monkeypatch):
    from processing.pipeline_torch import ParametrizedProcessing
    from raw2logit_b200 import synthetic as syn

    def boom(*a, **k):
        raise AssertionError("stock convolution called inside the staged path")

    mod = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], track_stages=True, batch_norm_output=False).cuda()
    raw = syn.smooth_scene(2, 64, 96, "drone", seed=2).cuda().requires_grad_()
    monkeypatch.setattr(F, "conv2d", boom)
    monkeypatch.setattr(torch, "conv2d", boom)
    out = mod(raw)
    out.mean().backward()
    assert list(mod.stages) == ["demosaic", "color_correct", "sharpening", "gaussian", "clipped", "gamma_correct"]
    assert all(t.grad is not None for t in mod.stages.values()) and raw.grad is not None
    fused = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], batch_norm_output=False).cuda()
    assert torch.allclose(out, fused(raw.detach()), atol=2e-5)           # staged = fused up to the YUV round trip (1.3e-5)
