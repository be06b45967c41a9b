"""Host logic of the staged mode (raw2logit_b200/staged.py): the combined stage weights reproduce the reference's
per-stage op chain (pipeline_torch.py:187-203) -- checked on CPU in float64 with stock torch ops; the kernels themselves
are checked on the GPU (tests/test_gpu_staged.py, tests/test_gpu_parity.py)."""
import torch
import torch.nn.functional as F

from raw2logit_b200 import staged


def _rand(*shape):
    return torch.rand(*shape, dtype=torch.float64)


def test_colour_stage_weight_is_debayer_then_white_balance_then_colour_matrix():
    torch.manual_seed(0)
    x, wd, wb, ccm = _rand(2, 3, 9, 11), _rand(3, 3, 3, 3), _rand(1, 3), _rand(3, 3)
    ref = F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), wd)                       # :187
    ref = torch.einsum('bchw,kc->bchw', ref, wb)                                        # :190
    ref = torch.einsum('bchw,kc->bkhw', ref, ccm)                                       # :191
    got = F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), staged.colour_stage_weight(wd, wb, ccm))
    assert (got - ref).abs().max() < 1e-12


def test_luma_stage_weight_is_rgb2yuv_filter_on_luma_yuv2rgb():
    torch.manual_seed(1)
    x, m1 = _rand(2, 3, 10, 13), _rand(3, 3)
    m2 = torch.linalg.inv(m1)
    for k, pad_mode in ((3, 'constant'), (5, 'reflect')):
        taps = _rand(1, 1, k, k)
        yuv = torch.einsum('bchw,kc->bkhw', x, m1).contiguous()                         # :194 / :199
        yuv[:, [0]] = F.conv2d(F.pad(yuv[:, [0]], (k // 2,) * 4, mode=pad_mode), taps)    # :195 / :202
        ref = torch.einsum('bchw,kc->bkhw', yuv, m2)                                    # :198 / :203
        got = F.conv2d(F.pad(x, (k // 2,) * 4, mode=pad_mode), staged.luma_stage_weight(taps, m1, m2))
        assert (got - ref).abs().max() < 1e-11


def test_stage_weights_are_differentiable_in_every_parameter():
    wd, wb, ccm = (_rand(3, 3, 3, 3).requires_grad_(), _rand(1, 3).requires_grad_(), _rand(3, 3).requires_grad_())
    staged.colour_stage_weight(wd, wb, ccm).sum().backward()
    assert wd.grad is not None and wb.grad is not None and ccm.grad is not None
    taps = _rand(1, 1, 5, 5).requires_grad_()
    staged.luma_stage_weight(taps, _rand(3, 3), _rand(3, 3)).sum().backward()
    assert taps.grad is not None and taps.grad.shape == taps.shape


def test_stage_operators_are_registered_for_cuda_only():
    import pytest
    ops = torch.ops.raw2logit_isp
    for name in ("stage_conv", "stage_conv_backward", "stage_clip", "stage_clip_backward", "stage_gamma",
                 "stage_gamma_backward"):
        assert hasattr(ops, name)
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.stage_clip(torch.zeros(4), 0.0, 1.0)                 # CPU tensor: no kernel, no fallback
