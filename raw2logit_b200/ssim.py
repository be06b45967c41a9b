"""Kernel-backed drop-in for the reference's ``utils/ssim.py`` (the SSIM regulariser of adversarial training,
train.py:261-262 through ``AuxLoss``, utils/base.py:346-358).

Same names and call signatures: ``SSIM(window_size=11, size_average=True)(img1, img2)`` and
``ssim(img1, img2, window_size=11, size_average=True)``; differentiable in both images.  One fused kernel each way
(``csrc/isp_ssim.cu``) instead of five grouped 11x11 convolutions, ~15 elementwise kernels and their saved
intermediates.  CUDA only; ``window_size`` must be 11 (the reference's only use).  ``processing``-style re-export:
``utils.ssim`` at the repo root.
"""
import torch

from . import ops  # noqa: F401  (loads the operator library)


def _check(img1, img2, window_size):
    if window_size != 11:
        raise NotImplementedError("the fused SSIM kernel is built for window_size=11 (sigma 1.5), the reference's use")
    if not (isinstance(img1, torch.Tensor) and img1.is_cuda and img2.is_cuda):
        raise RuntimeError("raw2logit_b200.ssim is CUDA-only (no CPU fallback)")


def ssim(img1, img2, window_size=11, size_average=True):
    """``utils/ssim.py:66-74``."""
    _check(img1, img2, window_size)
    return torch.ops.raw2logit_isp.ssim(img1, img2, int(window_size), bool(size_average))


class SSIM(torch.nn.Module):
    """``utils/ssim.py:41-64``: the window is a constant of the kernel, so nothing is cached per channel count."""

    def __init__(self, window_size=11, size_average=True):
        super().__init__()
        self.window_size = window_size
        self.size_average = size_average

    def forward(self, img1, img2):
        return ssim(img1, img2, self.window_size, self.size_average)
