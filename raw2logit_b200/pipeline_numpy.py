"""Kernel-backed stand-in for the reference's static numpy pipeline (``processing/pipeline_numpy.py``).

``--processing_mode static`` runs ``RawProcessingPipeline`` per image inside 16 DataLoader workers
(pipeline_numpy.py:36-68, train.py:316-320).  Here the same chain -- black level, bilinear demosaic, white balance,
colour matrix, sharpening filter, Gaussian denoising, clip, gamma, with the numpy chain's own boundary rules (scipy
half-sample reflection, zero fill, clip at 0) -- is one fused CUDA kernel over a whole batch (``csrc/isp_numpy.cu``).
Same class name, constructor arguments and call signature; CUDA only, forward only (the numpy chain has no gradient).

Served options: ``debayer='bilinear'``; ``sharpening='sharpening_filter'`` (any other string that the reference's
``processing`` would silently skip is skipped here too, except ``'unsharp_masking'``, which needs skimage's
unsharp_mask and raises); ``denoising='gaussian_denoising'`` (other names the reference would skip are skipped; the
median / FFT / TV / bilateral filters raise).
"""
import numpy as np
import torch

from . import ops  # noqa: F401  (loads the operator library)

_SKIP_OK_SHARP = ("sharpening_filter",)
_UNSERVED_SHARP = ("unsharp_masking",)
_UNSERVED_DENOISE = ("median_denoising", "fft_denoising", "tv_chambolle", "tv_bregman", "bilateral")


def processing(raw, black_level, white_balance, colour_matrix, debayer="bilinear", sharpening="unsharp_masking",
               denoising="median_filter", gaussian_sigma=0.5, gamma=2.2, bits=16):
    """``processing`` of pipeline_numpy.py:70-141 on a CUDA batch: raw (B, H, W) or (H, W), float or uint16
    (value = u / (2**bits - 1)) -> (B, 3, H, W) / (3, H, W) float32.  Defaults are the reference's."""
    if debayer != "bilinear":
        raise NotImplementedError(f"debayer={debayer!r}: only the bilinear demosaic is a kernel here")
    if sharpening in _UNSERVED_SHARP:
        raise NotImplementedError(f"sharpening={sharpening!r} needs skimage.filters.unsharp_mask; use 'sharpening_filter'")
    if denoising in _UNSERVED_DENOISE:
        raise NotImplementedError(f"denoising={denoising!r} is not a kernel here; use 'gaussian_denoising'")
    if not (isinstance(raw, torch.Tensor) and raw.is_cuda):
        raise RuntimeError("raw2logit_b200.pipeline_numpy is CUDA-only (no CPU fallback)")
    single = raw.ndim == 2
    x = raw[None] if single else raw
    out = torch.ops.raw2logit_isp.numpy_forward(
        x, [float(v) for v in black_level], [float(v) for v in white_balance],
        [float(v) for v in np.asarray(colour_matrix, dtype=np.float64).reshape(-1)], sharpening == "sharpening_filter",
        denoising == "gaussian_denoising", float(gaussian_sigma), float(gamma), float(2 ** bits - 1))
    return out[0] if single else out


class RawProcessingPipeline(object):
    """pipeline_numpy.py:36-68 (same constructor defaults, including the reference's ``denoising='gaussian'``, a name
    its ``processing`` matches to no filter).  ``__call__(img)``: (H, W) -> (3, H, W) tensor like the reference; a
    (B, H, W) batch is processed in one launch."""

    def __init__(self, camera_parameters, debayer='bilinear', sharpening='unsharp_masking', denoising='gaussian'):
        self.camera_parameters = camera_parameters
        self.debayer = debayer
        self.sharpening = sharpening
        self.denoising = denoising

    def __call__(self, img):
        black_level, white_balance, colour_matrix = self.camera_parameters
        if isinstance(img, np.ndarray):
            img = torch.from_numpy(np.ascontiguousarray(img)).cuda()
        return processing(img, black_level, white_balance, colour_matrix, debayer=self.debayer,
                          sharpening=self.sharpening, denoising=self.denoising)
