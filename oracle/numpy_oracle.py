"""CPU restatement of the reference's numpy research pipeline -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/`` and ``bench.py``'s ``cpu_baseline_numpy`` leg import this module.

What it restates: ``processing(img, black_level, white_balance, colour_matrix, debayer='bilinear',
sharpening='sharpening_filter', denoising='gaussian_denoising')`` of ``/root/reference/processing/pipeline_numpy.py``
(``:70-141``; the train.py defaults ``--sp_debayer bilinear --sp_sharpening sharpening_filter --sp_denoising
gaussian_denoising``, ``train.py:95-100``), i.e. the CPU chain the reference's 16 DataLoader workers run per image in
``--processing_mode static`` (``dataset.py`` transform, ``train.py:316-320``) and the chain its own cross-check
compares with the torch pipeline (``pipeline_torch.py:318-324``).

PARITY UNPINNED: three of its functions live in third-party packages that are absent here and un-vendored in the
reference: ``colour_demosaicing.demosaicing_CFA_Bayer_bilinear`` (colour-demosaicing 0.1.6, ``environment.yml:296``;
call site ``pipeline_numpy.py:93``) and ``skimage.color.rgb2yuv / yuv2rgb`` (scikit-image 0.18.1,
``environment.yml:343``; call sites ``:184,189,203,207``).  Their published algorithms are restated below.  The
reference holds no test or golden vector at this boundary; what pins this file is (a) the scipy functions being the
real ones (scipy is installed) and (b) agreement with the torch chain in the image interior
(``tests/test_numpy_oracle.py``: <= 2e-6 away from the border, where the two chains differ by design --
half-sample vs whole-sample reflection, clip at 0 vs 1e-5, ``pipeline_torch.py:233`` notes the mismatch).
"""
import numpy as np
from scipy import ndimage
from scipy.signal import convolve2d

# skimage.color.colorconv: yuv_from_rgb (the same matrix pipeline_torch.py:21-23 prints) and its inverse
_YUV_FROM_RGB = np.array([[0.299, 0.587, 0.114],
                          [-0.14714119, -0.28886916, 0.43601035],
                          [0.61497538, -0.51496512, -0.10001026]])
_RGB_FROM_YUV = np.linalg.inv(_YUV_FROM_RGB)
# colour_demosaicing.bayer.demosaicing.bilinear: H_G, H_RB
_H_G = np.array([[0, 1, 0], [1, 4, 1], [0, 1, 0]], dtype=np.float64) / 4
_H_RB = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=np.float64) / 4
_K_SHARP = np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]])          # pipeline_numpy.py:178


def remove_blacklv(raw, black_level):
    """pipeline_numpy.py:152-158 (in place on its argument, like the reference)."""
    raw[0::2, 0::2] -= black_level[0]
    raw[0::2, 1::2] -= black_level[1]
    raw[1::2, 0::2] -= black_level[2]
    raw[1::2, 1::2] -= black_level[3]
    return raw


def demosaicing_cfa_bayer_bilinear(cfa):
    """colour-demosaicing 0.1.6 ``demosaicing_CFA_Bayer_bilinear(CFA, pattern='RGGB')``: per-channel CFA masks, each
    masked plane convolved with H_RB / H_G by ``scipy.ndimage.convolve`` (default ``mode='reflect'``: half-sample
    symmetric).  Call site pipeline_numpy.py:93."""
    cfa = np.asarray(cfa, dtype=np.float64)
    h, w = cfa.shape
    r_m = np.zeros((h, w)); r_m[0::2, 0::2] = 1
    b_m = np.zeros((h, w)); b_m[1::2, 1::2] = 1
    g_m = np.zeros((h, w)); g_m[0::2, 1::2] = 1; g_m[1::2, 0::2] = 1
    r = ndimage.convolve(cfa * r_m, _H_RB)
    g = ndimage.convolve(cfa * g_m, _H_G)
    b = ndimage.convolve(cfa * b_m, _H_RB)
    return np.stack([r, g, b], axis=-1)


def rgb2yuv(img):
    """skimage 0.18 ``rgb2yuv``: ``arr @ yuv_from_rgb.T`` (call sites pipeline_numpy.py:184,203)."""
    return img @ _YUV_FROM_RGB.T


def yuv2rgb(img):
    """skimage 0.18 ``yuv2rgb``: ``arr @ inv(yuv_from_rgb).T`` (call sites :189,207)."""
    return img @ _RGB_FROM_YUV.T


def processing(img, black_level, white_balance, colour_matrix, gamma=2.2, gaussian_sigma=0.5):
    """pipeline_numpy.py:70-141 with debayer='bilinear', sharpening='sharpening_filter',
    denoising='gaussian_denoising'.  ``img`` (H, W) float; returns (H, W, 3) float64."""
    img = remove_blacklv(np.array(img, dtype=np.float64), black_level)                  # :90
    img = demosaicing_cfa_bayer_bilinear(img)                                           # :93
    img = img * np.asarray(white_balance, dtype=np.float64)                             # :105, :161-162
    img = np.einsum('ijk,lk->ijl', img, np.array(colour_matrix, dtype=np.float64).reshape(3, 3))   # :108, :165-167
    yuv = rgb2yuv(img)                                                                  # :111, :180-191
    yuv[:, :, 0] = convolve2d(yuv[:, :, 0], _K_SHARP, 'same', boundary='fill', fillvalue=0)
    img = yuv2rgb(yuv)
    yuv = rgb2yuv(img)                                                                  # :119, :203-209
    yuv[:, :, 0] = ndimage.gaussian_filter(yuv[:, :, 0], gaussian_sigma)
    img = yuv2rgb(yuv)
    img = np.clip(img, 0, 1)                                                            # :138
    return img ** (1.0 / gamma)                                                         # :139, :241-244


def process_chw(raw, camera_parameters):
    """``RawProcessingPipeline.__call__`` (pipeline_numpy.py:56-68): (H, W) -> float32 (3, H, W)."""
    bl, wb, ccm = camera_parameters
    return processing(raw, bl, wb, ccm).transpose(2, 0, 1).astype(np.float32)


def _worker(args):
    raw, cam = args
    return process_chw(raw, cam)


def process_batch(raws, camera_parameters, pool=None):
    """A batch through the chain, one image per task; ``pool`` = a ``multiprocessing.Pool`` (the reference uses 16
    DataLoader worker processes, train.py:318) or None for a single process."""
    jobs = [(r, camera_parameters) for r in raws]
    if pool is None:
        return [_worker(j) for j in jobs]
    return pool.map(_worker, jobs, chunksize=max(1, len(jobs) // (4 * pool._processes)))
