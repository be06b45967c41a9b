"""ctypes binding of the C ABI in include/r2l_isp.h.  No fallback: a missing library is a hard error."""
import ctypes
import os

from . import _build

PARAM_FIELDS = ("black_level", "white_balance", "colour_correction", "gamma_correct", "debayer_weight",
                "sharpen_weight", "gauss_weight", "rgb2yuv", "yuv2rgb")
NUM_PARAM_GRADS = 132
# name -> (offset, length, shape) inside the flat gradient vector (R2L_G_* in r2l_isp.h)
GRAD_LAYOUT = {
    "black_level": (0, 4, (4,)),
    "white_balance": (4, 3, (1, 3)),
    "colour_correction": (7, 9, (3, 3)),
    "gamma_correct": (16, 1, (1,)),
    "debayer_weight": (17, 81, (3, 3, 3, 3)),
    "sharpen_weight": (98, 9, (1, 1, 3, 3)),
    "gauss_weight": (107, 25, (1, 1, 5, 5)),
}
F32, U16 = 0, 1
ABI_VERSION = 6
EPOCH_DEVICE = 0xFFFFFFFF   # R2L_EPOCH_DEVICE: the kernel keeps the exchange epoch itself (graph-capturable)

EXPORTS = ("r2l_isp_abi_version", "r2l_isp_error_string", "r2l_isp_last_cuda_error", "r2l_isp_forward",
           "r2l_isp_workspace_bytes", "r2l_isp_forward_bn_train", "r2l_isp_bn_backward_prepare", "r2l_isp_backward",
           "r2l_isp_mosaic", "r2l_isp_mosaic_backward", "r2l_isp_batch_sum", "r2l_isp_saved_luma_floats",
           "r2l_isp_luma_supported", "r2l_isp_exchange_bytes", "r2l_isp_backward_dp", "r2l_isp_ssim_partial_count",
           "r2l_isp_ssim_forward", "r2l_isp_ssim_backward", "r2l_isp_numpy_forward", "r2l_isp_dihedral_copy",
           "r2l_isp_stage_workspace_bytes", "r2l_isp_stage_conv", "r2l_isp_stage_conv_backward", "r2l_isp_stage_clip",
           "r2l_isp_stage_clip_backward", "r2l_isp_stage_gamma", "r2l_isp_stage_gamma_backward")


class IspAllreduce(ctypes.Structure):
    """r2l_isp_allreduce (include/r2l_isp.h): the data-parallel exchange fused into the backward kernel."""
    _fields_ = [("world", ctypes.c_int), ("rank", ctypes.c_int), ("peers", ctypes.c_void_p), ("epoch", ctypes.c_uint),
                ("scale", ctypes.c_float)]


class IspParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in PARAM_FIELDS]


class IspTail(ctypes.Structure):
    _fields_ = [("additive", ctypes.c_void_p), ("affine", ctypes.c_void_p)]


_LIB = None


def load():
    """Returns the loaded CDLL.  Raises (never falls back) when the library has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("R2L_ISP_LIB") or _build.LIB_PATH      # R2L_ISP_LIB: another build of the same ABI (A/B timing)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: the raw2logit_b200 CUDA library has not been built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU fallback.")
    lib = ctypes.CDLL(path)
    vp, ci, cf, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
    lib.r2l_isp_abi_version.restype = ci
    lib.r2l_isp_error_string.restype = ctypes.c_char_p
    lib.r2l_isp_error_string.argtypes = [ci]
    lib.r2l_isp_last_cuda_error.restype = ci
    lib.r2l_isp_forward.restype = ci
    lib.r2l_isp_forward.argtypes = [vp, ci, cf, ci, ci, ci, ctypes.POINTER(IspParams), ctypes.POINTER(IspTail), vp, vp, vp]
    lib.r2l_isp_saved_luma_floats.restype = sz
    lib.r2l_isp_saved_luma_floats.argtypes = [ci, ci, ci]
    lib.r2l_isp_luma_supported.restype = ci
    lib.r2l_isp_luma_supported.argtypes = [vp, ci, ci, ci, ci, vp, vp]
    lib.r2l_isp_workspace_bytes.restype = sz
    lib.r2l_isp_workspace_bytes.argtypes = [ci, ci, ci]
    lib.r2l_isp_forward_bn_train.restype = ci
    lib.r2l_isp_forward_bn_train.argtypes = [vp, ci, cf, ci, ci, ci, ctypes.POINTER(IspParams), vp, vp, vp, vp, vp, cf, cf,
                                             vp, vp, vp, sz, vp]
    lib.r2l_isp_bn_backward_prepare.restype = ci
    lib.r2l_isp_bn_backward_prepare.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp, sz, vp]
    lib.r2l_isp_backward.restype = ci
    lib.r2l_isp_backward.argtypes = [vp, ci, cf, ci, ci, ci, ctypes.POINTER(IspParams), vp, vp, vp, vp, vp, vp, vp, vp, sz,
                                     vp]
    lib.r2l_isp_backward_dp.restype = ci
    lib.r2l_isp_backward_dp.argtypes = [vp, ci, cf, ci, ci, ci, ctypes.POINTER(IspParams), vp, vp, vp, vp, vp, vp, vp, vp,
                                        sz, ctypes.POINTER(IspAllreduce), vp]
    lib.r2l_isp_exchange_bytes.restype = sz
    lib.r2l_isp_exchange_bytes.argtypes = [ci]
    lib.r2l_isp_mosaic.restype = ci
    lib.r2l_isp_mosaic.argtypes = [vp, ci, cf, ci, ci, ci, vp, ci, ci, vp, vp]
    lib.r2l_isp_mosaic_backward.restype = ci
    lib.r2l_isp_mosaic_backward.argtypes = [vp, ci, ci, ci, ci, ci, vp, vp]
    lib.r2l_isp_batch_sum.restype = ci
    lib.r2l_isp_batch_sum.argtypes = [vp, vp, ci, ci, ci, vp, vp]
    lib.r2l_isp_dihedral_copy.restype = ci
    lib.r2l_isp_dihedral_copy.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_longlong), vp, ci, ctypes.POINTER(ctypes.c_longlong),
                                          ci, ci, ci, ci, ci, ci, ctypes.POINTER(ci), vp]
    lib.r2l_isp_numpy_forward.restype = ci
    lib.r2l_isp_numpy_forward.argtypes = [vp, ci, cf, ci, ci, ci, ctypes.POINTER(cf), ctypes.POINTER(cf), ctypes.POINTER(cf),
                                          ci, ci, cf, cf, vp, vp]
    lib.r2l_isp_ssim_partial_count.restype = sz
    lib.r2l_isp_ssim_partial_count.argtypes = [ci, ci, ci, ci]
    lib.r2l_isp_ssim_forward.restype = ci
    lib.r2l_isp_ssim_forward.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, vp]
    lib.r2l_isp_ssim_backward.restype = ci
    lib.r2l_isp_ssim_backward.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, vp, vp, vp]
    if lib.r2l_isp_abi_version() != ABI_VERSION:
        raise ImportError(f"{path}: ABI version {lib.r2l_isp_abi_version()} != expected {ABI_VERSION}; rebuild")
    _LIB = lib
    return lib


def check(rc, what):
    if rc != 0:
        lib = load()
        msg = lib.r2l_isp_error_string(rc).decode()
        extra = f" (cudaError {lib.r2l_isp_last_cuda_error()})" if rc == -6 else ""
        raise RuntimeError(f"{what} failed: {msg}{extra}")
