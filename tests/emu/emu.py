"""ctypes wrapper around the host emulation of the CUDA CTA code (tests/emu/isp_emu.cpp) -- TEST ONLY."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "isp_emu.cpp")
LIB = os.path.join(HERE, "libisp_emu.so")
DEPS = [SRC, os.path.join(ROOT, "raw2logit_b200", "csrc", "isp_core.cuh"),
        os.path.join(ROOT, "raw2logit_b200", "csrc", "isp_fwd2.cuh"),
        os.path.join(ROOT, "raw2logit_b200", "csrc", "isp_bwd3.cuh"),
        os.path.join(ROOT, "raw2logit_b200", "csrc", "isp_fwd3.cuh"),
        os.path.join(ROOT, "raw2logit_b200", "csrc", "isp_bwd4.cuh"),
        os.path.join(ROOT, "raw2logit_b200", "csrc", "isp_bwd5.cuh"),
        os.path.join(ROOT, "raw2logit_b200", "csrc", "isp_config.h"), os.path.join(ROOT, "include", "r2l_isp.h")]

PARAM_FIELDS = ["black_level", "white_balance", "colour_correction", "gamma_correct", "debayer.weight",
                "sharpening_filter.weight", "gaussian_blur.weight", "M_RGB_2_YUV", "M_YUV_2_RGB"]
GRAD_SLICES = {"black_level": (0, 4), "white_balance": (4, 7), "colour_correction": (7, 16),
               "gamma_correct": (16, 17), "debayer.weight": (17, 98), "sharpening_filter.weight": (98, 107),
               "gaussian_blur.weight": (107, 132)}


class Params(ctypes.Structure):
    _fields_ = [(n.replace(".", "_"), ctypes.c_void_p) for n in PARAM_FIELDS]


class Tail(ctypes.Structure):
    _fields_ = [("additive", ctypes.c_void_p), ("affine", ctypes.c_void_p)]


def build():
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS):
        return LIB
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _params(state):
    keep = [_f32(state[k].detach().numpy() if hasattr(state[k], "detach") else state[k]) for k in PARAM_FIELDS]
    p = Params(*[a.ctypes.data for a in keep])
    return p, keep


def forward(raw, state, additive=None, affine=None, n_cta=3, denom=65535.0, version=3, chan_sums=None, luma=None):
    """luma: optional float32 array (2, ceil(B/2), H, W, 2) that receives the saved Y0 / Y1 planes (version 3 only)."""
    raw = np.ascontiguousarray(raw)
    dtype = 1 if raw.dtype == np.uint16 else 0
    if dtype == 0:
        raw = _f32(raw)
    b, h, w = raw.shape
    p, keep = _params(state)
    add = None if additive is None else _f32(additive)
    aff = None if affine is None else _f32(affine)
    tail = Tail(None if add is None else add.ctypes.data, None if aff is None else aff.ctypes.data)
    out = np.full((b, 3, h, w), np.nan, dtype=np.float32)
    rc = lib().emu_isp_forward(ctypes.c_void_p(raw.ctypes.data), dtype, ctypes.c_float(denom), b, h, w,
                               ctypes.byref(p), ctypes.byref(tail), ctypes.c_void_p(out.ctypes.data), n_cta, version,
                               None if chan_sums is None else ctypes.c_void_p(chan_sums.ctypes.data),
                               None if luma is None else ctypes.c_void_p(luma.ctypes.data))
    assert rc == 0, rc
    return out


def backward(raw, state, grad_out, need_raw_grad=True, n_cta=3, denom=65535.0, grad_tail=None, additive=None,
             version=3, out=None, luma=None):
    raw = np.ascontiguousarray(raw)
    dtype = 1 if raw.dtype == np.uint16 else 0
    if dtype == 0:
        raw = _f32(raw)
    b, h, w = raw.shape
    p, keep = _params(state)
    g = _f32(grad_out)
    graw = np.full((b, h, w), np.nan, dtype=np.float32) if need_raw_grad else None
    gpar = np.full(132, np.nan, dtype=np.float32)
    gs = None if grad_tail is None else _f32(grad_tail)
    add = None if additive is None else _f32(additive)
    outc = None if out is None else _f32(out)
    lumac = None if luma is None else _f32(luma)
    rc = lib().emu_isp_backward(ctypes.c_void_p(raw.ctypes.data), dtype, ctypes.c_float(denom), b, h, w,
                                ctypes.byref(p), ctypes.c_void_p(g.ctypes.data),
                                None if gs is None else ctypes.c_void_p(gs.ctypes.data),
                                None if add is None else ctypes.c_void_p(add.ctypes.data),
                                None if graw is None else ctypes.c_void_p(graw.ctypes.data),
                                ctypes.c_void_p(gpar.ctypes.data), n_cta, version,
                                None if outc is None else ctypes.c_void_p(outc.ctypes.data),
                                None if lumac is None else ctypes.c_void_p(lumac.ctypes.data))
    assert rc == 0, rc
    grads = {k: gpar[a:b_] for k, (a, b_) in GRAD_SLICES.items()}
    if graw is not None:
        grads["raw"] = graw
    return grads
