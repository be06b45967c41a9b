"""torch custom ops over the C ABI + the hand-written autograd.Function of the fused ISP.

Ops are registered under the ``raw2logit_isp`` namespace for the CUDA dispatch key ONLY -- there is no CPU
implementation, no Triton path and no fallback: CPU tensors raise NotImplementedError from the dispatcher, and a
missing ``libr2l_isp.so`` raises ImportError from ``_lib.load()``.

Replaces the 79-node autograd graph the reference records per call (pipeline_torch.py:175-225) with
one forward kernel launch and one backward launch (+ a 1-CTA finish kernel for the 132 parameter gradients).
"""
import ctypes

import torch

from . import _lib

_NS = "raw2logit_isp"
_PARAMS_SCHEMA = ("Tensor black_level, Tensor white_balance, Tensor colour_correction, Tensor gamma_correct, "
                  "Tensor debayer_weight, Tensor sharpen_weight, Tensor gauss_weight, Tensor rgb2yuv, Tensor yuv2rgb")

_library = torch.library.Library(_NS, "DEF")
_library.define(f"forward(Tensor raw, {_PARAMS_SCHEMA}, Tensor? additive, Tensor? affine, float raw_denominator) -> Tensor")
_library.define(f"backward(Tensor raw, {_PARAMS_SCHEMA}, Tensor grad_out, Tensor? grad_scale, bool need_raw_grad, "
                "float raw_denominator) -> (Tensor, Tensor)")
_library.define("mosaic(Tensor raw, Tensor? black_level, bool reduce_size, int out_channels, float raw_denominator) -> Tensor")
_library.define("mosaic_backward(Tensor grad_out, int H, int W, bool reduce_size, int out_channels) -> Tensor")
_library.define("batch_sum(Tensor x, Tensor? scale) -> Tensor")


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _raw_input(raw):
    """(tensor, dtype code).  uint16 is ingested natively; every other dtype is computed in fp32 like the
    reference, whose output buffer is always fp32 (pipeline_torch.py:272)."""
    if raw.dtype == torch.uint16:
        return raw.contiguous(), _lib.U16
    if raw.dtype != torch.float32:
        raw = raw.to(torch.float32)
    return raw.contiguous(), _lib.F32


def _f32c(t, n, name):
    if t.dtype != torch.float32 or not t.is_cuda:
        raise TypeError(f"{name} must be a float32 CUDA tensor, got {t.dtype} on {t.device}")
    if t.numel() != n:
        raise ValueError(f"{name} must have {n} elements, got {tuple(t.shape)}")
    return t.contiguous()


_PARAM_SIZES = (4, 3, 9, 1, 81, 9, 25, 9, 9)


def _pack_params(tensors):
    keep = [_f32c(t, n, name) for t, n, name in zip(tensors, _PARAM_SIZES, _lib.PARAM_FIELDS)]
    return _lib.IspParams(*[t.data_ptr() for t in keep]), keep


def _check_shape(raw):
    if raw.ndim != 3:
        raise AssertionError(f"needs dims (B, H, W), got {raw.shape}")      # pipeline_torch.py:176
    b, h, w = raw.shape
    if h < 3 or w < 3:
        # the reference fails inside the Gaussian's reflect pad (pipeline_torch.py:165, 202)
        raise RuntimeError(f"Padding size should be less than the corresponding input dimension, got H={h}, W={w}")
    return b, h, w


def _forward_cuda(raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, additive, affine, raw_denominator):
    lib = _lib.load()
    b, h, w = _check_shape(raw)
    raw, code = _raw_input(raw)
    with torch.cuda.device(raw.device):
        params, keep = _pack_params((bl, wb, ccm, gamma, wd, ws, wg, m1, m2))
        add = None if additive is None else _f32c(additive, 3 * h * w, "additive")
        aff = None if affine is None else _f32c(affine, 6, "affine")
        tail = _lib.IspTail(None if add is None else add.data_ptr(), None if aff is None else aff.data_ptr())
        out = torch.empty((b, 3, h, w), dtype=torch.float32, device=raw.device)
        rc = lib.r2l_isp_forward(_ptr(raw), code, raw_denominator, b, h, w, ctypes.byref(params),
                                 ctypes.byref(tail), _ptr(out), _stream())
    _lib.check(rc, "r2l_isp_forward")
    return out


def _backward_cuda(raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, grad_out, grad_scale, need_raw_grad, raw_denominator):
    lib = _lib.load()
    b, h, w = _check_shape(raw)
    raw, code = _raw_input(raw)
    with torch.cuda.device(raw.device):
        params, keep = _pack_params((bl, wb, ccm, gamma, wd, ws, wg, m1, m2))
        g = _f32c(grad_out, b * 3 * h * w, "grad_out")
        gs = None if grad_scale is None else _f32c(grad_scale, 3, "grad_scale")
        graw = torch.empty((b, h, w), dtype=torch.float32, device=raw.device) if need_raw_grad else None
        gpar = torch.empty(_lib.NUM_PARAM_GRADS, dtype=torch.float32, device=raw.device)
        nbytes = lib.r2l_isp_backward_workspace_bytes(b, h, w)
        ws_buf = torch.empty(nbytes // 4, dtype=torch.float32, device=raw.device)
        rc = lib.r2l_isp_backward(_ptr(raw), code, raw_denominator, b, h, w, ctypes.byref(params), _ptr(g), _ptr(gs),
                                  _ptr(graw), _ptr(gpar), _ptr(ws_buf), nbytes, _stream())
    _lib.check(rc, "r2l_isp_backward")
    if graw is None:
        graw = torch.empty(0, dtype=torch.float32, device=raw.device)
    return graw, gpar


def _mosaic_cuda(raw, black_level, reduce_size, out_channels, raw_denominator):
    lib = _lib.load()
    assert out_channels in [3, 4]                                             # pipeline_torch.py:252
    if raw.ndim != 3:
        raise ValueError(f"needs dims (B, H, W), got {raw.shape}")
    b, h, w = raw.shape
    if reduce_size and (h % 2 or w % 2):
        # reference: assigning ceil(H/2) rows into an H//2 buffer raises (pipeline_torch.py:261-265)
        raise RuntimeError(f"The expanded size of the tensor must match the existing size: odd H={h} or W={w} "
                           "with reduce_size=True")
    raw, code = _raw_input(raw)
    with torch.cuda.device(raw.device):
        bl = None if black_level is None else _f32c(black_level, 4, "black_level")
        shape = (b, out_channels, h // 2, w // 2) if reduce_size else (b, out_channels, h, w)
        out = torch.empty(shape, dtype=torch.float32, device=raw.device)
        rc = lib.r2l_isp_mosaic(_ptr(raw), code, raw_denominator, b, h, w, _ptr(bl), int(reduce_size), out_channels,
                                _ptr(out), _stream())
    _lib.check(rc, "r2l_isp_mosaic")
    return out


def _mosaic_backward_cuda(grad_out, h, w, reduce_size, out_channels):
    lib = _lib.load()
    b = grad_out.shape[0]
    with torch.cuda.device(grad_out.device):
        g = _f32c(grad_out, grad_out.numel(), "grad_out")
        graw = torch.empty((b, h, w), dtype=torch.float32, device=g.device)
        rc = lib.r2l_isp_mosaic_backward(_ptr(g), b, h, w, int(reduce_size), out_channels, _ptr(graw), _stream())
    _lib.check(rc, "r2l_isp_mosaic_backward")
    return graw


def _batch_sum_cuda(x, scale):
    lib = _lib.load()
    b, c = x.shape[0], x.shape[1]
    hw = x.numel() // max(b * c, 1)
    with torch.cuda.device(x.device):
        xc = _f32c(x, x.numel(), "x")
        sc = None if scale is None else _f32c(scale, c, "scale")
        out = torch.empty((1,) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
        rc = lib.r2l_isp_batch_sum(_ptr(xc), _ptr(sc), b, c, hw, _ptr(out), _stream())
    _lib.check(rc, "r2l_isp_batch_sum")
    return out


_library.impl("forward", _forward_cuda, "CUDA")
_library.impl("backward", _backward_cuda, "CUDA")
_library.impl("mosaic", _mosaic_cuda, "CUDA")
_library.impl("mosaic_backward", _mosaic_backward_cuda, "CUDA")
_library.impl("batch_sum", _batch_sum_cuda, "CUDA")

_ops = getattr(torch.ops, _NS)


class FusedISP(torch.autograd.Function):
    """raw (B,H,W) + 7 parameter tensors + 2 buffers [+ additive, affine] -> (B,3,H,W).

    Saves only ``raw`` and the (tiny) parameters; the backward kernel recomputes the forward per tile.
    Gradients are returned for raw (if needed), the 7 parameter tensors and the additive layer; the colour-space
    buffers and the affine tail (BatchNorm running statistics) get none, as in the reference.
    """

    @staticmethod
    def forward(ctx, raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, additive, affine, raw_denominator):
        out = _ops.forward(raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, additive, affine, raw_denominator)
        ctx.save_for_backward(raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, affine)
        ctx.raw_denominator = raw_denominator
        ctx.has_additive = additive is not None
        ctx.additive_shape = None if additive is None else additive.shape
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, affine = ctx.saved_tensors
        need_raw = ctx.needs_input_grad[0]
        scale = None if affine is None else affine[:3]
        grad_out = grad_out.contiguous()
        grads = [None] * 13
        if need_raw or any(ctx.needs_input_grad[1:8]):
            graw, gpar = _ops.backward(raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, grad_out, scale, need_raw,
                                       ctx.raw_denominator)
            if need_raw:
                grads[0] = graw if raw.dtype == torch.float32 else graw.to(raw.dtype)
            for slot, name in enumerate(_lib.PARAM_FIELDS[:7], start=1):
                if ctx.needs_input_grad[slot]:
                    off, n, shape = _lib.GRAD_LAYOUT[name]
                    grads[slot] = gpar[off:off + n].view(shape)
        if ctx.has_additive and ctx.needs_input_grad[10]:
            grads[10] = _ops.batch_sum(grad_out, scale).view(ctx.additive_shape)
        return tuple(grads)


class Mosaic(torch.autograd.Function):
    """CFA split (raw2rgb, pipeline_torch.py:240-283) without black level; differentiable w.r.t. raw."""

    @staticmethod
    def forward(ctx, raw, reduce_size, out_channels, raw_denominator):
        ctx.hw = (raw.shape[1], raw.shape[2])
        ctx.cfg = (bool(reduce_size), int(out_channels))
        ctx.raw_dtype = raw.dtype
        return _ops.mosaic(raw, None, bool(reduce_size), int(out_channels), raw_denominator)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        g = _ops.mosaic_backward(grad_out.contiguous(), ctx.hw[0], ctx.hw[1], ctx.cfg[0], ctx.cfg[1])
        return g if ctx.raw_dtype == torch.float32 else g.to(ctx.raw_dtype), None, None, None


def fused_isp(raw, black_level, white_balance, colour_correction, gamma_correct, debayer_weight, sharpen_weight,
              gauss_weight, rgb2yuv, yuv2rgb, additive=None, affine=None, raw_denominator=65535.0):
    return FusedISP.apply(raw, black_level, white_balance, colour_correction, gamma_correct, debayer_weight,
                          sharpen_weight, gauss_weight, rgb2yuv, yuv2rgb, additive, affine, float(raw_denominator))
