"""torch custom ops over the C ABI + the hand-written autograd.Function of the fused ISP.

Ops are registered under the ``raw2logit_isp`` namespace for the CUDA dispatch key ONLY -- there is no CPU
implementation, no Triton path and no fallback: CPU tensors raise NotImplementedError from the dispatcher, and a
missing ``libr2l_isp.so`` raises ImportError from ``_lib.load()``.

Replaces the 79-node autograd graph the reference records per call (pipeline_torch.py:175-225) with
one forward kernel launch and one backward launch (+ a 1-CTA finish kernel for the 132 parameter gradients).
"""
import ctypes
import os

import torch

from . import _lib

_NS = "raw2logit_isp"
_PARAMS_SCHEMA = ("Tensor black_level, Tensor white_balance, Tensor colour_correction, Tensor gamma_correct, "
                  "Tensor debayer_weight, Tensor sharpen_weight, Tensor gauss_weight, Tensor rgb2yuv, Tensor yuv2rgb")

_library = torch.library.Library(_NS, "DEF")
_library.define(f"forward(Tensor raw, {_PARAMS_SCHEMA}, Tensor? additive, Tensor? affine, float raw_denominator, "
                "bool save_luma=False) -> (Tensor, Tensor)")
_library.define(f"forward_bn_train(Tensor raw, {_PARAMS_SCHEMA}, Tensor? additive, Tensor(a!)? running_mean, "
                "Tensor(b!)? running_var, float momentum, float eps, float raw_denominator, bool save_luma=False) "
                "-> (Tensor, Tensor, Tensor)")
_library.define("bn_backward_prepare(Tensor grad_out, Tensor out, Tensor saved_affine) -> Tensor")
_library.define(f"backward(Tensor raw, {_PARAMS_SCHEMA}, Tensor grad_out, Tensor? grad_tail, Tensor? additive, "
                "Tensor? out, Tensor? luma, bool need_raw_grad, float raw_denominator) -> (Tensor, Tensor)")
_library.define("mosaic(Tensor raw, Tensor? black_level, bool reduce_size, int out_channels, float raw_denominator) -> Tensor")
_library.define("mosaic_backward(Tensor grad_out, int H, int W, bool reduce_size, int out_channels) -> Tensor")
_library.define("batch_sum(Tensor x, Tensor? scale) -> Tensor")


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    """The current CUDA stream of the current device as a cudaStream_t (the raw-handle query: torch.cuda.current_stream()
    builds a Stream object and costs ~15 us of host time per call, a tenth of the module-path step)."""
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def _on_device(device):
    """Device guard for the launch: nothing to switch (the usual case) costs nothing, torch.cuda.device() ~10 us."""
    return _NO_GUARD if device.index == torch.cuda.current_device() else torch.cuda.device(device)


_workspaces = {}

# data-parallel gradient exchange fused into the backward kernel (set by parallel.enable_fused_gradient_exchange)
_exchange = None
_exchange_average = True


def set_gradient_exchange(exchange, average=True):
    """``exchange``: a ``parallel.PeerExchange`` (every later fused backward on this process sums / averages its 132
    parameter gradients over the ranks inside the kernel) or None (local gradients).  Every rank must run the same
    sequence of backward calls while it is set."""
    global _exchange, _exchange_average
    _exchange, _exchange_average = exchange, bool(average)


def _workspace(lib, b, h, w, device):
    """One scratch buffer per (device, stream), kept for the life of the process: r2l_isp_workspace_bytes is a fixed upper
    bound, and calls on one stream run in order, so they can share it."""
    key = (device.index, torch._C._cuda_getCurrentRawStream(device.index))
    nbytes = lib.r2l_isp_workspace_bytes(b, h, w)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() * 4 < nbytes:
        buf = torch.empty(nbytes // 4, dtype=torch.float32, device=device)
        _workspaces[key] = buf
    return buf, nbytes


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


def _luma_buffer(lib, raw, code, b, h, w, out, add, save_luma):
    """The (2, ceil(B/2), H, W, 2) tensor for the Y0 / Y1 planes the forward keeps for the backward, or an empty
    tensor when not asked for / when the call does not take the kernel that writes them (r2l_isp_luma_supported)."""
    if save_luma and b > 0 and lib.r2l_isp_luma_supported(_ptr(raw), code, b, h, w, _ptr(out), _ptr(add)):
        return torch.empty((2, (b + 1) // 2, h, w, 2), dtype=torch.float32, device=raw.device)
    return torch.empty(0, dtype=torch.float32, device=raw.device)


def _forward_cuda(raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, additive, affine, raw_denominator, save_luma=False):
    lib = _lib.load()
    b, h, w = _check_shape(raw)
    raw, code = _raw_input(raw)
    with _on_device(raw.device):
        params, keep = _pack_params((bl, wb, ccm, gamma, wd, ws, wg, m1, m2))
        add = None if additive is None else _f32c(additive, 3 * h * w, "additive")
        aff = None if affine is None else _f32c(affine, 6, "affine")
        tail = _lib.IspTail(None if add is None else add.data_ptr(), None if aff is None else aff.data_ptr())
        out = torch.empty((b, 3, h, w), dtype=torch.float32, device=raw.device)
        luma = _luma_buffer(lib, raw, code, b, h, w, out, add, save_luma)
        rc = lib.r2l_isp_forward(_ptr(raw), code, raw_denominator, b, h, w, ctypes.byref(params),
                                 ctypes.byref(tail), _ptr(out), _ptr(luma) if luma.numel() else None, _stream())
    _lib.check(rc, "r2l_isp_forward")
    return out, luma


def _forward_bn_train_cuda(raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, additive, running_mean, running_var,
                           momentum, eps, raw_denominator, save_luma=False):
    lib = _lib.load()
    b, h, w = _check_shape(raw)
    if b * h * w < 2:
        raise ValueError("Expected more than 1 value per channel when training")
    raw, code = _raw_input(raw)
    with _on_device(raw.device):
        params, keep = _pack_params((bl, wb, ccm, gamma, wd, ws, wg, m1, m2))
        add = None if additive is None else _f32c(additive, 3 * h * w, "additive")
        for t, name in ((running_mean, "running_mean"), (running_var, "running_var")):
            if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != 3):
                raise TypeError(f"{name} must be a contiguous float32 tensor with 3 elements")
        out = torch.empty((b, 3, h, w), dtype=torch.float32, device=raw.device)
        saved = torch.empty(6, dtype=torch.float32, device=raw.device)
        ws_buf, nbytes = _workspace(lib, b, h, w, raw.device)
        luma = _luma_buffer(lib, raw, code, b, h, w, out, add, save_luma)
        rc = lib.r2l_isp_forward_bn_train(_ptr(raw), code, raw_denominator, b, h, w, ctypes.byref(params), _ptr(add),
                                          _ptr(out), _ptr(running_mean), _ptr(running_var), momentum, eps,
                                          _ptr(saved), _ptr(luma) if luma.numel() else None, _ptr(ws_buf), nbytes,
                                          _stream())
    _lib.check(rc, "r2l_isp_forward_bn_train")
    return out, saved, luma


def _bn_backward_prepare_cuda(grad_out, out, saved_affine):
    lib = _lib.load()
    b, _, h, w = out.shape
    with _on_device(out.device):
        g = _f32c(grad_out, out.numel(), "grad_out")
        y = _f32c(out, out.numel(), "out")
        sa = _f32c(saved_affine, 6, "saved_affine")
        tail = torch.empty(15, dtype=torch.float32, device=out.device)
        ws_buf, nbytes = _workspace(lib, b, h, w, out.device)
        rc = lib.r2l_isp_bn_backward_prepare(_ptr(g), _ptr(y), _ptr(sa), b, h, w, _ptr(tail), _ptr(ws_buf), nbytes,
                                             _stream())
    _lib.check(rc, "r2l_isp_bn_backward_prepare")
    return tail


def _backward_cuda(raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, grad_out, grad_tail, additive, out, luma, need_raw_grad,
                   raw_denominator):
    lib = _lib.load()
    b, h, w = _check_shape(raw)
    raw, code = _raw_input(raw)
    with _on_device(raw.device):
        params, keep = _pack_params((bl, wb, ccm, gamma, wd, ws, wg, m1, m2))
        g = _f32c(grad_out, b * 3 * h * w, "grad_out")
        gs = None if grad_tail is None else _f32c(grad_tail, 15, "grad_tail")
        add = None if additive is None else _f32c(additive, 3 * h * w, "additive")
        y = None if out is None else _f32c(out, b * 3 * h * w, "out")
        lum = None
        if luma is not None and luma.numel():
            lum = _f32c(luma, lib.r2l_isp_saved_luma_floats(b, h, w), "luma")
        graw = torch.empty((b, h, w), dtype=torch.float32, device=raw.device) if need_raw_grad else None
        gpar = torch.empty(_lib.NUM_PARAM_GRADS, dtype=torch.float32, device=raw.device)
        ws_buf, nbytes = _workspace(lib, b, h, w, raw.device)
        if _exchange is not None:
            # data-parallel: the 132 gradients leave the kernel already reduced over the ranks (parallel.PeerExchange)
            if y is None or lum is None:
                raise RuntimeError("the fused gradient exchange needs the saved output and luma planes (unset "
                                   "R2L_ISP_RECOMPUTE / R2L_ISP_NO_LUMA, shapes with W % 4 == 0)")
            desc = _exchange.next(average=_exchange_average)
            rc = lib.r2l_isp_backward_dp(_ptr(raw), code, raw_denominator, b, h, w, ctypes.byref(params), _ptr(g),
                                         _ptr(gs), _ptr(add), _ptr(y), _ptr(lum), _ptr(graw), _ptr(gpar), _ptr(ws_buf),
                                         nbytes, ctypes.byref(desc), _stream())
        else:
            rc = lib.r2l_isp_backward(_ptr(raw), code, raw_denominator, b, h, w, ctypes.byref(params), _ptr(g), _ptr(gs),
                                      _ptr(add), _ptr(y), _ptr(lum), _ptr(graw), _ptr(gpar), _ptr(ws_buf), nbytes,
                                      _stream())
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
    with _on_device(raw.device):
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
    with _on_device(grad_out.device):
        g = _f32c(grad_out, grad_out.numel(), "grad_out")
        graw = torch.empty((b, h, w), dtype=torch.float32, device=g.device)
        rc = lib.r2l_isp_mosaic_backward(_ptr(g), b, h, w, int(reduce_size), out_channels, _ptr(graw), _stream())
    _lib.check(rc, "r2l_isp_mosaic_backward")
    return graw


def _batch_sum_cuda(x, scale):
    lib = _lib.load()
    b, c = x.shape[0], x.shape[1]
    hw = x.numel() // max(b * c, 1)
    with _on_device(x.device):
        xc = _f32c(x, x.numel(), "x")
        sc = None if scale is None else _f32c(scale, c, "scale")
        out = torch.empty((1,) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
        rc = lib.r2l_isp_batch_sum(_ptr(xc), _ptr(sc), b, c, hw, _ptr(out), _stream())
    _lib.check(rc, "r2l_isp_batch_sum")
    return out


_library.impl("forward", _forward_cuda, "CUDA")
_library.impl("forward_bn_train", _forward_bn_train_cuda, "CUDA")
_library.impl("bn_backward_prepare", _bn_backward_prepare_cuda, "CUDA")
_library.impl("backward", _backward_cuda, "CUDA")
_library.impl("mosaic", _mosaic_cuda, "CUDA")
_library.impl("mosaic_backward", _mosaic_backward_cuda, "CUDA")
_library.impl("batch_sum", _batch_sum_cuda, "CUDA")

_ops = getattr(torch.ops, _NS)


class FusedISP(torch.autograd.Function):
    """raw (B,H,W) + 7 parameter tensors + 2 buffers [+ additive] [+ BatchNorm tail] -> (B,3,H,W).

    Saves ``raw``, the (tiny) parameters, the output tensor (which the consumer of the output keeps alive anyway)
    and the two luma planes Y0 / Y1 the forward kernel computes on the way (8 B/px, what autograd would keep for the
    two convolutions): the backward kernel then recomputes nothing -- the clip mask and the gamma derivative are read
    off the saved output, the weight statistics read their Y1 / Y0 / raw centres from memory.
    ``R2L_ISP_NO_LUMA=1`` keeps only the output (third-generation backward: Y0 / Y1 rebuilt per tile from ``raw``),
    ``R2L_ISP_RECOMPUTE=1`` recomputes everything from ``raw``.  Gradients are returned for raw (if needed), the 7 parameter
    tensors and the additive layer; the colour-space buffers and the BatchNorm statistics get none, as in the
    reference.

    bn_mode: 0 = no tail, 1 = eval (affine from running statistics), 2 = train (batch statistics, running
    statistics updated in place by the kernel).
    """

    @staticmethod
    def forward(ctx, raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, additive, bn_mode, running_mean, running_var,
                momentum, eps, raw_denominator):
        params = (bl, wb, ccm, gamma, wd, ws, wg, m1, m2)
        saved_affine = None
        # the backward reads the luma planes only together with the saved output (R2L_ISP_RECOMPUTE=1: neither)
        save_luma = (os.environ.get("R2L_ISP_RECOMPUTE", "0") == "0" and os.environ.get("R2L_ISP_NO_LUMA", "0") != "1"
                     and any(ctx.needs_input_grad[:8]))
        if bn_mode == 2:
            out, saved_affine, luma = _ops.forward_bn_train(raw, *params, additive, running_mean, running_var,
                                                            momentum, eps, raw_denominator, save_luma)
            ctx.mark_non_differentiable(saved_affine)
        elif bn_mode == 1:
            scale = torch.rsqrt(running_var + eps)
            saved_affine = torch.cat([scale, -running_mean * scale])
            out, luma = _ops.forward(raw, *params, additive, saved_affine, raw_denominator, save_luma)
        else:
            out, luma = _ops.forward(raw, *params, additive, None, raw_denominator, save_luma)
        ctx.save_for_backward(raw, *params, additive, saved_affine, out, luma)
        ctx.bn_mode = bn_mode
        ctx.raw_denominator = raw_denominator
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, additive, saved_affine, out_saved, luma = ctx.saved_tensors
        params = (bl, wb, ccm, gamma, wd, ws, wg, m1, m2)
        need_raw = ctx.needs_input_grad[0]
        grad_out = grad_out.contiguous()
        grads = [None] * 17
        # 15-float description of the tail the forward applied: {gs, c1, c2, ysc, ysh} x 3 channels (r2l_isp.h)
        tail = None
        if ctx.bn_mode == 2:
            tail = _ops.bn_backward_prepare(grad_out, out_saved, saved_affine)
        elif ctx.bn_mode == 1:
            zeros = saved_affine.new_zeros(6)
            tail = torch.cat([saved_affine[:3], zeros, saved_affine])        # gs = ysc = scale, c1 = c2 = 0, ysh = shift
        elif additive is not None:
            one, zero = grad_out.new_ones(3), grad_out.new_zeros(3)
            tail = torch.cat([one, zero, zero, one, zero])                   # identity affine around the additive layer
        use_out = os.environ.get("R2L_ISP_RECOMPUTE", "0") != "1"
        if need_raw or any(ctx.needs_input_grad[1:8]):
            graw, gpar = _ops.backward(raw, *params, grad_out, tail, additive if tail is not None else None,
                                       out_saved if use_out else None, luma if use_out else None, need_raw,
                                       ctx.raw_denominator)
            if need_raw:
                grads[0] = graw if raw.dtype == torch.float32 else graw.to(raw.dtype)
            for slot, name in enumerate(_lib.PARAM_FIELDS[:7], start=1):
                if ctx.needs_input_grad[slot]:
                    off, n, shape = _lib.GRAD_LAYOUT[name]
                    grads[slot] = gpar[off:off + n].view(shape)
        if additive is not None and ctx.needs_input_grad[10]:
            if ctx.bn_mode == 2:
                # d/d(additive) = sum_b gs*(G - c1 - c2*yhat); yhat is the saved output
                gs, c1, c2 = (tail[0:3].view(1, 3, 1, 1), tail[3:6].view(1, 3, 1, 1), tail[6:9].view(1, 3, 1, 1))
                geff = gs * (grad_out - c1 - c2 * out_saved)
                grads[10] = _ops.batch_sum(geff, None).view(additive.shape)
            else:
                scale = None if saved_affine is None else saved_affine[:3]
                grads[10] = _ops.batch_sum(grad_out, scale).view(additive.shape)
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
              gauss_weight, rgb2yuv, yuv2rgb, additive=None, bn_mode=0, running_mean=None, running_var=None,
              momentum=0.1, eps=1e-5, raw_denominator=65535.0):
    return FusedISP.apply(raw, black_level, white_balance, colour_correction, gamma_correct, debayer_weight,
                          sharpen_weight, gauss_weight, rgb2yuv, yuv2rgb, additive, int(bn_mode), running_mean,
                          running_var, float(momentum), float(eps), float(raw_denominator))
