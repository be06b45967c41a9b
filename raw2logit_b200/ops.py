"""torch custom ops of the fused ISP: ``torch.ops.raw2logit_isp.*``.

The operators and the autograd node of the fused ISP are C++ (``csrc_torch/r2l_torch.cpp``: ``TORCH_LIBRARY`` +
``torch::autograd::Function``, built in-tree into ``libr2l_torch.so``) over the C ABI of ``libr2l_isp.so``
(``include/r2l_isp.h``); this module only loads them and keeps the keyword-argument entry points the module classes
call.  Ops are registered for the CUDA dispatch key ONLY -- there is no CPU implementation, no Triton path and no
fallback: CPU tensors raise NotImplementedError from the dispatcher, and a missing library raises ImportError here.

Replaces the 79-node autograd graph the reference records per call (pipeline_torch.py:175-225) with one
dispatcher call, one forward kernel launch and one backward launch.

Operators (schemas in r2l_torch.cpp): ``forward``, ``forward_bn_train``, ``bn_backward_prepare``, ``backward``,
``mosaic``, ``mosaic_backward``, ``batch_sum`` (plain, one C-ABI call each); ``fused`` and ``mosaic_ad`` (differentiable:
C++ autograd nodes); ``set_exchange`` (state of the fused data-parallel gradient exchange).
"""
import os

import torch

from . import _build, _lib

_NS = "raw2logit_isp"


def _load_ops():
    _lib.load()                                     # libr2l_isp.so first: a clear ImportError when it was never built
    path = _build.TORCH_LIB_PATH
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: the torch operator shim has not been built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs g++ and nvcc). There is no fallback.")
    torch.ops.load_library(path)
    return getattr(torch.ops, _NS)


_ops = _load_ops()

# data-parallel gradient exchange fused into the backward kernel (set by parallel.enable_fused_gradient_exchange)
_exchange = None
_exchange_average = True


def set_gradient_exchange(exchange, average=True):
    """``exchange``: a ``parallel.PeerExchange`` (every later fused backward on this process sums / averages its 132
    parameter gradients over the ranks inside the kernel) or None (local gradients).  Every rank must run the same
    sequence of backward calls while it is set.  The kernel keeps the exchange epoch in the buffer itself
    (``R2L_EPOCH_DEVICE``), so the step may be captured in a CUDA graph and replayed."""
    global _exchange, _exchange_average
    _exchange, _exchange_average = exchange, bool(average)
    if exchange is None or not hasattr(exchange, "peers"):
        _ops.set_exchange(1, 0, 0, 1.0, False)
    else:
        _ops.set_exchange(exchange.world, exchange.rank, exchange.peers, 1.0 / exchange.world if average else 1.0, True)


def fused_isp(raw, black_level, white_balance, colour_correction, gamma_correct, debayer_weight, sharpen_weight,
              gauss_weight, rgb2yuv, yuv2rgb, additive=None, bn_mode=0, running_mean=None, running_var=None,
              momentum=0.1, eps=1e-5, raw_denominator=65535.0, num_batches_tracked=None):
    """raw (B,H,W) + 7 parameter tensors + 2 buffers [+ additive] [+ BatchNorm tail] -> (B,3,H,W), differentiable in
    raw, the 7 parameter tensors and the additive layer (``FusedISPFn`` in r2l_torch.cpp).
    bn_mode: 0 = no tail, 1 = eval (affine from running statistics), 2 = train (batch statistics, running statistics
    updated in place by the kernel; ``num_batches_tracked``, when given, is incremented by the kernel as well)."""
    return _ops.fused(raw, black_level, white_balance, colour_correction, gamma_correct, debayer_weight,
                      sharpen_weight, gauss_weight, rgb2yuv, yuv2rgb, additive, int(bn_mode), running_mean,
                      running_var, num_batches_tracked, float(momentum), float(eps), float(raw_denominator))


def mosaic(raw, black_level=None, reduce_size=True, out_channels=3, raw_denominator=65535.0):
    """CFA split (raw2rgb, pipeline_torch.py:240-283), differentiable in raw and black_level like the reference."""
    return _ops.mosaic_ad(raw, black_level, bool(reduce_size), int(out_channels), float(raw_denominator))
