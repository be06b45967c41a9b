"""Staged mode of the parametrized ISP (``track_stages=True``): one repo kernel per stage, every stage its own tensor.

Reference: ``ParametrizedProcessing.forward`` with ``track_stages=True`` (pipeline_torch.py:183-221) keeps
``stages['demosaic' | 'color_correct' | 'sharpening' | 'gaussian' | 'clipped' | 'gamma_correct' (| 'noise')]`` in the
autograd graph so that ``model.track_images`` (model.py:229-254) can read each stage and its ``.grad``.  The fused
kernels never materialise them, so this mode runs the chain stage by stage -- on the kernels of ``csrc/isp_stages.cu``
(``torch.ops.raw2logit_isp.stage_*``), not on stock ATen / cuDNN ops:

* every linear stage is ONE 3 -> 3 channel K x K correlation of the stage before it; its combined weight is formed here
  from the parameters with a few tiny differentiable torch ops (einsum over <= 225 numbers), so autograd carries the
  kernel's weight gradient back to white balance / colour matrix / demosaic taps / sharpening / Gaussian taps;
* clip and gamma are pointwise kernels with hand-written backward (gamma also reduces d/dgamma);
* CFA split + black level is the ``mosaic`` operator the ``RawToRGB`` class already uses.

The YUV -> RGB -> YUV round trip between the sharpening and the Gaussian stage (:197-200) is part of the chain, as in
the reference.  This is an inspection path (a few images per epoch): simple kernels, deterministic reductions.
"""
import torch

from . import ops

_ops = ops._ops


class _StageConv(torch.autograd.Function):
    """y = corr(pad(x), weight): x (B,3,H,W), weight (3,3,K,K), reflect or zero padding K//2."""

    @staticmethod
    def forward(ctx, x, weight, reflect):
        ctx.save_for_backward(x, weight)
        ctx.reflect = bool(reflect)
        return _ops.stage_conv(x, weight, bool(reflect))

    @staticmethod
    def backward(ctx, grad_y):
        x, weight = ctx.saved_tensors
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_x or need_w):
            return None, None, None
        gx, gw = _ops.stage_conv_backward(x, weight, grad_y.contiguous(), ctx.reflect, need_x, need_w)
        return (gx if need_x else None), (gw if need_w else None), None


class _StageClip(torch.autograd.Function):
    """torch.clip(x, lo, hi) (:206): the gradient passes where lo <= x <= hi."""

    @staticmethod
    def forward(ctx, x, lo, hi):
        ctx.save_for_backward(x)
        ctx.lo, ctx.hi = float(lo), float(hi)
        return _ops.stage_clip(x, float(lo), float(hi))

    @staticmethod
    def backward(ctx, grad_y):
        (x,) = ctx.saved_tensors
        return _ops.stage_clip_backward(x, grad_y.contiguous(), ctx.lo, ctx.hi), None, None


class _StageGamma(torch.autograd.Function):
    """exp((1 / gamma) * log(x)) (:209), differentiable in x and gamma."""

    @staticmethod
    def forward(ctx, x, gamma):
        y = _ops.stage_gamma(x, gamma)
        ctx.save_for_backward(x, y, gamma)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        x, y, gamma = ctx.saved_tensors
        gx, gg = _ops.stage_gamma_backward(x, y, grad_y.contiguous(), gamma, ctx.needs_input_grad[0])
        return (gx if ctx.needs_input_grad[0] else None), (gg.reshape(gamma.shape) if ctx.needs_input_grad[1] else None)


def colour_stage_weight(debayer_weight, white_balance, colour_correction):
    """Combined weight of Debayer -> white balance -> colour matrix (:187-191): W[k][i] = sum_c ccm[k][c] wb[c] Wd[c][i]."""
    return torch.einsum('kc,c,cidj->kidj', colour_correction, white_balance.reshape(3), debayer_weight).contiguous()


def luma_stage_weight(luma_taps, rgb2yuv, yuv2rgb):
    """Combined weight of RGB -> YUV, a K x K filter on Y alone, YUV -> RGB (:194-198 / :199-203):
    W[o][i][a][b] = M2[o][0] taps[a][b] M1[0][i] + [a, b centre] (M2[o][1] M1[1][i] + M2[o][2] M1[2][i])."""
    k = luma_taps.shape[-1]
    taps = luma_taps.reshape(k, k)
    w = torch.einsum('o,i,ab->oiab', yuv2rgb[:, 0], rgb2yuv[0, :], taps)
    centre = torch.zeros((k, k), device=taps.device, dtype=taps.dtype)
    centre[k // 2, k // 2] = 1.0
    chroma = yuv2rgb[:, 1:] @ rgb2yuv[1:, :]                      # (3, 3): the U, V planes pass through
    return (w + chroma[:, :, None, None] * centre).contiguous()


def forward_staged(self, raw):
    """``ParametrizedProcessing.forward`` for ``track_stages=True`` (bound as ``_forward_staged``)."""
    if ops._exchange is not None:
        import warnings
        warnings.warn("track_stages=True runs the stage kernels: its parameter gradients are rank-local, the fused "
                      "data-parallel gradient exchange does not apply to this call", RuntimeWarning, stacklevel=3)
    self.stages['demosaic'] = rgb = ops.mosaic(raw, self.black_level, reduce_size=False, out_channels=3,
                                               raw_denominator=float(2 ** self.raw_bits - 1))

    w = colour_stage_weight(self.debayer.weight, self.white_balance, self.colour_correction)
    self.stages['color_correct'] = rgb = _StageConv.apply(rgb, w, True)

    w = luma_stage_weight(self.sharpening_filter.weight, self.M_RGB_2_YUV, self.M_YUV_2_RGB)
    self.stages['sharpening'] = rgb = _StageConv.apply(rgb, w, False)

    w = luma_stage_weight(self.gaussian_blur.weight, self.M_RGB_2_YUV, self.M_YUV_2_RGB)
    self.stages['gaussian'] = rgb = _StageConv.apply(rgb, w, True)

    self.stages['clipped'] = rgb = _StageClip.apply(rgb, 1e-5, 1.0)
    self.stages['gamma_correct'] = rgb = _StageGamma.apply(rgb, self.gamma_correct)

    if self.additive_layer is not None:
        self.stages['noise'] = rgb = rgb + self.additive_layer
    if self.batch_norm is not None:
        rgb = self.batch_norm(rgb)

    if raw.requires_grad:
        for stage in self.stages.values():
            stage.retain_grad()
    self.buffer['processed_rgb'] = rgb
    return rgb
