"""Kernel-backed drop-in for the reference's ``processing/pipeline_torch.py``.

Same class names, constructor arguments, parameter / buffer names (``state_dict`` keys) and forward signatures as
the reference (pipeline_torch.py:43-283), so ``model.py`` / ``train.py`` wiring (``processing_mode``,
``freeze_processor``, ``track_processing_gradients``, ``adv_parameters`` substring matching, deepcopy, pickling,
``load_state_dict(strict=True)`` of reference checkpoints) keeps working -- but ``forward`` is one fused sm_100a
kernel (and ``backward`` a second one) instead of ~50 ATen launches and a 79-node autograd graph.

CUDA only.  There is no CPU path: CPU inputs raise.  The sub-modules ``debayer``, ``sharpening_filter`` and
``gaussian_blur`` are kept as ``nn.Conv2d`` *parameter holders* (their ``forward`` is never called).
"""
import torch
import torch.nn as nn

from . import ops

# Calibration constants of the reference (pipeline_torch.py:13-40); numeric values are the contract.
K_G = torch.tensor([[0., 1., 0.], [1., 4., 1.], [0., 1., 0.]]) / 4
K_RB = torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]]) / 4
M_RGB_2_YUV = torch.tensor([[0.299, 0.587, 0.114],
                            [-0.14714119, -0.28886916, 0.43601035],
                            [0.61497538, -0.51496512, -0.10001026]])
M_YUV_2_RGB = torch.tensor([[1.0000000000e+00, -4.1827794561e-09, 1.1398830414e+00],
                            [1.0000000000e+00, -3.9464232326e-01, -5.8062183857e-01],
                            [1.0000000000e+00, 2.0320618153e+00, -1.2232658220e-09]])
_g5 = [6.9625e-08, 2.8089e-05, 2.0755e-04, 1.1332e-02, 8.3731e-02, 6.1869e-01]
K_BLUR = torch.tensor([[_g5[0], _g5[1], _g5[2], _g5[1], _g5[0]],
                       [_g5[1], _g5[3], _g5[4], _g5[3], _g5[1]],
                       [_g5[2], _g5[4], _g5[5], _g5[4], _g5[2]],
                       [_g5[1], _g5[3], _g5[4], _g5[3], _g5[1]],
                       [_g5[0], _g5[1], _g5[2], _g5[1], _g5[0]]])
K_SHARP = torch.tensor([[0., -1., 0.], [-1., 5., -1.], [0., -1., 0.]])
DEFAULT_CAMERA_PARAMS = ([0., 0., 0., 0.], [1., 1., 1.], [1., 0., 0., 0., 1., 0., 0., 0., 1.])

_REF_MODULE = "processing.pipeline_torch"   # pickles of reference models name this module path


def _require_cuda(raw):
    if not isinstance(raw, torch.Tensor) or not raw.is_cuda:
        raise RuntimeError("raw2logit_b200 is CUDA-only (sm_100a kernels, no CPU fallback): move the module and "
                           f"the raw batch to a CUDA device (got {getattr(raw, 'device', type(raw))})")


def raw2rgb(raw, black_level=None, reduce_size=True, out_channels=3, raw_denominator=65535.0):
    """CFA split of an RGGB mosaic (reference ``raw2rgb``, pipeline_torch.py:240-283).

    raw (B,H,W) -> (B,C,H//2,W//2) packed (``reduce_size=True``; greens averaged for C=3) or (B,C,H,W) zero-filled.
    Output is always float32.  ``raw_denominator`` only matters for uint16 input.
    """
    assert out_channels in [3, 4]
    _require_cuda(raw)
    bl = None
    if black_level is not None:
        bl = black_level if isinstance(black_level, torch.Tensor) else torch.as_tensor(black_level, dtype=torch.float32)
        bl = bl.to(device=raw.device, dtype=torch.float32).reshape(4)
    # differentiable in raw and black_level, like the reference's indexing arithmetic (:256-259, :273-277)
    return ops.mosaic(raw, bl, reduce_size, out_channels, raw_denominator)


class RawToRGB(nn.Module):
    """'none' processing mode: CFA split only (reference ``RawToRGB``, pipeline_torch.py:43-80)."""

    def __init__(self, reduce_size=True, out_channels=3, track_stages=False, normalize_mosaic=None):
        super().__init__()
        self.stages = None
        self.buffer = None
        self.reduce_size = reduce_size
        self.out_channels = out_channels
        self.track_stages = track_stages
        self.normalize_mosaic = normalize_mosaic

    def forward(self, raw):
        self.stages = {}
        self.buffer = {}
        rgb = raw2rgb(raw, reduce_size=self.reduce_size, out_channels=self.out_channels)
        self.stages['demosaic'] = rgb
        if self.normalize_mosaic:
            rgb = self.normalize_mosaic(rgb)
        if self.track_stages and raw.requires_grad:
            for stage in self.stages.values():
                stage.retain_grad()
        self.buffer['processed_rgb'] = rgb
        return rgb


class NNProcessing(nn.Module):
    """Out of the hot-path scope: the reference's learned U-Net++ processor (pipeline_torch.py:83-126)
    is a CNN from ``segmentation_models_pytorch``, not an ISP kernel.  Import-safe stub with the reference's constructor
    signature: constructing it raises."""

    def __init__(self, track_stages=False, normalize_mosaic=None, batch_norm_output=True):
        super().__init__()
        raise NotImplementedError("NNProcessing (U-Net++ from segmentation_models_pytorch) is outside the ISP hot path "
                                  "this package replaces; use the reference's class for processing_mode=neural_network")


def append_additive_layer(processor):
    """Adversarial-mode additive layer (reference pipeline_torch.py:129-131; hard-codes 256x256 like the reference)."""
    ref = processor.black_level
    processor.additive_layer = nn.Parameter(torch.zeros((1, 3, 256, 256), device=ref.device, dtype=ref.dtype))


class Debayer(nn.Conv2d):
    """Parameter holder for the trainable 3->3 3x3 demosaic taps, bilinear at init (reference :228-237)."""

    def __init__(self):
        super().__init__(3, 3, kernel_size=3, padding=1, padding_mode='reflect', bias=False)
        self.weight.data.fill_(0)
        self.weight.data[0, 0] = K_RB.clone()
        self.weight.data[1, 1] = K_G.clone()
        self.weight.data[2, 2] = K_RB.clone()


class ParametrizedProcessing(nn.Module):
    """Differentiable raw -> RGB pipeline, fused on the GPU (reference ``ParametrizedProcessing``, :134-225).

    Args:
        camera_parameters (tuple(list), optional): (black_level[4], white_balance[3], colour_matrix[9])
        track_stages (bool, optional): keep every intermediate stage (and its gradient) -- staged kernels
        batch_norm_output (bool, optional): BatchNorm2d(3, affine=False) at the end
    """

    def __init__(self, camera_parameters=None, track_stages=False, batch_norm_output=True):
        super().__init__()
        self.stages = None
        self.buffer = None
        self.track_stages = track_stages
        if camera_parameters is None:
            camera_parameters = DEFAULT_CAMERA_PARAMS
        black_level, white_balance, colour_matrix = camera_parameters

        self.black_level = nn.Parameter(torch.as_tensor(black_level, dtype=torch.float32).clone())
        self.white_balance = nn.Parameter(torch.as_tensor(white_balance, dtype=torch.float32).reshape(1, 3).clone())
        self.colour_correction = nn.Parameter(torch.as_tensor(colour_matrix, dtype=torch.float32).reshape(3, 3).clone())
        self.gamma_correct = nn.Parameter(torch.tensor([2.2]))
        self.debayer = Debayer()
        self.sharpening_filter = nn.Conv2d(1, 1, kernel_size=3, padding=1, bias=False)
        self.sharpening_filter.weight.data[0][0] = K_SHARP.clone()
        self.gaussian_blur = nn.Conv2d(1, 1, kernel_size=5, padding=2, padding_mode='reflect', bias=False)
        self.gaussian_blur.weight.data[0][0] = K_BLUR.clone()
        self.batch_norm = nn.BatchNorm2d(3, affine=False) if batch_norm_output else None
        self.register_buffer('M_RGB_2_YUV', M_RGB_2_YUV.clone())
        self.register_buffer('M_YUV_2_RGB', M_YUV_2_RGB.clone())
        self.additive_layer = None  # this can be added in later
        self.raw_bits = 16          # uint16 ingest: value = u / (2**raw_bits - 1)   (dataset.py:87)

    def forward(self, raw):
        assert raw.ndim == 3, f"needs dims (B, H, W), got {raw.shape}"
        _require_cuda(raw)
        d = self.__dict__                      # plain attributes: nn.Module.__setattr__ costs ~10 us per assignment
        d['stages'] = {}
        d['buffer'] = {}
        if self.track_stages:
            return self._forward_staged(raw)

        bn = self.batch_norm
        bn_mode, rm, rv, nbt, momentum, eps = 0, None, None, None, 0.1, 1e-5
        if bn is not None:
            eps = bn.eps
            use_batch_stats = bn.training or bn.running_mean is None
            bn_mode = 2 if use_batch_stats else 1
            rm, rv = bn.running_mean, bn.running_var
            if use_batch_stats and bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
                # nn.BatchNorm2d bookkeeping: the counter is advanced by the forward kernel (no extra launch, and a
                # captured step counts its replays); only the cumulative-average mode needs its value on the host
                if bn.momentum is not None and bn.num_batches_tracked.is_cuda:
                    nbt, momentum = bn.num_batches_tracked, bn.momentum
                else:
                    bn.num_batches_tracked.add_(1)
                    momentum = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
            elif use_batch_stats:
                rm, rv = None, None                                  # batch statistics, nothing to update
        additive = self.additive_layer
        if additive is not None and additive.numel() != 3 * raw.shape[1] * raw.shape[2]:
            raise RuntimeError(f"additive_layer {tuple(additive.shape)} does not match a "
                               f"{raw.shape[1]}x{raw.shape[2]} frame")
        rgb = ops.fused_isp(raw, self.black_level, self.white_balance, self.colour_correction, self.gamma_correct,
                            self.debayer.weight, self.sharpening_filter.weight, self.gaussian_blur.weight,
                            self.M_RGB_2_YUV, self.M_YUV_2_RGB, additive=additive, bn_mode=bn_mode,
                            running_mean=rm, running_var=rv, momentum=momentum, eps=eps,
                            raw_denominator=float(2 ** self.raw_bits - 1), num_batches_tracked=nbt)
        self.buffer['processed_rgb'] = rgb
        return rgb


def _forward_staged(self, raw):
    """``track_stages=True``: every intermediate of the chain is materialised as its own autograd node so that
    ``processor.stages[name]`` and ``processor.stages[name].grad`` exist for ``model.track_images``
    (reference model.py:204-300, pipeline_torch.py:183-221) -- one repo kernel per stage, see ``staged.py``."""
    from . import staged
    return staged.forward_staged(self, raw)


ParametrizedProcessing._forward_staged = _forward_staged

for _cls in (RawToRGB, NNProcessing, Debayer, ParametrizedProcessing):
    _cls.__module__ = _REF_MODULE
