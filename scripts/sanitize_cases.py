#!/usr/bin/env python
"""A handful of small cases through every kernel family of the library, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/sanitize_cases.py
    compute-sanitizer --tool racecheck python scripts/sanitize_cases.py
Fused forward / backward (multi-tile, partial tiles, odd batch, fp32 + raw gradient, uint16, BatchNorm tail in train mode --
the grid-wide barrier of the fused forward -- and additive layer), staged mode, RawToRGB, SSIM, the dihedral hand-off copy,
the numpy-compatible pipeline.  Prints one line per case with a checksum."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import isp_oracle  # noqa: E402
from processing.pipeline_torch import ParametrizedProcessing, RawToRGB, append_additive_layer  # noqa: E402
from raw2logit_b200 import synthetic as syn  # noqa: E402

STATE = syn.perturbed_state(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))


def fused(shape, need_raw, u16=False, bn=False, additive=False, stages=False):
    raw = syn.smooth_scene(*shape, "drone", seed=31)
    x0 = syn.to_uint16(raw).cuda() if u16 else raw.cuda()
    g = isp_oracle.cotangent((shape[0], 3, shape[1], shape[2]), "ramp").cuda()
    mod = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], batch_norm_output=bn, track_stages=stages)
    mod.load_state_dict(STATE, strict=not bn)
    if additive:
        append_additive_layer(mod)
    mod = mod.cuda().train()
    x = x0.clone().requires_grad_(True) if need_raw else x0
    mod(x).backward(g)
    torch.cuda.synchronize()
    flat = torch.cat([p.grad.flatten() for p in mod.parameters() if p.grad is not None]).cpu()
    print("OK fused", shape, "raw_grad", need_raw, "u16", u16, "bn", bn, "additive", additive, "stages", stages,
          float(flat.abs().sum()), flush=True)


def main():
    for shape in [(2, 64, 64), (3, 96, 200), (5, 256, 256)]:
        fused(shape, True)
        fused(shape, False, u16=True)
        fused(shape, False, bn=True)
    fused((2, 256, 256), True, bn=True, additive=True)
    fused((2, 64, 96), True, stages=True)
    fused((1, 37, 53), True)                                        # generic scalar kernels (W % 4 != 0)
    raw = syn.smooth_scene(2, 64, 96, "drone", seed=3).cuda().requires_grad_(True)
    for rs, ch in ((True, 3), (True, 4), (False, 3)):
        out = RawToRGB(reduce_size=rs, out_channels=ch).cuda()(raw)
        out.sum().backward()
    torch.cuda.synchronize()
    print("OK raw2rgb", float(raw.grad.abs().sum()), flush=True)
    from utils.ssim import SSIM
    a = torch.rand(2, 3, 70, 90, device="cuda", requires_grad=True)
    b = torch.rand(2, 3, 70, 90, device="cuda")
    v = SSIM()(a, b)
    v.backward()
    torch.cuda.synchronize()
    print("OK ssim", float(v), float(a.grad.abs().sum()), flush=True)
    import utils.augmentation as aug
    t = aug.ComposeState(aug.augmentation_weak.transforms, memory_format=torch.channels_last, dtype=torch.bfloat16)
    y = t(torch.rand(2, 3, 64, 64, device="cuda"))
    torch.cuda.synchronize()
    print("OK handoff", tuple(y.shape), y.dtype, float(y.float().abs().sum()), flush=True)
    from processing.pipeline_numpy import processing
    cam = syn.CAMERA_PRESETS["drone"]
    out = processing(syn.smooth_scene(2, 64, 96, "drone", seed=5).cuda(), cam[0], cam[1], cam[2], sharpening="sharpening_filter",
                     denoising="gaussian_denoising")
    torch.cuda.synchronize()
    print("OK numpy-mode", tuple(out.shape), float(out.abs().sum()), flush=True)


if __name__ == "__main__":
    main()
