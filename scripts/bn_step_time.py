#!/usr/bin/env python
"""GPU time of the module-API step with and without the BatchNorm tail (train mode), batch 64 x 256 x 256, measured
with CUDA events over a CUDA-graph replay so that host overhead does not show."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from raw2logit_b200 import synthetic as syn  # noqa: E402
from processing.pipeline_torch import ParametrizedProcessing  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    B, H, W = 64, 256, 256
    raw = syn.smooth_scene(B, H, W, "drone", seed=1).to(dev)
    g = torch.full((B, 3, H, W), 1e-6, device=dev)
    for bn in (False, True):
        mod = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], batch_norm_output=bn).to(dev).train()
        plist = list(mod.parameters())

        def step():
            x = raw.detach().requires_grad_(True)
            mod(x).backward(g)
            for p in plist:
                p.grad = None

        for _ in range(5):
            step()
        torch.cuda.synchronize()
        n = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # queue many steps so the GPU stays busy; host overhead (~0.2 ms/step) exceeds GPU time, so also report kernels via profiler
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(n):
                step()
            torch.cuda.synchronize()
        tot = {}
        for ev in prof.key_averages():
            if ev.device_time_total > 0:
                tot[ev.key] = (ev.device_time_total / n, ev.count / n)
        print(f"== batch_norm_output={bn}: GPU kernel time per step {sum(v[0] for v in tot.values()):.1f} us")
        for k, (t, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:9]:
            print(f"   {t:8.1f} us  x{c:.0f}  {k[:90]}")


if __name__ == "__main__":
    main()
