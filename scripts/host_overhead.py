#!/usr/bin/env python
"""Host-side cost of one module-API step (forward + autograd backward) when the GPU is not the limiter (tiny batch):
wall time per step without synchronising, and a cProfile of where it goes."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from raw2logit_b200 import synthetic as syn  # noqa: E402
from processing.pipeline_torch import ParametrizedProcessing  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    mod = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], batch_norm_output=False).to(dev)
    raw = syn.smooth_scene(2, 64, 64, "drone", seed=1).to(dev)
    raw16 = syn.to_uint16(raw.cpu()).to(dev)
    g = torch.full((2, 3, 64, 64), 1e-4, device=dev)
    plist = list(mod.parameters())

    def step(x):
        if x.dtype == torch.float32:
            x = x.detach().requires_grad_(True)
        out = mod(x)
        out.backward(g)
        for p in plist:
            p.grad = None

    for name, x in (("float32", raw), ("uint16", raw16)):
        for _ in range(50):
            step(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 500
        for _ in range(n):
            step(x)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"{name}: {1e6 * (t1 - t0) / n:.1f} us of host time per step (forward + backward, module API)")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300):
        step(raw16)
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()
