#!/usr/bin/env python
"""CUDA-graph capture of the module-API step (forward + autograd backward): the result must equal the eager step bit for
bit, and one replay costs one graph launch of host time.  usage (GPU box): python scripts/graph_check.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from raw2logit_b200 import synthetic as syn  # noqa: E402
from processing.pipeline_torch import ParametrizedProcessing  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    ok = True
    for bn, u16 in ((False, False), (True, False), (False, True), (True, True)):
        torch.manual_seed(0)
        mod = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], batch_norm_output=bn).to(dev)
        B, H, W = 64, 256, 256
        raw = syn.smooth_scene(B, H, W, "drone", seed=1)
        x_static = (syn.to_uint16(raw) if u16 else raw).to(dev)
        g = torch.full((B, 3, H, W), 1e-6, device=dev)
        plist = list(mod.parameters())

        def step():
            out = mod(x_static)
            out.backward(g)
            return out

        # eager result
        step()
        eager = torch.cat([p.grad.flatten() for p in plist]).clone()
        for p in plist:
            p.grad = None
        # capture
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
                for p in plist:
                    p.grad = None
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = step()
            flat = torch.cat([p.grad.flatten() for p in plist])
        graph.replay()
        torch.cuda.synchronize()
        same = torch.equal(flat, eager)
        ok = ok and same
        n = 300
        t0 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            graph.replay()
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"bn={bn} u16={u16}: graph == eager {same}; host {1e6 * (t1 - t0) / n:.1f} us / replay, "
              f"GPU {1e3 * e0.elapsed_time(e1) / n:.1f} us / step", flush=True)
        # eager host time for comparison
        for _ in range(20):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            step()
            for p in plist:
                p.grad = None
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"             eager: host {1e6 * (t1 - t0) / n:.1f} us / step, GPU-side {1e3 * e0.elapsed_time(e1) / n:.1f} us / step",
              flush=True)
    print("GRAPH OK" if ok else "GRAPH MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
