#!/usr/bin/env python
"""The UNMODIFIED reference ParametrizedProcessing (processing/pipeline_torch.py, loaded by oracle/ref_loader.py from
/root/reference or the baseline/_ref staging) in stock eager PyTorch ON THE SAME B200, TF32 off and on: the honest
"beat this on the same box" bar next to this repo's fused kernels.  Timing only; prints JSON lines.
usage (GPU box): python scripts/ref_on_gpu.py [--batch 64 --size 256]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import ref_loader  # noqa: E402
from raw2logit_b200 import synthetic as syn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    ref = ref_loader.load_reference()
    cam = syn.CAMERA_PRESETS["drone"]
    B, H = args.batch, args.size
    pix = B * H * H
    sets = 4
    raws = [syn.smooth_scene(B, H, H, "drone", seed=1234 + s).to(dev) for s in range(sets)]
    g = torch.full((B, 3, H, H), 1.0 / (3 * pix), device=dev)
    for bn in (False, True):
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            mod = ref.ParametrizedProcessing(cam, batch_norm_output=bn).to(dev).train()

            def step(i, backward=True):
                x = raws[i % sets].detach().requires_grad_(backward)
                for p in mod.parameters():
                    p.grad = None
                out = mod(x)
                if backward:
                    out.backward(g)

            res = {}
            for name, bw in (("forward", False), ("forward_backward", True)):
                ctx = torch.no_grad() if not bw else torch.enable_grad()
                with ctx:
                    for i in range(5):
                        step(i, bw)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for i in range(args.steps):
                        step(i, bw)
                    e1.record()
                    torch.cuda.synchronize()
                res[name + "_ms"] = e0.elapsed_time(e1) / args.steps
            print(json.dumps({"impl": "reference module, stock eager PyTorch on the B200", "tf32": tf32, "batch_norm_output": bn,
                              "batch": B, "size": H, **{k: round(v, 4) for k, v in res.items()},
                              "fwd_bwd_mpixel_per_s": round(pix / (res["forward_backward_ms"] * 1e-3) / 1e6, 1),
                              "fwd_mpixel_per_s": round(pix / (res["forward_ms"] * 1e-3) / 1e6, 1),
                              "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0)}), flush=True)


if __name__ == "__main__":
    main()
