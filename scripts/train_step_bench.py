#!/usr/bin/env python
"""BASELINE.json configs[2] / configs[3]: one training step of the parametrized ISP + a stock-PyTorch task model,
data-parallel over N GPUs (torchrun), ISP and task model timed separately (north star).

  configs[2]  Microscopy classification: ISP + ResNet18 (train.py:86), CrossEntropy, Adam lr 1e-5, batch 32/GPU, 256^2
  configs[3]  Drone segmentation: ISP + U-Net stand-in (smp is not installable), Dice loss, tiles 256^2 or larger

Usage: python -m torch.distributed.run --nproc-per-node N scripts/train_step_bench.py --task {microscopy,drone}
Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from raw2logit_b200 import model as r2l_model, parallel, synthetic as syn  # noqa: E402
from processing.pipeline_torch import ParametrizedProcessing  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="microscopy", choices=["microscopy", "drone"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    preset = "microscopy" if args.task == "microscopy" else "drone"
    proc = ParametrizedProcessing(syn.CAMERA_PRESETS[preset], track_stages=False, batch_norm_output=True)   # train.py:195
    if args.task == "microscopy":
        clf, loss = r2l_model.resnet_model("resnet18", pretrained=False, fc_out_features=16), torch.nn.CrossEntropyLoss()
    else:
        clf, loss = r2l_model.SmallUNet(), r2l_model.dice_loss
    lit = r2l_model.LitModel(clf, loss, lr=1e-5, processor=proc, is_segmentation_task=args.task == "drone").to(dev).train()
    opt = lit.configure_optimizers()
    B, H = args.batch, args.size
    scale = 0.25 if preset == "microscopy" else 1.0
    raw = (syn.smooth_scene(B, H, H, preset, seed=100 + rank) * scale).to(dev)
    if args.task == "microscopy":
        y = torch.randint(0, 16, (B,), device=dev)
    else:
        y = (torch.rand(B, 1, H, H, device=dev) > 0.5).float()

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    t_isp_f = t_isp_b = t_total = 0.0
    for it in range(args.warmup + args.steps):
        e = [ev() for _ in range(6)]
        for p in lit.parameters():
            p.grad = None
        # keep the GPU busy for ~0.3 ms so that the host runs ahead: the events then bracket GPU time, not the
        # dispatch latency of an idle queue (at batch 32 the ISP kernels are shorter than their launch path)
        torch.cuda._sleep(600_000)
        e[0].record()
        rgb = lit.processor(raw)
        e[1].record()
        logits = lit.classifier(rgb)
        l = lit.loss_fn(logits, y)
        e[2].record()
        # one backward: the hook fires when d loss / d rgb is complete, i.e. right before the ISP's backward node runs
        # (it is the last node of the graph: the processor is the first module of LitModel.forward, model.py:78)
        rgb.register_hook(lambda g, ev3=e[3]: ev3.record())
        l.backward()
        e[4].record()
        parallel.allreduce_gradients(lit.parameters(), world=world)
        opt.step()
        e[5].record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            t_isp_f += e[0].elapsed_time(e[1])
            t_isp_b += e[3].elapsed_time(e[4])
            t_total += e[0].elapsed_time(e[5])
    n = args.steps
    res = torch.tensor([t_isp_f / n, t_isp_b / n, t_total / n], device=dev)
    if world > 1:
        dist.all_reduce(res, op=dist.ReduceOp.MAX)
    if rank == 0:
        f, b, tot = res.tolist()
        pix = B * H * H * world
        print(json.dumps({"task": args.task, "n_gpus": world, "batch_per_gpu": B, "size": H,
                          "isp_forward_ms": round(f, 4), "isp_backward_ms": round(b, 4), "step_ms": round(tot, 3),
                          "isp_share_of_step": round((f + b) / tot, 4),
                          "isp_mpixel_per_s": round(pix / ((f + b) * 1e-3) / 1e6, 1),
                          "step_mpixel_per_s": round(pix / (tot * 1e-3) / 1e6, 1),
                          "note": "ISP = fused forward + BN-train tail and fused backward (no raw grad); task model = "
                                  "stock PyTorch; ISP times are GPU time between events (the queue is kept busy), one backward pass"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
