#!/usr/bin/env python
"""BASELINE.json configs[4]: ISP throughput sweep -- static (frozen, forward only) vs parametrized (forward + backward),
256^2 .. 4096^2 synthetic Bayer frames, on N B200s (one process per GPU under torchrun, the frames sharded: every rank
runs the same per-GPU batch, weak scaling, no collective on the data path).  Prints one JSON line per point: aggregate
Mpixel/s (max time over ranks), achieved GB/s per GPU from the algorithmic bytes and the fraction of the
measured HBM peak.  Working sets are kept above the L2 (rotating buffer sets).  Kernel-level timing through the C ABI
with CUDA events.
usage: python scripts/sweep.py   |   python -m torch.distributed.run --nproc-per-node N scripts/sweep.py"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from raw2logit_b200 import _lib, synthetic as syn  # noqa: E402
from processing.pipeline_torch import ParametrizedProcessing  # noqa: E402


def main():
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    mod = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], batch_norm_output=False).to(dev)
    pt = [mod.black_level, mod.white_balance, mod.colour_correction, mod.gamma_correct, mod.debayer.weight,
          mod.sharpening_filter.weight, mod.gaussian_blur.weight, mod.M_RGB_2_YUV, mod.M_YUV_2_RGB]
    params = _lib.IspParams(*[t.data_ptr() for t in pt])
    vp = ctypes.c_void_p
    sp = vp(torch.cuda.current_stream().cuda_stream)
    gpar = torch.empty(132, device=dev)
    for size in (256, 512, 1024, 2048, 4096):
        B = max(2, (64 * 256 * 256) // (size * size) * 4)          # 16.8 Mpx per buffer set (> L2 with outputs)
        S = 3
        pix = B * size * size
        base = syn.smooth_scene(min(B, 4), size, size, "drone", seed=7).to(dev)
        raws = [base.repeat((B + base.shape[0] - 1) // base.shape[0], 1, 1)[:B].contiguous() for _ in range(S)]
        outs = [torch.empty(B, 3, size, size, device=dev) for _ in range(S)]
        gouts = [torch.full((B, 3, size, size), 1.0 / (3 * pix), device=dev) for _ in range(S)]
        graws = [torch.empty(B, size, size, device=dev) for _ in range(S)]
        lumas = [torch.empty(lib.r2l_isp_saved_luma_floats(B, size, size), device=dev) for _ in range(S)]
        nws = lib.r2l_isp_workspace_bytes(B, size, size)
        ws = torch.empty(nws // 4, device=dev)

        def fwd(s, save_luma=True):            # the static (frozen) processor has no backward: no luma planes to keep
            return lib.r2l_isp_forward(vp(raws[s].data_ptr()), _lib.F32, 65535.0, B, size, size, ctypes.byref(params),
                                       None, vp(outs[s].data_ptr()), vp(lumas[s].data_ptr()) if save_luma else None, sp)

        def bwd(s, need_raw):
            return lib.r2l_isp_backward(vp(raws[s].data_ptr()), _lib.F32, 65535.0, B, size, size, ctypes.byref(params),
                                        vp(gouts[s].data_ptr()), None, None, vp(outs[s].data_ptr()),
                                        vp(lumas[s].data_ptr()), vp(graws[s].data_ptr()) if need_raw else None, vp(gpar.data_ptr()),
                                        vp(ws.data_ptr()), nws, sp)

        def time_it(fn, n=12, warm=3):
            for i in range(warm):
                _lib.check(fn(i % S), "kernel")
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                fn(i % S)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = t.item()
            return ms

        t_fs = time_it(lambda s: fwd(s, False))
        t_f = time_it(fwd)
        t_b = time_it(lambda s: bwd(s, True))
        t_bn = time_it(lambda s: bwd(s, False))
        for mode, ms, bpp in (("static: forward only", t_fs, 16), ("parametrized: forward + backward (raw grad)", t_f + t_b, 36),
                              ("parametrized: forward + backward (no raw grad)", t_f + t_bn, 32)):
            gbs = bpp * pix / (ms * 1e-3) / 1e9
            if rank == 0:
                print(json.dumps({"n_gpus": world, "size": size, "batch_per_gpu": B, "mode": mode, "ms": round(ms, 4),
                                  "mpixel_per_s": round(world * pix / (ms * 1e-3) / 1e6, 1), "algorithmic_bytes_per_px": bpp,
                                  "achieved_gbs_per_gpu": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 4)}),
                      flush=True)
        del raws, outs, gouts, graws
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
