#!/usr/bin/env python
"""bench.py -- ISP forward+backward throughput (Mpixel/s) on N B200s, with roofline, e2e and CPU baseline.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
A "step" is one fused forward + backward (input gradient + all 132 parameter gradients) of the parametrized ISP
over one synthetic RGGB batch (BASELINE.json configs[1]: batch 64, 256x256, fp32).  `value` times the kernels
through the C ABI with inputs resident in HBM; `e2e` times the same step through the nn.Module / autograd public
API with the raw batch starting in pinned HOST memory and the parameter gradients read back to the host.
`--impl reference` times the reference's CPU implementation of the path (oracle port: the same ATen CPU ops the
reference module runs) on the host cores.
"""
import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "isp_fwd_bwd_mpixel_per_s"
UNIT = "Mpixel/s"
BYTES_FWD, BYTES_BWD = 16, 20          # algorithmic bytes per pixel: raw 4 + rgb 12 | raw 4 + g 12 + graw 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--sets", type=int, default=8, help="rotating buffer sets (working set must exceed the L2)")
    ap.add_argument("--preset", default="drone")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_traffic(kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of `kernel` from the committed ncu capture
    of this same command (profiles/traffic.json, written from `ncu --set full`), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)["kernels"][kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port on host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_fwd_bwd_seconds(batch, size, preset, repeats):
    from oracle import isp_oracle
    from raw2logit_b200 import synthetic as syn
    raw = syn.smooth_scene(batch, size, size, preset, seed=1234)
    state = isp_oracle.default_state(syn.CAMERA_PRESETS[preset])
    g = torch.full((batch, 3, size, size), 1.0 / (batch * 3 * size * size))
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        isp_oracle.forward_backward(raw, state, grad_out=g, raw_grad=True)
        times.append(time.perf_counter() - t0)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    per_image = 0.025                                   # s, fwd+bwd of one 256^2 image on ~8 cores
    budget = 120.0
    scale = (args.size / 256.0) ** 2
    b = int(max(1, min(args.batch, budget / max(1, args.steps + args.warmup) / (per_image * scale))))
    cpu_fwd_bwd_seconds(b, args.size, args.preset, args.warmup if args.warmup < 3 else 3)
    times = cpu_fwd_bwd_seconds(b, args.size, args.preset, args.steps)
    dt = sum(times) / len(times)
    value = b * args.size * args.size / dt / 1e6
    sample = f"{b}x{args.size}x{args.size} fp32 per step, fwd+bwd incl. raw grad, oracle port (torch CPU ops)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sample_batch=b),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, sample_batch=None):
    cfg = {"workload": f"parametrized ISP fwd+bwd (raw grad + 132 param grads), RGGB {args.size}x{args.size}, "
                       f"batch {args.batch} per GPU, fp32, preset {args.preset}, no BN tail",
           "batch_per_gpu": args.batch, "height": args.size, "width": args.size,
           "l2_policy": f"{args.sets} rotating buffer sets (working set > 126 MB L2)",
           "parallelism": f"dp{args.gpus}"}
    if sample_batch is not None:
        cfg["sample_batch"] = sample_batch
    return cfg


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "sw_power_cap": 0x4, "hw_power_brake_slowdown": 0x80}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from oracle import isp_oracle          # cpu_baseline leg only
    from raw2logit_b200 import _lib, ops, synthetic as syn
    from processing.pipeline_torch import ParametrizedProcessing

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    B, H, W, S = args.batch, args.size, args.size, args.sets
    pix = B * H * W
    cam = syn.CAMERA_PRESETS[args.preset]
    mod = ParametrizedProcessing(cam, batch_norm_output=False).to(dev)
    base = syn.smooth_scene(B, H, W, args.preset, seed=1234 + rank)
    vp0 = ctypes.c_void_p
    host_raw = [torch.roll(base, shifts=2 * s, dims=0).contiguous().pin_memory() for s in range(S)]
    raws = [h.to(dev) for h in host_raw]
    gouts = [torch.full((B, 3, H, W), 1.0 / (3 * pix), device=dev) * (1.0 + 0.01 * s) for s in range(S)]
    outs = [torch.empty(B, 3, H, W, device=dev) for _ in range(S)]
    graws = [torch.empty(B, H, W, device=dev) for _ in range(S)]
    # Y0 / Y1 planes the forward keeps for the backward (8 B/px, r2l_isp.h: saved_luma)
    lumas = [torch.empty(lib.r2l_isp_saved_luma_floats(B, H, W), device=dev) for _ in range(S)]
    assert lib.r2l_isp_luma_supported(vp0(raws[0].data_ptr()), _lib.F32, B, H, W, vp0(outs[0].data_ptr()), None) == 1
    gpar = torch.empty(_lib.NUM_PARAM_GRADS, device=dev)
    nws = lib.r2l_isp_workspace_bytes(B, H, W)
    wsb = torch.empty(nws // 4, device=dev)
    ptensors = [mod.black_level, mod.white_balance, mod.colour_correction, mod.gamma_correct, mod.debayer.weight,
                mod.sharpening_filter.weight, mod.gaussian_blur.weight, mod.M_RGB_2_YUV, mod.M_YUV_2_RGB]
    params = _lib.IspParams(*[t.data_ptr() for t in ptensors])
    vp = ctypes.c_void_p
    stream = torch.cuda.current_stream()
    sp = vp(stream.cuda_stream)

    def step_kernels(i):
        s = i % S
        rc = lib.r2l_isp_forward(vp(raws[s].data_ptr()), _lib.F32, 65535.0, B, H, W, ctypes.byref(params), None,
                                 vp(outs[s].data_ptr()), vp(lumas[s].data_ptr()), sp)
        return rc, s

    # N > 1: the 132-float gradient all-reduce is fused into the backward kernel (its last CTA exchanges the gradients over
    # NVLink peer memory, r2l_isp_backward_dp); NCCL is the fallback when symmetric memory cannot be set up
    xch = None
    if world > 1 and os.environ.get("R2L_BENCH_NCCL", "0") != "1":
        try:
            from raw2logit_b200 import parallel
            xch = parallel.PeerExchange()
        except Exception as e:                                      # noqa: BLE001 - report and fall back
            if rank == 0:
                print(f"[bench] fused exchange unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr)
            xch = None
        # all ranks must take the same path (a rank that waits in the fused exchange for one that went to NCCL never returns)
        ok = torch.tensor([1 if xch is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            xch = None

    def step_backward(s):
        if xch is not None:
            d = xch.next(average=False)
            return lib.r2l_isp_backward_dp(vp(raws[s].data_ptr()), _lib.F32, 65535.0, B, H, W, ctypes.byref(params),
                                           vp(gouts[s].data_ptr()), None, None, vp(outs[s].data_ptr()),
                                           vp(lumas[s].data_ptr()), vp(graws[s].data_ptr()),
                                           vp(gpar.data_ptr()), vp(wsb.data_ptr()), nws, ctypes.byref(d), sp)
        return lib.r2l_isp_backward(vp(raws[s].data_ptr()), _lib.F32, 65535.0, B, H, W, ctypes.byref(params),
                                    vp(gouts[s].data_ptr()), None, None, vp(outs[s].data_ptr()),
                                    vp(lumas[s].data_ptr()), vp(graws[s].data_ptr()),
                                    vp(gpar.data_ptr()), vp(wsb.data_ptr()), nws, sp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) -----------------------------------------------------------------
    for i in range(args.warmup):
        rc, s = step_kernels(i)
        _lib.check(rc, "forward")
        _lib.check(step_backward(s), "backward")
        if world > 1 and xch is None:
            dist.all_reduce(gpar)
    K = args.steps
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_begin.record()
    for i in range(K):
        rc, s = step_kernels(i)
        rc2 = step_backward(s)
        if world > 1 and xch is None and os.environ.get("R2L_BENCH_NO_EXCHANGE", "0") != "1":
            dist.all_reduce(gpar)
    t_end.record()
    barrier()
    sampler.stop_flag = True
    _lib.check(rc, "forward")
    _lib.check(rc2, "backward")
    total_ms = t_begin.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = t.item()
    ms_per_step = total_ms / K
    value = world * pix / (ms_per_step * 1e-3) / 1e6

    # per-kernel launch durations for the roofline: same loop, CUDA events around each launch (kept out of the
    # loop above so the event records do not perturb the headline number)
    KR = max(3, min(K, 100))
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(KR)]
    barrier()
    for i in range(KR):
        ev[i][0].record()
        rc, s = step_kernels(i)
        ev[i][1].record()
        rc2 = step_backward(s)
        ev[i][2].record()
    barrier()
    fwd_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in ev)
    bwd_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in ev)

    # ---- end-to-end through the public module API, host buffers ------------------------------------------
    # Every step copies its raw batch from pinned host memory (side stream, double-buffered like a DataLoader
    # prefetcher, so the copy of step i+1 overlaps the kernels of step i), runs ParametrizedProcessing.forward and the
    # autograd backward, and reads the 132 parameter gradients back to the host.
    host_grads = torch.empty(132, dtype=torch.float32).pin_memory()
    plist = [p for p in mod.parameters()]
    copy_stream = torch.cuda.Stream()

    def e2e_run(host_batches, steps):
        bufs = [torch.empty_like(host_batches[0], device=dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        main = torch.cuda.current_stream()
        for ev_ in consumed:
            ev_.record(main)

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i % 2])
                bufs[i % 2].copy_(host_batches[i % len(host_batches)], non_blocking=True)
                ready[i % 2].record(copy_stream)

        prefetch(0)
        for i in range(steps):
            if i + 1 < steps:
                prefetch(i + 1)
            main.wait_event(ready[i % 2])
            x = bufs[i % 2]
            if x.dtype == torch.float32:
                x = x.detach().requires_grad_(True)
            out = mod(x)
            out.backward(gouts[i % S])
            consumed[i % 2].record(main)
            flat = torch.cat([p.grad.reshape(-1) for p in plist])
            if world > 1 and not fused_e2e:
                dist.all_reduce(flat)
            host_grads.copy_(flat, non_blocking=True)
            for p in plist:
                p.grad = None

    def e2e_measure(host_batches):
        steps = max(3, min(K, 200))
        e2e_run(host_batches, max(3, min(args.warmup, 10)))
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_run(host_batches, steps)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return world * pix / (ms / steps * 1e-3) / 1e6, steps

    # N > 1: the module path exchanges the ISP gradients inside its backward kernel too (parallel.enable_fused_...)
    fused_e2e = False
    if xch is not None:
        from raw2logit_b200 import parallel
        parallel.enable_fused_gradient_exchange(average=False)
        fused_e2e = True
    e2e_value, e2e_steps = e2e_measure(host_raw)
    # same step fed with the sensor's uint16 words (2 B/px over PCIe; the divide by 2^16-1 happens in the kernel,
    # dataset.py:87); parameter gradients only -- an integer input has no gradient
    host_u16 = [syn.to_uint16(h).pin_memory() for h in host_raw]
    e2e_u16_value, _ = e2e_measure(host_u16)
    if fused_e2e:
        parallel.disable_fused_gradient_exchange()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    default_cfg = (B, H, W) == (64, 256, 256)          # the configuration the committed ncu capture was taken on
    bwd_gbs = BYTES_BWD * pix / (bwd_ms * 1e-3) / 1e9
    fwd_gbs = BYTES_FWD * pix / (fwd_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args), gradient_exchange=(
            "none (1 GPU)" if world == 1 else
            "fused into the backward kernel: one-shot all-reduce of the 132 gradients over NVLink peer memory "
            "(r2l_isp_backward_dp)" if xch is not None else "NCCL all-reduce of the 132 gradients after the backward")),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pix * 4, "d2h_bytes_per_step": 132 * 4,
                "steps": e2e_steps, "api": "ParametrizedProcessing.forward + autograd backward; fp32 raw batch copied "
                                           "from pinned host memory every step on a prefetch stream (double-buffered)"
                                           + ("; gradients exchanged inside the backward kernel "
                                              "(parallel.enable_fused_gradient_exchange)" if fused_e2e else
                                              ("; NCCL all-reduce of the gradients" if world > 1 else ""))},
        "e2e_uint16": {"value": e2e_u16_value, "unit": UNIT, "h2d_bytes_per_step": pix * 2, "d2h_bytes_per_step": 132 * 4,
                       "note": "same step with uint16 raw words over PCIe (normalised in the kernel); parameter "
                               "gradients only"},
        "gpu_launches": 2 * K,
        "roofline": {"bound": "hbm", "kernel": "isp_backward5_kernel", "achieved": bwd_gbs, "peak": peak, "unit": "GB/s",
                     "frac": bwd_gbs / peak, "traffic": measured_traffic("isp_backward5_kernel") if default_cfg else None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_BWD * pix, "avg_launch_ms": bwd_ms,
                     "note": "traffic > algorithmic bytes by design: the backward reads the saved forward output "
                             "(12 B/px) and luma planes (8 B/px) instead of recomputing them (DESIGN.md 4.2)"},
        "roofline_forward": {"bound": "hbm", "kernel": "isp_forward3_kernel", "achieved": fwd_gbs, "peak": peak,
                             "unit": "GB/s", "frac": fwd_gbs / peak, "algorithmic_bytes_per_launch": BYTES_FWD * pix,
                             "traffic": measured_traffic("isp_forward3_kernel") if default_cfg else None,
                             "avg_launch_ms": fwd_ms},
        "roofline_step_frac": (BYTES_FWD + BYTES_BWD) * pix / (ms_per_step * 1e-3) / 1e9 / peak,
        "clocks": sampler.summary(),
    }
    if not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        cb = min(B, 16)
        cpu_fwd_bwd_seconds(cb, H, args.preset, 1)
        ts = cpu_fwd_bwd_seconds(cb, H, args.preset, 3)
        line["cpu_baseline"] = {"value": cb * H * W / min(ts) / 1e6, "unit": UNIT, "cores": torch.get_num_threads(),
                                "kind": "port",
                                "sample": f"{cb}x{H}x{W} fp32 fwd+bwd incl. raw grad, best of 3, oracle port "
                                          "(same ATen CPU ops as the reference module)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
