#!/usr/bin/env python
"""bench.py -- ISP forward+backward throughput (Mpixel/s) on N B200s, with roofline, e2e and CPU baselines.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
A "step" is one fused forward + backward (input gradient + all 132 parameter gradients) of the parametrized ISP
over one synthetic RGGB batch (BASELINE.json configs[1]: batch 64, 256x256, fp32).

  value            the two kernels through the C ABI, inputs resident in HBM, 8 rotating buffer sets (> L2)
  e2e              the training step through the nn.Module / autograd API from HOST buffers: the sensor's uint16 words
                   (dataset.py:87 divides them by 2^16-1; here the kernel does) copied from pinned memory every step,
                   ParametrizedProcessing.forward + backward captured in a CUDA graph (raw2logit_b200.graphs), the 132
                   gradients read back to the host.  e2e_fp32 is the same with an fp32 batch and the raw gradient,
                   e2e_eager the uint16 step without graph capture.
  module_resident  the value's step through the module API with device-resident input (eager and graphed)
  train_case       forward + backward without the raw gradient (what training asks for, model.py:77-83): 32 B/px
  bn_tail          the step with the reference's default BatchNorm2d tail (train.py:196), train mode
  backward_cold_l2 the backward with the L2 flushed between forward and backward (a task model runs in between)
  cpu_baseline     the UNMODIFIED reference module (baseline/_ref staging of processing/pipeline_torch.py, loaded by
                   oracle/ref_loader.py) on the host cores, bounded sample; cpu_baseline_numpy: the numpy chain
                   (pipeline_numpy.py restated, oracle/numpy_oracle.py), one process and a pool of all cores
`--impl reference` times the reference's CPU implementation of the path on the host cores (same metric / config).
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
BYTES_BWD_NORAW = 16                   # training case: no raw gradient


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
    ap.add_argument("--kernels-only", action="store_true", help="value + roofline only (A/B timing of kernel builds)")
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


def workload_config(args):
    return {"workload": f"parametrized ISP fwd+bwd (raw grad + 132 param grads), RGGB {args.size}x{args.size}, "
                        f"batch {args.batch} per GPU, fp32, preset {args.preset}, no BN tail",
            "batch_per_gpu": args.batch, "height": args.size, "width": args.size,
            "l2_policy": f"{args.sets} rotating buffer sets (working set > 126 MB L2)",
            "parallelism": f"dp{args.gpus}"}


# ------------------------------------------------------------------------------------------------------------
# reference arm / CPU baselines (host cores)
# ------------------------------------------------------------------------------------------------------------
class CpuReference:
    """fwd+bwd (raw gradient + parameter gradients) of the path on the host: the unmodified reference module when its
    file is present (/root/reference here, baseline/_ref on the GPU box), else the oracle port (same ATen CPU ops)."""

    def __init__(self, preset):
        from oracle import isp_oracle, ref_loader
        from raw2logit_b200 import synthetic as syn
        self.syn, self.oracle, self.preset = syn, isp_oracle, preset
        self.cam = syn.CAMERA_PRESETS[preset]
        self.mod = None
        self.kind = "port"
        if ref_loader.available():
            try:
                ref = ref_loader.load_reference()
                self.mod = ref.ParametrizedProcessing(self.cam, batch_norm_output=False)
                self.kind = "reference"
                self.where = ref_loader.REF_ROOT
            except Exception as e:                                   # noqa: BLE001 - report, fall back to the port
                print(f"[bench] reference module not loadable ({type(e).__name__}: {e}); timing the oracle port",
                      file=sys.stderr)
        self.state = isp_oracle.default_state(self.cam)

    def describe(self):
        if self.kind == "reference":
            return "unmodified reference ParametrizedProcessing (processing/pipeline_torch.py), stock torch CPU ops"
        return "oracle port (same ATen CPU ops as the reference module)"

    def seconds(self, batch, size, repeats):
        raw = self.syn.smooth_scene(batch, size, size, self.preset, seed=1234)
        g = torch.full((batch, 3, size, size), 1.0 / (batch * 3 * size * size))
        times = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            if self.mod is not None:
                x = raw.clone().requires_grad_(True)
                for p in self.mod.parameters():
                    p.grad = None
                self.mod(x).backward(g)
            else:
                self.oracle.forward_backward(raw, self.state, grad_out=g, raw_grad=True)
            times.append(time.perf_counter() - t0)
        return times


def numpy_baseline(batch, size, preset):
    """pipeline_numpy.processing (restated, oracle/numpy_oracle.py) over a batch: one process, and a pool of all cores
    (the reference runs it in 16 DataLoader workers, train.py:318).  Forward only -- the numpy chain has no backward."""
    import multiprocessing as mp
    from oracle import numpy_oracle
    from raw2logit_b200 import synthetic as syn
    cam = syn.CAMERA_PRESETS[preset]
    raws = list(syn.smooth_scene(batch, size, size, preset, seed=1234).numpy())
    numpy_oracle.process_batch(raws[:2], cam)
    t0 = time.perf_counter()
    numpy_oracle.process_batch(raws, cam)
    single = time.perf_counter() - t0
    cores = os.cpu_count() or 1
    with mp.get_context("spawn").Pool(cores) as pool:
        numpy_oracle.process_batch(raws, cam, pool)                  # start-up, imports
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            numpy_oracle.process_batch(raws, cam, pool)
        pooled = (time.perf_counter() - t0) / reps
    px = batch * size * size / 1e6
    return {"value": px / pooled, "unit": UNIT, "cores": cores, "kind": "port", "single_process": px / single,
            "sample": f"{batch}x{size}x{size} forward only (the numpy chain has no backward): pipeline_numpy.processing "
                      "(bilinear / sharpening_filter / gaussian_denoising) restated in oracle/numpy_oracle.py -- parity "
                      f"unpinned for its third-party parts; multiprocessing.Pool({cores}) vs one process"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cpu = CpuReference(args.preset)
    per_image = 0.006                                   # s, fwd+bwd of one 256^2 image on the box's host cores
    budget = 150.0
    scale = (args.size / 256.0) ** 2
    b = int(max(1, min(args.batch, budget / max(1, args.steps + args.warmup) / (per_image * scale))))
    cpu.seconds(b, args.size, max(1, min(args.warmup, 3)))
    times = cpu.seconds(b, args.size, args.steps)
    dt = sum(times) / len(times)
    value = b * args.size * args.size / dt / 1e6
    sample = f"{b}x{args.size}x{args.size} fp32 per step, fwd+bwd incl. raw grad, {cpu.describe()}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args), "sample_batch": b,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": cpu.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.stamps, self.window = [], (0.0, float("inf"))
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
                self.stamps.append(time.perf_counter())
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
        """Median over every sample taken under load (warm-up, the timed region, the per-kernel timing loops -- the same
        two kernels back to back throughout); `samples_in_timed_region` counts those inside the timed region itself,
        which at the driver's 20 steps lasts only ~2.5 ms."""
        inside = [c for c, t in zip(self.samples, self.stamps) if self.window[0] <= t <= self.window[1]]
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "samples_in_timed_region": len(inside),
                "sm_mhz_in_timed_region": statistics.median(inside) if inside else None}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs next to its GPU BEFORE the pinned staging buffers are allocated (first touch puts
    them on that NUMA node): at 8 ranks the host-to-device copies otherwise cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, wd in enumerate(mask) for b in range(64) if (wd >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from raw2logit_b200 import _lib, graphs, synthetic as syn
    from processing.pipeline_torch import ParametrizedProcessing

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    B, H, W, S = args.batch, args.size, args.size, args.sets
    pix = B * H * W
    cam = syn.CAMERA_PRESETS[args.preset]
    mod = ParametrizedProcessing(cam, batch_norm_output=False).to(dev)
    base = syn.smooth_scene(B, H, W, args.preset, seed=1234 + rank)
    vp0 = ctypes.c_void_p
    def pinned(t):
        """A page-locked copy (allocated as bytes: pin_memory() of some dtypes silently stays pageable, and a pageable
        source turns the 'non-blocking' copy into a staged, synchronous one)."""
        buf = torch.empty(t.numel() * t.element_size(), dtype=torch.uint8, pin_memory=True).view(t.dtype).view(t.shape)
        buf.copy_(t)
        assert buf.is_pinned()
        return buf

    host_raw = [pinned(torch.roll(base, shifts=2 * s, dims=0).contiguous()) for s in range(S)]
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

    def step_forward(i):
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

    def step_backward(s, with_raw=True, fused=True, grads=None):
        gp = gpar if grads is None else grads
        gr = vp(graws[s].data_ptr()) if with_raw else None
        if xch is not None and fused:
            d = xch.next(average=False)
            return lib.r2l_isp_backward_dp(vp(raws[s].data_ptr()), _lib.F32, 65535.0, B, H, W, ctypes.byref(params),
                                           vp(gouts[s].data_ptr()), None, None, vp(outs[s].data_ptr()),
                                           vp(lumas[s].data_ptr()), gr, vp(gp.data_ptr()), vp(wsb.data_ptr()), nws,
                                           ctypes.byref(d), sp)
        return lib.r2l_isp_backward(vp(raws[s].data_ptr()), _lib.F32, 65535.0, B, H, W, ctypes.byref(params),
                                    vp(gouts[s].data_ptr()), None, None, vp(outs[s].data_ptr()),
                                    vp(lumas[s].data_ptr()), gr, vp(gp.data_ptr()), vp(wsb.data_ptr()), nws, sp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return ms

    # ---- N > 1: the fused exchange against NCCL on the same inputs, before anything is timed ---------------------
    exchange_check = None
    if xch is not None:
        rc, s = step_forward(rank)                                   # different shards per rank (seed) and set per rank
        _lib.check(rc, "forward")
        g_fused, g_local = torch.empty_like(gpar), torch.empty_like(gpar)
        _lib.check(step_backward(s, grads=g_fused), "backward_dp")
        _lib.check(step_backward(s, fused=False, grads=g_local), "backward")
        dist.all_reduce(g_local)
        torch.cuda.synchronize()
        rel = ((g_fused - g_local).abs().max() / g_local.abs().max().clamp_min(1e-30)).item()
        gathered = [torch.empty_like(g_fused) for _ in range(world)]
        dist.all_gather(gathered, g_fused)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        relt = torch.tensor([rel], device=dev)
        dist.all_reduce(relt, op=dist.ReduceOp.MAX)
        exchange_check = {"max_rel_err_vs_nccl": relt.item(), "ranks_bit_identical": bool(same),
                          "finite": bool(torch.isfinite(g_fused).all().item())}
        if not same or relt.item() > 1e-5 or not exchange_check["finite"]:
            raise RuntimeError(f"fused gradient exchange disagrees with NCCL: {exchange_check}")

    # ---- device-resident timing (value) -----------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    for i in range(max(args.warmup, 50)):                            # >= 50: the clock sampler needs a few ms of load
        rc, s = step_forward(i)
        _lib.check(rc, "forward")
        _lib.check(step_backward(s), "backward")
        if world > 1 and xch is None:
            dist.all_reduce(gpar)
    K = args.steps
    barrier()
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t_begin.record()
    for i in range(K):
        rc, s = step_forward(i)
        rc2 = step_backward(s)
        if world > 1 and xch is None and os.environ.get("R2L_BENCH_NO_EXCHANGE", "0") != "1":
            dist.all_reduce(gpar)
    t_end.record()
    barrier()
    sampler.window = (w0, time.perf_counter())
    _lib.check(rc, "forward")
    _lib.check(rc2, "backward")
    ms_per_step = max_over_ranks(t_begin.elapsed_time(t_end)) / K
    value = world * pix / (ms_per_step * 1e-3) / 1e6

    # per-kernel launch durations for the roofline: same loop, CUDA events around each launch (kept out of the
    # loop above so the event records do not perturb the headline number)
    KR = max(3, min(K, 100))

    def per_kernel(with_raw):
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(KR)]
        barrier()
        for i in range(KR):
            ev[i][0].record()
            rc_, s_ = step_forward(i)
            ev[i][1].record()
            step_backward(s_, with_raw=with_raw)
            ev[i][2].record()
        barrier()
        return (statistics.mean(e[0].elapsed_time(e[1]) for e in ev), statistics.mean(e[1].elapsed_time(e[2]) for e in ev),
                statistics.mean(e[0].elapsed_time(e[2]) for e in ev))

    fwd_ms, bwd_ms, _ = per_kernel(True)
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    peak, peak_src = measured_peak()
    default_cfg = (B, H, W) == (64, 256, 256)          # the configuration the committed ncu capture was taken on
    bwd_gbs = BYTES_BWD * pix / (bwd_ms * 1e-3) / 1e9
    fwd_gbs = BYTES_FWD * pix / (fwd_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args),
        "gradient_exchange": (
            "none (1 GPU)" if world == 1 else
            "fused into the backward kernel: one-shot all-reduce of the 132 gradients over NVLink peer memory "
            "(r2l_isp_backward_dp)" if xch is not None else "NCCL all-reduce of the 132 gradients after the backward"),
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
    if exchange_check is not None:
        line["exchange_check"] = exchange_check
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if default_cfg and "step" in tj:
            line["roofline_step_traffic"] = tj["step"]
    except Exception:
        pass

    def finish():
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()

    if args.kernels_only:
        return finish()

    # ---- training case: no raw gradient (model.py:77-83 -- raw.requires_grad only in track_images, :229) ----------
    f2, b2, st2 = per_kernel(False)
    line["train_case"] = {"ms_per_step": st2, "forward_ms": f2, "backward_ms": b2,
                          "value": world * pix / (st2 * 1e-3) / 1e6, "unit": UNIT,
                          "algorithmic_bytes_per_pixel": BYTES_FWD + BYTES_BWD_NORAW,
                          "roofline_step_frac": (BYTES_FWD + BYTES_BWD_NORAW) * pix / (st2 * 1e-3) / 1e9 / peak,
                          "note": "forward + backward with need_raw_grad=False (132 parameter gradients only), CUDA "
                                  "events per step incl. the events' own gaps"}

    # ---- backward with a cold L2: a buffer larger than the L2 is written between forward and backward --------------
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    evc = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(20)]
    barrier()
    for i in range(20):
        rc, s = step_forward(i)
        flush.fill_(i & 1)
        evc[i][0].record()
        step_backward(s)
        evc[i][1].record()
    barrier()
    cold_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evc)
    line["backward_cold_l2"] = {"avg_launch_ms": cold_ms, "warm_avg_launch_ms": bwd_ms,
                                "frac": BYTES_BWD * pix / (cold_ms * 1e-3) / 1e9 / peak,
                                "note": "256 MiB written between the forward and the backward of each step"}

    # ---- through the public module API -----------------------------------------------------------------------
    fused_e2e = False
    if xch is not None:
        from raw2logit_b200 import parallel
        parallel.enable_fused_gradient_exchange(average=False)
        fused_e2e = True
    host_grads = pinned(torch.zeros(132, dtype=torch.float32))
    torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)   # GraphedStep warms up on a side stream
    plist = [p for p in mod.parameters()]
    copy_stream = torch.cuda.Stream()
    main = torch.cuda.current_stream()

    def timed(fn, steps, warm):
        fn(warm)
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(steps)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    # (a) device-resident input through ParametrizedProcessing + autograd: eager, then the captured graph
    def resident_eager(steps):
        for i in range(steps):
            x = raws[i % S].detach().requires_grad_(True)
            mod(x).backward(gouts[i % S])
            for p in plist:
                p.grad = None

    n_mod = max(3, min(K, 200))
    eager_ms = timed(resident_eager, n_mod, 10)
    gstep_f32 = graphs.GraphedStep(mod, raws[0], gouts[0], need_raw_grad=True)

    def resident_graphed(steps):
        for i in range(steps):
            gstep_f32.replay()

    graphed_ms = timed(resident_graphed, n_mod, 10)
    line["module_resident"] = {
        "ms_per_step_eager": eager_ms, "ms_per_step_graphed": graphed_ms, "kernels_ms_per_step": ms_per_step,
        "ratio_eager": eager_ms / ms_per_step, "ratio_graphed": graphed_ms / ms_per_step,
        "api": "ParametrizedProcessing.forward + autograd backward (C++ operator shim), device-resident fp32 batch with "
               "raw gradient; graphed = raw2logit_b200.graphs.GraphedStep (one static buffer set, so its working set "
               "fits the L2 -- compare eager, which rotates the 8 sets, with kernels_ms_per_step)"}

    # (b) end to end from pinned host memory: NBUF graphed steps (one static input buffer each), copies on a side
    # stream so that the copy of step i+1 / i+2 overlaps the kernels of step i, gradients read back every step
    NBUF = 3

    def make_e2e(host_batches, need_raw_grad, graphed):
        steps_g = [graphs.GraphedStep(mod, host_batches[0].to(dev), gouts[0], need_raw_grad=need_raw_grad)
                   for _ in range(NBUF)] if graphed else None
        bufs = [g.raw for g in steps_g] if graphed else [torch.empty_like(host_batches[0], device=dev) for _ in range(NBUF)]
        ready = [torch.cuda.Event() for _ in range(NBUF)]
        consumed = [torch.cuda.Event() for _ in range(NBUF)]

        def run(steps):
            for ev_ in consumed:
                ev_.record(main)

            def prefetch(i):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[i % NBUF])
                    with torch.no_grad():
                        bufs[i % NBUF].copy_(host_batches[i % len(host_batches)], non_blocking=True)
                    ready[i % NBUF].record(copy_stream)

            for i in range(min(NBUF - 1, steps)):
                prefetch(i)
            for i in range(steps):
                if i + NBUF - 1 < steps:
                    prefetch(i + NBUF - 1)
                main.wait_event(ready[i % NBUF])
                if graphed:
                    g = steps_g[i % NBUF]
                    g.replay()
                    flat = g.flat_grads
                else:
                    x = bufs[i % NBUF]
                    if need_raw_grad:
                        x = x.detach().requires_grad_(True)
                    mod(x).backward(gouts[i % S])
                    flat = torch.cat([p.grad.reshape(-1) for p in plist])
                    for p in plist:
                        p.grad = None
                consumed[i % NBUF].record(main)
                if world > 1 and not fused_e2e:
                    dist.all_reduce(flat)
                host_grads.copy_(flat, non_blocking=True)
        return run

    n_e2e = max(3, min(K, 200))
    host_u16 = [pinned(syn.to_uint16(h)) for h in host_raw]
    xnote = ("; gradients exchanged inside the backward kernel (parallel.enable_fused_gradient_exchange)" if fused_e2e
             else ("; NCCL all-reduce of the gradients" if world > 1 else ""))
    ms = timed(make_e2e(host_u16, False, True), n_e2e, 10)
    line["e2e"] = {"value": world * pix / (ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": pix * 2,
                   "d2h_bytes_per_step": 132 * 4, "steps": n_e2e, "ms_per_step": ms, "ingest": "uint16",
                   "api": "raw2logit_b200.graphs.GraphedStep over ParametrizedProcessing.forward + autograd backward: the "
                          "sensor's uint16 words (2 B/px; the reference divides them by 2^16-1 on the host, dataset.py:87, "
                          "here the kernel does) copied from pinned host memory every step on a copy stream, three "
                          "buffers in flight; 132 parameter gradients read back (an integer batch has no raw gradient: "
                          "the training case)" + xnote}
    ms = timed(make_e2e(host_raw, True, True), n_e2e, 10)
    line["e2e_fp32"] = {"value": world * pix / (ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": pix * 4,
                        "d2h_bytes_per_step": 132 * 4, "ms_per_step": ms,
                        "note": "same with the fp32 batch (4 B/px over PCIe) and the raw gradient"}
    ms = timed(make_e2e(host_u16, False, False), n_e2e, 10)
    line["e2e_eager"] = {"value": world * pix / (ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": pix * 2,
                         "d2h_bytes_per_step": 132 * 4, "ms_per_step": ms,
                         "note": "the uint16 step without graph capture (dispatcher + autograd engine every step)"}
    line["e2e"]["numa_bound_cpus"] = numa_cpus
    if fused_e2e:
        parallel.disable_fused_gradient_exchange()

    # ---- the reference's default BatchNorm tail (train.py:196), train mode, graphed module step ---------------------
    if world == 1:
        mod_bn = ParametrizedProcessing(cam, batch_norm_output=True).to(dev).train()
        g_bn = graphs.GraphedStep(mod_bn, raws[0], gouts[0], need_raw_grad=True)
        bn_ms = timed(lambda n: [g_bn.replay() for _ in range(n)], n_mod, 10)
        line["bn_tail"] = {"ms_per_step": bn_ms, "no_tail_ms_per_step": graphed_ms, "ratio": bn_ms / graphed_ms,
                           "note": "graphed module step with BatchNorm2d(3, affine=False) in train mode vs the same "
                                   "without the tail (both on one static buffer set)"}

    if rank == 0 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        cpu = CpuReference(args.preset)
        cb = min(B, 16)
        cpu.seconds(cb, H, 1)
        ts = cpu.seconds(cb, H, 3)
        line["cpu_baseline"] = {"value": cb * H * W / min(ts) / 1e6, "unit": UNIT, "cores": torch.get_num_threads(),
                                "kind": cpu.kind,
                                "sample": f"{cb}x{H}x{W} fp32 fwd+bwd incl. raw grad, best of 3, {cpu.describe()}"}
        try:
            line["cpu_baseline_numpy"] = numpy_baseline(min(B, 64), H, args.preset)
        except Exception as e:                                       # noqa: BLE001 - a baseline, not the product
            line["cpu_baseline_numpy"] = {"unavailable": f"{type(e).__name__}: {e}"}
    finish()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
