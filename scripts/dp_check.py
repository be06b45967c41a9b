#!/usr/bin/env python
"""Two or more GPUs: the all-reduce fused into the backward kernel (r2l_isp_backward_dp) against r2l_isp_backward + an
NCCL all-reduce, on different shards per rank, several steps (both epoch parities), float and uint16 raw.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/dp_check.py
Prints "dp_check ok" on rank 0; any mismatch raises."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from raw2logit_b200 import _lib, parallel, synthetic as syn  # noqa: E402
from processing.pipeline_torch import ParametrizedProcessing  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    xch = parallel.PeerExchange()
    mod = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], batch_norm_output=False).to(dev)
    pt = [mod.black_level, mod.white_balance, mod.colour_correction, mod.gamma_correct, mod.debayer.weight,
          mod.sharpening_filter.weight, mod.gaussian_blur.weight, mod.M_RGB_2_YUV, mod.M_YUV_2_RGB]
    params = _lib.IspParams(*[t.data_ptr() for t in pt])
    vp = ctypes.c_void_p
    sp = vp(torch.cuda.current_stream().cuda_stream)
    worst = 0.0
    for step, (B, H, W, u16) in enumerate([(6, 96, 128, False), (3, 72, 136, False), (8, 256, 256, False), (5, 64, 64, True),
                                           (2, 40, 72, False), (64, 256, 256, False)]):
        raw = syn.smooth_scene(B, H, W, "drone", seed=100 * step + rank)
        raw = (syn.to_uint16(raw) if u16 else raw).to(dev)
        code = _lib.U16 if u16 else _lib.F32
        out = torch.empty(B, 3, H, W, device=dev)
        luma = torch.empty(lib.r2l_isp_saved_luma_floats(B, H, W), device=dev)
        gout = torch.randn(B, 3, H, W, device=dev, generator=torch.Generator(dev).manual_seed(7 + rank)) / (B * H * W)
        graw = torch.empty(B, H, W, device=dev)
        nws = lib.r2l_isp_workspace_bytes(B, H, W)
        ws = torch.empty(nws // 4, device=dev)
        g_ref = torch.empty(132, device=dev)
        g_dp = torch.empty(132, device=dev)
        _lib.check(lib.r2l_isp_forward(vp(raw.data_ptr()), code, 65535.0, B, H, W, ctypes.byref(params), None,
                                       vp(out.data_ptr()), vp(luma.data_ptr()), sp), "forward")
        _lib.check(lib.r2l_isp_backward(vp(raw.data_ptr()), code, 65535.0, B, H, W, ctypes.byref(params),
                                        vp(gout.data_ptr()), None, None, vp(out.data_ptr()), vp(luma.data_ptr()),
                                        vp(graw.data_ptr()), vp(g_ref.data_ptr()), vp(ws.data_ptr()), nws, sp), "backward")
        dist.all_reduce(g_ref)
        for rep in range(3):                                         # both epoch parities, back to back
            d = xch.next(average=False)
            _lib.check(lib.r2l_isp_backward_dp(vp(raw.data_ptr()), code, 65535.0, B, H, W, ctypes.byref(params),
                                               vp(gout.data_ptr()), None, None, vp(out.data_ptr()), vp(luma.data_ptr()),
                                               vp(graw.data_ptr()), vp(g_dp.data_ptr()), vp(ws.data_ptr()), nws,
                                               ctypes.byref(d), sp), "backward_dp")
            torch.cuda.synchronize()
            err = (g_dp - g_ref).abs().max().item() / max(1.0, g_ref.abs().max().item())
            worst = max(worst, err)
            assert err <= 1e-5, (step, rep, err)
            gathered = [torch.empty_like(g_dp) for _ in range(world)]
            dist.all_gather(gathered, g_dp)
            assert all(torch.equal(gathered[0], g) for g in gathered), "ranks disagree bitwise"
    # the same through the module API: autograd backward with the exchange enabled == local gradients + NCCL average
    raw = syn.smooth_scene(4, 128, 192, "drone", seed=900 + rank).to(dev)
    gout = torch.randn(4, 3, 128, 192, device=dev, generator=torch.Generator(dev).manual_seed(70 + rank)) / raw.numel()
    plist = list(mod.parameters())

    def grads():
        for p in plist:
            p.grad = None
        x = raw.clone().requires_grad_(True)
        mod(x).backward(gout)
        return torch.cat([p.grad.reshape(-1) for p in plist]), x.grad

    g_local, gx_local = grads()
    dist.all_reduce(g_local)
    g_local /= world
    parallel.enable_fused_gradient_exchange(average=True)
    for rep in range(2):
        g_fused, gx_fused = grads()
        torch.cuda.synchronize()
        err = (g_fused - g_local).abs().max().item() / max(1.0, g_local.abs().max().item())
        worst = max(worst, err)
        assert err <= 1e-5, ("module API", rep, err)
        assert torch.equal(gx_fused, gx_local), "the raw gradient must stay local"
    parallel.disable_fused_gradient_exchange()
    if rank == 0:
        print(f"dp_check ok: world {world}, worst relative error vs NCCL {worst:.2e}, ranks bit-identical", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
