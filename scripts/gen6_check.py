"""GPU check of the sixth-generation backward against the fifth (R2L_ISP_BWD_GEN=5) and the fp64 oracle.
usage (GPU box): timeout 300 python scripts/gen6_check.py"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import isp_oracle
from raw2logit_b200 import synthetic as syn
from processing.pipeline_torch import ParametrizedProcessing

state = syn.perturbed_state(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))


def grads(shape, gen, need_raw=True, u16=False, bn=False, add=False):
    if gen == 5:
        os.environ["R2L_ISP_BWD_GEN"] = "5"
    else:
        os.environ.pop("R2L_ISP_BWD_GEN", None)
    raw = syn.smooth_scene(*shape, "drone", seed=31)
    x0 = syn.to_uint16(raw).cuda() if u16 else raw.cuda()
    g = isp_oracle.cotangent((shape[0], 3, shape[1], shape[2]), "ramp").cuda()
    mod = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], batch_norm_output=bn)
    mod.load_state_dict(state, strict=not bn)
    if add:
        torch.manual_seed(3)
        mod.additive_layer = torch.nn.Parameter(0.01 * torch.randn(1, 3, shape[1], shape[2]))
    mod = mod.cuda().train()
    x = x0.clone().requires_grad_(True) if (need_raw and not u16) else x0
    mod(x).backward(g)
    torch.cuda.synchronize()
    flat = torch.cat([p.grad.flatten() for p in mod.parameters()]).cpu()
    return flat, (x.grad.cpu() if (need_raw and not u16) else None)


bad = 0
for shape in [(2, 64, 64), (5, 256, 256), (2, 96, 200), (3, 72, 136), (1, 8, 8), (2, 37, 8), (64, 256, 256), (2, 1024, 1024)]:
    for kw in ({}, {"need_raw": False}, {"u16": True}, {"bn": True}, {"add": True}):
        t0 = time.time()
        p6, r6 = grads(shape, 6, **kw)
        p6b, r6b = grads(shape, 6, **kw)
        p5, r5 = grads(shape, 5, **kw)
        ep = float((p6 - p5).abs().max() / max(1.0, float(p5.abs().max())))
        er = 0.0 if r6 is None else float((r6 - r5).abs().max() / max(1.0, float(r5.abs().max())))
        rep = torch.equal(p6, p6b) and (r6 is None or torch.equal(r6, r6b))
        ok = ep <= 2e-5 and er <= 2e-5 and rep and bool(torch.isfinite(p6).all())
        bad += 0 if ok else 1
        print("OK " if ok else "BAD", shape, kw, "param err %.2e raw err %.2e reproducible %s  %.2fs" % (ep, er, rep, time.time() - t0), flush=True)
print("FAILED" if bad else "ALL OK", bad)
sys.exit(1 if bad else 0)
