import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import isp_oracle
from raw2logit_b200 import synthetic as syn
from processing.pipeline_torch import ParametrizedProcessing
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
state = syn.perturbed_state(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))
def run(shape, force, need_raw, u16=False, bn=False):
    os.environ["R2L_ISP_FORCE_GENERIC"] = force
    raw = syn.smooth_scene(*shape, "drone", seed=31)
    x0 = syn.to_uint16(raw).cuda() if u16 else raw.cuda()
    g = isp_oracle.cotangent((shape[0], 3, shape[1], shape[2]), "ramp").cuda()
    mod = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], batch_norm_output=bn)
    mod.load_state_dict(state, strict=not bn)
    mod = mod.cuda().train()
    x = x0.clone().requires_grad_(True) if need_raw else x0
    try:
        mod(x).backward(g)
        torch.cuda.synchronize()
        flat = torch.cat([p.grad.flatten() for p in mod.parameters()]).cpu()
        print("OK ", shape, "force", force, "need_raw", need_raw, "u16", u16, "bn", bn, float(flat.abs().sum()), flush=True)
    except Exception as e:
        print("ERR", shape, "force", force, "need_raw", need_raw, "u16", u16, "bn", bn, str(e)[:80], flush=True)
        sys.exit(1)
only_u16 = len(sys.argv) > 1 and sys.argv[1] == "u16"
for shape in [(2, 64, 64), (5, 256, 256), (2, 96, 200)]:
    for u16 in ((True,) if only_u16 else (False, True)):
        for need_raw in ((True, False) if not u16 else (False,)):
            for force in ("0", "1"):
                run(shape, force, need_raw, u16)
