import os, sys, torch
sys.path.insert(0, os.getcwd())
from raw2logit_b200 import synthetic as syn
from processing.pipeline_torch import ParametrizedProcessing
dev = torch.device("cuda:0")
raw = syn.smooth_scene(64, 256, 256, "drone", seed=1).to(dev)
mod = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], batch_norm_output=True).to(dev).train()
with torch.no_grad():
    for _ in range(10): mod(raw)
    torch.cuda.synchronize()
    torch.cuda._sleep(2_000_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): mod(raw)
    e1.record(); torch.cuda.synchronize()
print("bn forward us", 1e3 * e0.elapsed_time(e1) / 200)
