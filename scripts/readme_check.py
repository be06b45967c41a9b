"""The README quick-start snippet, run as a check (from anywhere: the repo root goes on sys.path first)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from processing.pipeline_torch import ParametrizedProcessing
from raw2logit_b200 import synthetic as syn
isp = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], track_stages=False, batch_norm_output=True).cuda()
raw = syn.smooth_scene(64, 256, 256, "drone", seed=0).cuda()
rgb = isp(raw)
rgb.mean().backward()
from raw2logit_b200.graphs import GraphedStep
step = GraphedStep(isp, raw, torch.full_like(rgb, 1e-3))
step.raw.copy_(raw, non_blocking=True); step.replay(); grads = step.flat_grads
torch.cuda.synchronize()
print("readme snippet ok", tuple(rgb.shape), grads.shape, bool(torch.isfinite(grads).all()))
