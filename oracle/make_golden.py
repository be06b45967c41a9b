"""Generates ``tests/golden/*.npz`` by running the UNMODIFIED reference -- TEST INFRASTRUCTURE.

Run in the build container (needs ``/root/reference``):   python oracle/make_golden.py
The fp64 "truth" twin of every case is produced in a child process (``--fp64``) because the reference's constants
follow the process-wide default dtype (see oracle/ref_loader.py).

Each fixture stores the inputs (raw, every parameter/buffer, flags) and, from the reference module itself:
the fp32 forward, the fp32 gradients of every parameter and of raw under two cotangents ('mean', 'ramp'), the same
in fp64, and -- where the case asks -- stage tensors, stage gradients and BatchNorm running statistics.
The reference has no golden vectors of its own for this path; these are its outputs, pinned.
"""
import argparse
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import isp_oracle, ref_loader                     # noqa: E402
from raw2logit_b200 import synthetic as syn                    # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
PARAM_KEYS = isp_oracle.PARAM_KEYS


def _demo_crop(name, y0, x0, size):
    """Real-data inputs: the reference's demo mosaics (app.py:31 uses channel 0 / 255)."""
    from PIL import Image
    img = np.asarray(Image.open(os.path.join(ref_loader.REF_ROOT, "demo-files", name)))
    plane = img[..., 0] if img.ndim == 3 else img
    crop = plane[y0:y0 + size, x0:x0 + size].astype(np.float32) / 255.0
    return torch.from_numpy(crop).unsqueeze(0).contiguous()


def cases():
    """name -> dict(raw, preset, perturb, flags).  Small shapes: the whole set stays ~1 MB."""
    c = {}
    c["drone_g1"] = dict(raw=syn.smooth_scene(2, 64, 64, "drone"), preset="drone", perturb=False)
    c["drone_g1_pert"] = dict(raw=syn.smooth_scene(2, 64, 64, "drone"), preset="drone", perturb=True)
    c["micro_g1_pert"] = dict(raw=syn.smooth_scene(2, 64, 64, "microscopy"), preset="microscopy", perturb=True)
    c["default_g1"] = dict(raw=syn.smooth_scene(1, 32, 48, "default"), preset="default", perturb=False)
    c["noise_g2_pert"] = dict(raw=syn.noise_stress(1, 32, 32), preset="drone", perturb=True)
    c["odd_31x34"] = dict(raw=syn.smooth_scene(2, 31, 34, "drone"), preset="drone", perturb=True)
    c["odd_33x31"] = dict(raw=syn.smooth_scene(1, 33, 31, "drone"), preset="drone", perturb=True)
    c["tiny_3x3"] = dict(raw=syn.smooth_scene(1, 3, 3, "drone"), preset="drone", perturb=True)
    c["tiny_4x5"] = dict(raw=syn.smooth_scene(2, 4, 5, "drone"), preset="drone", perturb=True)
    c["tiny_6x6"] = dict(raw=syn.smooth_scene(1, 6, 6, "drone"), preset="drone", perturb=True)
    c["wide_8x200"] = dict(raw=syn.smooth_scene(1, 8, 200, "drone"), preset="drone", perturb=True)
    c["tall_150x8"] = dict(raw=syn.smooth_scene(1, 150, 8, "drone"), preset="drone", perturb=True)
    c["impulses"] = dict(raw=syn.impulses(16, 16, syn.impulse_positions(16, 16)), preset="drone", perturb=True)
    c["car_crop"] = dict(raw=_demo_crop("car.png", 96, 64, 64), preset="drone", perturb=False)
    c["micro_crop"] = dict(raw=_demo_crop("micro.png", 64, 96, 64), preset="microscopy", perturb=False)
    c["stages_pert"] = dict(raw=syn.smooth_scene(2, 32, 32, "drone"), preset="drone", perturb=True,
                            track_stages=True)
    c["additive"] = dict(raw=syn.smooth_scene(2, 32, 32, "drone"), preset="drone", perturb=True, additive=True)
    c["bn_train"] = dict(raw=syn.smooth_scene(4, 32, 32, "drone"), preset="drone", perturb=True, bn="train")
    c["bn_eval"] = dict(raw=syn.smooth_scene(2, 32, 32, "drone"), preset="drone", perturb=True, bn="eval")
    c["bn_train_additive"] = dict(raw=syn.smooth_scene(3, 32, 32, "drone"), preset="drone", perturb=True,
                                  bn="train", additive=True)
    return c


def build_state(case):
    st = isp_oracle.default_state(syn.CAMERA_PRESETS[case["preset"]])
    if case["perturb"]:
        st = syn.perturbed_state(st)
    return st


def case_extras(case):
    """Deterministic extra inputs (additive layer, BN running statistics)."""
    g = torch.Generator().manual_seed(99)
    ex = {}
    if case.get("additive"):
        h, w = case["raw"].shape[1:]
        ex["additive"] = 0.01 * torch.randn(1, 3, h, w, generator=g)
    if case.get("bn"):
        ex["running_mean"] = 0.4 + 0.1 * torch.rand(3, generator=g)
        ex["running_var"] = 0.02 + 0.02 * torch.rand(3, generator=g)
    return ex


def run_reference(ref, case, state, extras, dtype):
    """One fresh reference module per cotangent; returns a flat dict of numpy arrays."""
    out = {}
    for cot in ("mean", "ramp"):
        mod = ref.ParametrizedProcessing(syn.CAMERA_PRESETS[case["preset"]],
                                         track_stages=bool(case.get("track_stages")),
                                         batch_norm_output=bool(case.get("bn")))
        sd = {k: v.to(dtype) for k, v in state.items()}
        if case.get("bn"):
            sd["batch_norm.running_mean"] = extras["running_mean"].to(dtype)
            sd["batch_norm.running_var"] = extras["running_var"].to(dtype)
            sd["batch_norm.num_batches_tracked"] = torch.tensor(3)
        mod.load_state_dict(sd, strict=True)
        if case.get("additive"):
            mod.additive_layer = torch.nn.Parameter(extras["additive"].to(dtype).clone())
        mod.train(case.get("bn") != "eval")
        raw = case["raw"].to(dtype).clone().requires_grad_(True)
        y = mod(raw)
        g = isp_oracle.cotangent(tuple(y.shape), cot, dtype)
        y.backward(g)
        if cot == "mean":
            out["out"] = y.detach().numpy()
            if case.get("track_stages"):
                for name, t in mod.stages.items():
                    out[f"stage.{name}"] = t.detach().numpy()
            if case.get("bn"):
                out["bn.running_mean"] = mod.batch_norm.running_mean.numpy().copy()
                out["bn.running_var"] = mod.batch_norm.running_var.numpy().copy()
        params = dict(mod.named_parameters())
        for k in PARAM_KEYS:
            out[f"grad.{cot}.{k}"] = params[k].grad.numpy().copy()
        out[f"grad.{cot}.raw"] = raw.grad.numpy().copy()
        if case.get("additive"):
            out[f"grad.{cot}.additive"] = mod.additive_layer.grad.numpy().copy()
        if case.get("track_stages"):
            for name, t in mod.stages.items():
                out[f"stagegrad.{cot}.{name}"] = t.grad.numpy().copy()
    return out


def run_raw2rgb(ref):
    """``raw2rgb`` / ``RawToRGB`` in every mode (pipeline_torch.py:43-80, 240-283), incl. odd sizes."""
    out = {}
    g = torch.Generator().manual_seed(5)
    for tag, (h, w) in {"even": (8, 12), "odd": (7, 9)}.items():
        raw = syn.quantise16(torch.rand(2, h, w, generator=g)).float()
        out[f"{tag}.raw"] = raw.numpy()
        bl = [0.0625, 0.0626, 0.0627, 0.0628]
        out[f"{tag}.black_level"] = np.asarray(bl, dtype=np.float32)
        for rs in (True, False):
            if rs and tag == "odd":      # the reference raises RuntimeError for odd H/W with reduce_size=True
                continue
            for ch in (3, 4):
                out[f"{tag}.rs{int(rs)}.c{ch}"] = ref.raw2rgb(raw, reduce_size=rs, out_channels=ch).numpy()
                out[f"{tag}.rs{int(rs)}.c{ch}.bl"] = ref.raw2rgb(raw, black_level=bl, reduce_size=rs,
                                                                out_channels=ch).numpy()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fp64", action="store_true")
    ap.add_argument("--out", default=GOLDEN_DIR)
    args = ap.parse_args()
    dtype = torch.float64 if args.fp64 else torch.float32
    os.makedirs(args.out, exist_ok=True)
    suffix = "f64" if args.fp64 else "f32"
    if not args.fp64:
        # inputs are generated once, under the default fp32 dtype, and stored; the fp64 child re-reads them
        for name, case in cases().items():
            state = build_state(case)
            extras = case_extras(case)
            inputs = {"raw": case["raw"].numpy()}
            inputs.update({f"state.{k}": v.numpy() for k, v in state.items()})
            inputs.update({f"extra.{k}": v.numpy() for k, v in extras.items()})
            inputs["preset"] = np.asarray(case["preset"])
            inputs["flags"] = np.asarray([int(bool(case.get("track_stages"))), int(bool(case.get("additive"))),
                                          {None: 0, "train": 1, "eval": 2}[case.get("bn")]], dtype=np.int32)
            np.savez_compressed(os.path.join(args.out, f"{name}.in.npz"), **inputs)
    ref = ref_loader.load_reference(fp64=args.fp64)
    import glob
    for path in sorted(glob.glob(os.path.join(args.out, "*.in.npz"))):
        name = os.path.basename(path)[:-len(".in.npz")]
        ins = np.load(path)
        flags = ins["flags"]
        case = dict(raw=torch.from_numpy(ins["raw"]), preset=str(ins["preset"]),
                    track_stages=bool(flags[0]), additive=bool(flags[1]),
                    bn={0: None, 1: "train", 2: "eval"}[int(flags[2])])
        state = {k[6:]: torch.from_numpy(ins[k]) for k in ins.files if k.startswith("state.")}
        extras = {k[6:]: torch.from_numpy(ins[k]) for k in ins.files if k.startswith("extra.")}
        res = run_reference(ref, case, state, extras, dtype)
        np.savez_compressed(os.path.join(args.out, f"{name}.{suffix}.npz"), **res)
        print(f"[{suffix}] {name}: out {res['out'].shape}")
    if not args.fp64:
        np.savez_compressed(os.path.join(args.out, "raw2rgb.f32.npz"), **run_raw2rgb(ref))
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--fp64", "--out", args.out])


if __name__ == "__main__":
    main()
