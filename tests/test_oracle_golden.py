"""Pins the CPU oracle (oracle/isp_oracle.py) to the reference's own outputs (tests/golden, and the live reference
when /root/reference is mounted).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import isp_oracle, ref_loader
from tests.golden_util import GoldenCase, case_names, maxabs, GOLDEN_DIR


@pytest.mark.parametrize("name", case_names())
def test_oracle_forward_matches_reference_fp32(name):
    c = GoldenCase(name)
    out, stages = isp_oracle.forward(c.raw, c.state, track_stages=c.track_stages, additive=c.additive,
                                     bn=c.bn_dict())
    # same arithmetic, possibly different op spelling: fp32 rounding only
    assert maxabs(out, c.f32["out"]) <= 2e-6, name
    for k, v in c.f32.items():
        if k.startswith("stage."):
            assert maxabs(stages[k[6:]], v) <= 2e-6, (name, k)


@pytest.mark.parametrize("name", case_names())
def test_oracle_fp64_matches_reference_fp64(name):
    c = GoldenCase(name)
    st = isp_oracle.cast_state(c.state, torch.float64)
    add = None if c.additive is None else c.additive.double()
    out, _ = isp_oracle.forward(c.raw.double(), st, track_stages=c.track_stages, additive=add,
                                bn=c.bn_dict(torch.float64), dtype=torch.float64)
    assert maxabs(out, c.f64["out"]) <= 1e-12, name


@pytest.mark.parametrize("cot", ["mean", "ramp"])
@pytest.mark.parametrize("name", case_names())
def test_oracle_gradients_match_reference_fp64(name, cot):
    c = GoldenCase(name)
    add = None if c.additive is None else c.additive.double()
    _, grads = isp_oracle.forward_backward(c.raw, c.state, grad_out=cot, dtype=torch.float64,
                                           track_stages=c.track_stages, additive=add,
                                           bn=c.bn_dict(torch.float64))
    for k, g in grads.items():
        ref = c.f64[f"grad.{cot}.{k}"]
        scale = max(1.0, float(np.max(np.abs(ref))))
        assert maxabs(g, ref) <= 1e-10 * scale, (name, k)


def test_mosaic_modes_bit_exact():
    z = np.load(f"{GOLDEN_DIR}/raw2rgb.f32.npz")
    for tag in ("even", "odd"):
        raw = torch.from_numpy(z[f"{tag}.raw"])
        bl = z[f"{tag}.black_level"].tolist()
        for rs in (True, False):
            if rs and tag == "odd":
                continue
            for ch in (3, 4):
                got = isp_oracle.mosaic(raw, None, rs, ch)
                assert np.array_equal(got.numpy(), z[f"{tag}.rs{int(rs)}.c{ch}"])
                got = isp_oracle.mosaic(raw, bl, rs, ch)
                assert np.array_equal(got.numpy(), z[f"{tag}.rs{int(rs)}.c{ch}.bl"])


def test_state_keys_and_shapes():
    st = isp_oracle.default_state()
    shapes = {k: tuple(v.shape) for k, v in st.items()}
    assert shapes == {"black_level": (4,), "white_balance": (1, 3), "colour_correction": (3, 3),
                      "gamma_correct": (1,), "M_RGB_2_YUV": (3, 3), "M_YUV_2_RGB": (3, 3),
                      "debayer.weight": (3, 3, 3, 3), "sharpening_filter.weight": (1, 1, 3, 3),
                      "gaussian_blur.weight": (1, 1, 5, 5)}


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_oracle_matches_live_reference():
    ref = ref_loader.load_reference()
    from raw2logit_b200 import synthetic as syn
    cam = syn.CAMERA_PRESETS["drone"]
    mod = ref.ParametrizedProcessing(cam, batch_norm_output=False)
    st = {k: v.clone() for k, v in mod.state_dict().items()}
    mine = isp_oracle.default_state(cam)
    for k in st:
        assert torch.equal(st[k], mine[k]), k
    raw = syn.smooth_scene(3, 40, 56, "drone", seed=3)
    with torch.no_grad():
        want = mod(raw)
    got, _ = isp_oracle.forward(raw, mine)
    assert maxabs(got, want) <= 2e-6
