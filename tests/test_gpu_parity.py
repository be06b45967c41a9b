"""Parity of the sm_100a kernels (through the module / C ABI) against the reference's outputs (tests/golden) and
the CPU oracle on seeded inputs.  Tolerances are BASELINE.json's: forward max-abs 1e-5 (+1e-5 relative),
parameter gradients 1e-4; CFA indexing bit-exact."""
import numpy as np
import pytest
import torch

from oracle import isp_oracle
from raw2logit_b200 import synthetic as syn
from tests.golden_util import GoldenCase, case_names, maxabs

pytestmark = pytest.mark.gpu

FWD_ATOL, FWD_RTOL, GRAD_ATOL = 1e-5, 1e-5, 1e-4
FUSED_CASES = [n for n in case_names() if not GoldenCase(n).track_stages]


def _module(state, preset="drone", bn=False, dev="cuda"):
    from processing.pipeline_torch import ParametrizedProcessing
    mod = ParametrizedProcessing(syn.CAMERA_PRESETS[preset], batch_norm_output=bn)
    mod.load_state_dict(state, strict=not bn)
    return mod.to(dev)


def _close_fwd(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return bool(np.all(np.abs(got - want) <= FWD_ATOL + FWD_RTOL * np.abs(want)))


def _run_case(c, cot=None):
    from processing.pipeline_torch import ParametrizedProcessing
    mod = ParametrizedProcessing(batch_norm_output=c.bn is not None)
    sd = dict(c.state)
    if c.bn is not None:
        sd["batch_norm.running_mean"] = c.extra["running_mean"].clone()
        sd["batch_norm.running_var"] = c.extra["running_var"].clone()
        sd["batch_norm.num_batches_tracked"] = torch.tensor(3)
    mod.load_state_dict(sd, strict=True)
    if c.additive is not None:
        mod.additive_layer = torch.nn.Parameter(c.additive.clone())
    mod = mod.cuda()
    mod.train(c.bn != "eval")
    x = c.raw.cuda().requires_grad_(cot is not None)
    out = mod(x)
    if cot is not None:
        out.backward(isp_oracle.cotangent(tuple(out.shape), cot).cuda())
    return mod, x, out


@pytest.mark.parametrize("name", FUSED_CASES)
def test_forward_matches_reference_golden(name):
    c = GoldenCase(name)
    mod, _, out = _run_case(c)
    out = out.detach().cpu().numpy()
    assert np.isfinite(out).all()
    scale = 1.0 if c.bn is None else 8.0     # BN divides by sqrt(var) ~ 0.15: errors and tolerance scale alike
    err = maxabs(out, c.f32["out"])
    assert err <= scale * FWD_ATOL + FWD_RTOL * np.abs(c.f32["out"]).max(), (name, err)
    if c.bn == "train":
        assert maxabs(mod.batch_norm.running_mean.cpu().numpy(), c.f32["bn.running_mean"]) <= 1e-6
        assert maxabs(mod.batch_norm.running_var.cpu().numpy(), c.f32["bn.running_var"]) <= 1e-6


@pytest.mark.parametrize("cot", ["mean", "ramp"])
@pytest.mark.parametrize("name", FUSED_CASES)
def test_gradients_match_reference_golden(name, cot):
    c = GoldenCase(name)
    mod, x, _ = _run_case(c, cot)
    named = dict(mod.named_parameters())
    for k in isp_oracle.PARAM_KEYS:
        got = named[k].grad.cpu().numpy()
        assert np.isfinite(got).all(), (name, k)
        assert maxabs(got, c.f64[f"grad.{cot}.{k}"]) <= GRAD_ATOL, (name, k)
        assert maxabs(got, c.f32[f"grad.{cot}.{k}"]) <= GRAD_ATOL, (name, k)
    ref = c.f64[f"grad.{cot}.raw"]
    assert maxabs(x.grad.cpu().numpy(), ref) <= GRAD_ATOL * max(1.0, float(np.abs(ref).max())), name
    if c.additive is not None:
        assert maxabs(mod.additive_layer.grad.cpu().numpy(), c.f64[f"grad.{cot}.additive"]) <= GRAD_ATOL


def test_cfa_indexing_is_bit_exact_on_impulses():
    """One-hot raw at every CFA phase / corner / edge: the set of non-zero pre-gamma responses must be identical
    to the reference's, i.e. demosaic indexing, pattern handling and the three border rules are exact."""
    c = GoldenCase("impulses")
    _, _, out = _run_case(c)
    got = out.detach().cpu().numpy()
    want = c.f32["out"]
    floor = np.float32(1e-5) ** np.float32(1 / c.state["gamma_correct"].item())
    # a response is "on" where the output differs from the all-clipped floor value
    on_got = np.abs(got - floor) > 2e-6
    on_want = np.abs(want - floor) > 2e-6
    assert np.array_equal(on_got, on_want)


@pytest.mark.parametrize("shape,preset", [((8, 256, 256), "drone"), ((64, 256, 256), "drone"),
                                          ((3, 200, 328), "microscopy"), ((2, 255, 258), "drone"),
                                          ((1, 1024, 1024), "drone")])
def test_seeded_parity_with_oracle(shape, preset):
    raw = syn.smooth_scene(*shape, preset, seed=1234)
    state = syn.perturbed_state(isp_oracle.default_state(syn.CAMERA_PRESETS[preset]))
    want, grads = isp_oracle.forward_backward(raw, state, grad_out="mean")
    mod = _module(state, preset)
    x = raw.cuda().requires_grad_(True)
    out = mod(x)
    out.mean().backward()
    assert _close_fwd(out.detach().cpu().numpy(), want.numpy())
    named = dict(mod.named_parameters())
    for k in isp_oracle.PARAM_KEYS:
        assert maxabs(named[k].grad.cpu(), grads[k]) <= GRAD_ATOL, k
    assert maxabs(x.grad.cpu(), grads["raw"]) <= GRAD_ATOL * max(1.0, grads["raw"].abs().max().item())


def test_noise_stress_no_worse_than_twice_the_reference_fp32_error():
    c = GoldenCase("noise_g2_pert")
    _, _, out = _run_case(c)
    floor = maxabs(c.f32["out"], c.f64["out"])
    assert maxabs(out.detach().cpu().numpy(), c.f64["out"]) <= 2 * floor


def test_large_frame_crops_match_oracle():
    """Full-size property (4096^2 is too slow for the CPU oracle): an output window depends on raw within +-4 only,
    so interior windows of the big frame must equal the oracle run on the padded crop."""
    h = w = 4096
    raw = syn.smooth_scene(1, h, w, "drone", seed=77)
    state = syn.perturbed_state(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))
    mod = _module(state)
    with torch.no_grad():
        out = mod(raw.cuda()).cpu()
    for (y0, x0) in [(0, 0), (1000, 2046), (4096 - 72, 4096 - 72), (31, 63), (2048 - 36, 0)]:
        y1, x1 = min(y0 + 72, h), min(x0 + 72, w)
        ya, xa = max(y0 - 8, 0) & ~1, max(x0 - 8, 0) & ~1            # even origin keeps the CFA phase
        yb, xb = min(y1 + 8, h), min(x1 + 8, w)
        crop, _ = isp_oracle.forward(raw[:, ya:yb, xa:xb], state)
        # compare only sites at least 4 away from an artificial crop edge
        my0 = 0 if ya == 0 else 4
        mx0 = 0 if xa == 0 else 4
        my1 = (yb - ya) if yb == h else (yb - ya - 4)
        mx1 = (xb - xa) if xb == w else (xb - xa - 4)
        got = out[:, :, ya + my0:ya + my1, xa + mx0:xa + mx1]
        assert _close_fwd(got.numpy(), crop[:, :, my0:my1, mx0:mx1].numpy()), (y0, x0)


def test_backward_is_linear_in_the_cotangent_and_deterministic():
    raw = syn.smooth_scene(16, 256, 256, "drone", seed=9).cuda()
    state = syn.perturbed_state(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))
    mod = _module(state)
    g1 = torch.rand(16, 3, 256, 256, device="cuda") / raw.numel()
    g2 = torch.rand(16, 3, 256, 256, device="cuda") / raw.numel()

    def run(g):
        mod.zero_grad(set_to_none=True)
        x = raw.clone().requires_grad_(True)
        mod(x).backward(g)
        return torch.cat([p.grad.flatten() for p in mod.parameters()]), x.grad

    pa, ra = run(g1)
    pb, rb = run(g2)
    pc, rc = run(g1 + g2)
    assert maxabs(pc.cpu(), (pa + pb).cpu()) <= 1e-6 * max(1.0, pc.abs().max().item())
    assert maxabs(rc.cpu(), (ra + rb).cpu()) <= 1e-6 * max(1.0, rc.abs().max().item())
    pa2, ra2 = run(g1)
    assert torch.equal(pa, pa2) and torch.equal(ra, ra2)       # bit-reproducible run to run


def test_batch_items_are_independent():
    raw = syn.smooth_scene(5, 130, 70, "drone", seed=4).cuda()
    mod = _module(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))
    with torch.no_grad():
        full = mod(raw)
        for i in range(5):
            assert torch.equal(full[i:i + 1], mod(raw[i:i + 1]))


def test_uint16_ingest_matches_float_path():
    raw = syn.smooth_scene(4, 128, 192, "drone", seed=5)
    u16 = syn.to_uint16(raw)
    mod = _module(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))
    as_float = u16.to(torch.int32).to(torch.float32) / 65535.0
    with torch.no_grad():
        a = mod(u16.cuda())
        b = mod(as_float.cuda())
    assert torch.equal(a, b)
    want, _ = isp_oracle.forward(as_float, isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))
    assert _close_fwd(a.cpu().numpy(), want.numpy())


def test_raw2rgb_modes_bit_exact():
    from processing.pipeline_torch import RawToRGB, raw2rgb
    from tests.golden_util import GOLDEN_DIR
    z = np.load(f"{GOLDEN_DIR}/raw2rgb.f32.npz")
    for tag in ("even", "odd"):
        raw = torch.from_numpy(z[f"{tag}.raw"]).cuda()
        bl = z[f"{tag}.black_level"].tolist()
        for rs in (True, False):
            if rs and tag == "odd":
                with pytest.raises(RuntimeError):
                    raw2rgb(raw, reduce_size=True)
                continue
            for ch in (3, 4):
                assert np.array_equal(raw2rgb(raw, reduce_size=rs, out_channels=ch).cpu().numpy(),
                                      z[f"{tag}.rs{int(rs)}.c{ch}"])
                assert np.array_equal(raw2rgb(raw, black_level=bl, reduce_size=rs, out_channels=ch).cpu().numpy(),
                                      z[f"{tag}.rs{int(rs)}.c{ch}.bl"])
    # adjoint of the split w.r.t. raw against torch autograd over the oracle restatement
    raw = torch.from_numpy(z["even.raw"])
    for rs in (True, False):
        for ch in (3, 4):
            x = raw.clone().requires_grad_(True)
            y = isp_oracle.mosaic(x, None, rs, ch)
            g = torch.rand_like(y)
            y.backward(g)
            xc = raw.cuda().requires_grad_(True)
            m = RawToRGB(reduce_size=rs, out_channels=ch)
            m(xc).backward(g.cuda())
            assert torch.equal(xc.grad.cpu(), x.grad)


def test_errors_follow_the_reference():
    from processing.pipeline_torch import ParametrizedProcessing
    mod = ParametrizedProcessing(batch_norm_output=False).cuda()
    with pytest.raises(AssertionError):
        mod(torch.rand(1, 1, 8, 8, device="cuda"))
    with pytest.raises(RuntimeError):
        mod(torch.rand(1, 2, 8, device="cuda"))
    with pytest.raises(RuntimeError):
        mod.cpu()(torch.rand(1, 8, 8))


def test_frozen_processor_runs_forward_only():
    from processing.pipeline_torch import ParametrizedProcessing
    mod = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"]).cuda().eval()
    for p in mod.parameters():
        p.requires_grad = False
    out = mod(syn.smooth_scene(2, 64, 64).cuda())
    assert not out.requires_grad and out.shape == (2, 3, 64, 64)
    assert mod.buffer["processed_rgb"] is out


def test_tma_loader_and_generic_loader_agree_bitwise():
    """Aligned shapes take the TMA staging path (UTMALDG + mbarrier); R2L_ISP_NO_TMA=1 forces the generic loader."""
    import os
    state = syn.perturbed_state(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))
    mod = _module(state)
    for shape in [(5, 256, 256), (2, 96, 200), (1, 40, 8), (3, 1024, 512)]:
        raw = syn.smooth_scene(*shape, "drone", seed=21).cuda()
        u16 = syn.to_uint16(raw.cpu()).cuda() if shape[2] % 8 == 0 else None
        with torch.no_grad():
            a = mod(raw)
            au = mod(u16) if u16 is not None else None
            os.environ["R2L_ISP_NO_TMA"] = "1"
            try:
                b = mod(raw)
                bu = mod(u16) if u16 is not None else None
            finally:
                os.environ["R2L_ISP_NO_TMA"] = "0"
        assert torch.equal(a, b), shape
        if au is not None:
            assert torch.equal(au, bu), shape


@pytest.mark.parametrize("shape", [(5, 256, 256), (2, 96, 200), (3, 72, 136), (1, 40, 8), (2, 37, 8), (4, 128, 192),
                                   (2, 68, 132), (1, 33, 68)])
@pytest.mark.parametrize("bn", [False, True])
def test_vectorised_backward_agrees_with_generic_backward(shape, bn):
    """The third-generation backward (padded-domain phases, fold passes, TMA-fed) against the generic scalar kernel
    (R2L_ISP_FORCE_GENERIC=1) on multi-tile shapes, odd batches, partial tiles, uint16 raw and the BatchNorm tail."""
    import os
    from processing.pipeline_torch import ParametrizedProcessing
    state = syn.perturbed_state(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))
    raw = syn.smooth_scene(*shape, "drone", seed=31)
    inputs = [raw.cuda()]
    if shape[2] % 8 == 0:
        inputs.append(syn.to_uint16(raw).cuda())
    g = isp_oracle.cotangent((shape[0], 3, shape[1], shape[2]), "ramp").cuda()
    for x0 in inputs:
        for need_raw in ([True, False] if x0.dtype == torch.float32 else [False]):
            res = []
            for force in ("0", "1"):
                os.environ["R2L_ISP_FORCE_GENERIC"] = force
                try:
                    torch.manual_seed(0)
                    mod = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], batch_norm_output=bn)
                    mod.load_state_dict(state, strict=not bn)
                    mod = mod.cuda().train()
                    x = x0.clone().requires_grad_(True) if need_raw else x0
                    mod(x).backward(g)
                    flat = torch.cat([p.grad.flatten() for p in mod.parameters()]).cpu()
                    res.append((flat, x.grad.cpu() if need_raw else None))
                finally:
                    os.environ["R2L_ISP_FORCE_GENERIC"] = "0"
            (pa, ra), (pb, rb) = res
            scale = max(1.0, pb.abs().max().item())
            assert maxabs(pa, pb) <= 2e-5 * scale, (shape, bn, x0.dtype, need_raw, maxabs(pa, pb))
            if need_raw:
                assert maxabs(ra, rb) <= 1e-6 * max(1.0, rb.abs().max().item()), (shape, bn)


@pytest.mark.parametrize("cot", ["mean", "ramp"])
def test_track_stages_mode_matches_reference_stages_and_stage_gradients(cot):
    """track_stages=True (model.track_images, reference model.py:229-254): every stage tensor, its .grad, the output
    (incl. the YUV->RGB->YUV round trip of pipeline_torch.py:197-200) and all gradients against the golden fixture."""
    from processing.pipeline_torch import ParametrizedProcessing
    c = GoldenCase("stages_pert")
    mod = ParametrizedProcessing(track_stages=True, batch_norm_output=False)
    mod.load_state_dict(c.state, strict=True)
    mod = mod.cuda()
    x = c.raw.cuda().requires_grad_(True)
    out = mod(x)
    assert list(mod.stages) == ["demosaic", "color_correct", "sharpening", "gaussian", "clipped", "gamma_correct"]
    assert mod.buffer["processed_rgb"] is out
    out.backward(isp_oracle.cotangent(tuple(out.shape), cot).cuda())
    assert maxabs(out.detach().cpu().numpy(), c.f32["out"]) <= FWD_ATOL
    for name, t in mod.stages.items():
        assert maxabs(t.detach().cpu().numpy(), c.f32[f"stage.{name}"]) <= FWD_ATOL, name
        ref = c.f64[f"stagegrad.{cot}.{name}"]
        assert t.grad is not None, name
        assert maxabs(t.grad.cpu().numpy(), ref) <= GRAD_ATOL * max(1.0, float(np.abs(ref).max())), name
    named = dict(mod.named_parameters())
    for k in isp_oracle.PARAM_KEYS:
        assert maxabs(named[k].grad.cpu().numpy(), c.f64[f"grad.{cot}.{k}"]) <= GRAD_ATOL, k
    ref = c.f64[f"grad.{cot}.raw"]
    assert maxabs(x.grad.cpu().numpy(), ref) <= GRAD_ATOL * max(1.0, float(np.abs(ref).max()))
    # without a raw gradient the stages are populated but nothing is retained (pipeline_torch.py:219-221)
    out2 = mod(c.raw.cuda())
    assert len(mod.stages) == 6 and out2.shape == out.shape


@pytest.mark.parametrize("shape", [(5, 256, 256), (2, 96, 200), (3, 72, 136)])
@pytest.mark.parametrize("tail", ["none", "bn_train", "bn_eval", "additive"])
def test_backward_from_saved_output_agrees_with_full_recompute(shape, tail):
    """Default backward (clip mask / gamma derivative read off the saved forward output) against the variant that
    recomputes everything from raw (R2L_ISP_RECOMPUTE=1), with and without the BatchNorm / additive tails."""
    import os
    from processing.pipeline_torch import ParametrizedProcessing
    state = syn.perturbed_state(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))
    raw = syn.smooth_scene(*shape, "drone", seed=41).cuda()
    g = isp_oracle.cotangent((shape[0], 3, shape[1], shape[2]), "ramp").cuda()
    res = []
    for mode in ("0", "1"):
        os.environ["R2L_ISP_RECOMPUTE"] = mode
        try:
            bn = tail.startswith("bn")
            mod = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], batch_norm_output=bn)
            mod.load_state_dict(state, strict=not bn)
            if tail == "additive":
                torch.manual_seed(3)
                mod.additive_layer = torch.nn.Parameter(0.01 * torch.randn(1, 3, shape[1], shape[2]))
            mod = mod.cuda().train(tail != "bn_eval")
            x = raw.clone().requires_grad_(True)
            mod(x).backward(g)
            flat = torch.cat([p.grad.flatten() for p in mod.parameters()]).cpu()
            res.append((flat, x.grad.cpu()))
        finally:
            os.environ["R2L_ISP_RECOMPUTE"] = "0"
    (pa, ra), (pb, rb) = res
    assert maxabs(pa, pb) <= 2e-5 * max(1.0, pb.abs().max().item()), (shape, tail, maxabs(pa, pb))
    assert maxabs(ra, rb) <= 2e-5 * max(1.0, rb.abs().max().item()), (shape, tail, maxabs(ra, rb))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(5, 256, 256), (2, 96, 200), (3, 72, 136), (1, 8, 8), (2, 37, 8)])
@pytest.mark.parametrize("tail", ["none", "bn_train", "additive"])
@pytest.mark.parametrize("u16", [False, True])
def test_tmem_backward_agrees_with_register_backward(shape, tail, u16):
    """Default backward (saved output + luma planes) against the third generation (no saved luma planes,
    R2L_ISP_NO_LUMA=1: Y0 / Y1 rebuilt per tile) on multi-tile, partial-tile, odd-batch and tiny shapes, float and uint16
    raw, with and without tails; and bit-reproducible."""
    import os
    from processing.pipeline_torch import ParametrizedProcessing
    state = syn.perturbed_state(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))
    raw = syn.smooth_scene(*shape, "drone", seed=43)
    raw = (syn.to_uint16(raw) if u16 else raw).cuda()
    g = isp_oracle.cotangent((shape[0], 3, shape[1], shape[2]), "ramp").cuda()
    res = []
    for env in ({}, {}, {"R2L_ISP_NO_LUMA": "1"}):
        os.environ.update(env)
        try:
            bn = tail.startswith("bn")
            mod = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], batch_norm_output=bn)
            mod.load_state_dict(state, strict=not bn)
            if tail == "additive":
                torch.manual_seed(3)
                mod.additive_layer = torch.nn.Parameter(0.01 * torch.randn(1, 3, shape[1], shape[2]))
            mod = mod.cuda().train()
            x = raw.clone() if u16 else raw.clone().requires_grad_(True)
            mod(x).backward(g)
            flat = torch.cat([p.grad.flatten() for p in mod.parameters()]).cpu()
            res.append((flat, None if u16 else x.grad.cpu()))
        finally:
            for k in env:
                os.environ.pop(k, None)
    (p5, r5), (p5b, r5b), (p3, r3) = res
    assert torch.equal(p5, p5b) and (u16 or torch.equal(r5, r5b)), "TMEM backward is not reproducible run to run"
    for name, p, r in (("gen3", p3, r3),):
        assert maxabs(p5, p) <= 2e-5 * max(1.0, p.abs().max().item()), (name, shape, tail, maxabs(p5, p))
        if not u16:
            assert maxabs(r5, r) <= 2e-5 * max(1.0, r.abs().max().item()), (name, shape, tail, maxabs(r5, r))


@pytest.mark.parametrize("shape", [(6, 256, 256), (3, 72, 136)])
def test_deferred_batchnorm_tail_equals_the_separately_finished_one(shape):
    """Train-mode BatchNorm backward: the tail finished in the backward kernel's prologue from the statistics kernel's
    per-CTA sums (default) against the one a separate finish kernel completes (R2L_ISP_BN_TAIL_COMPLETE=1): the same
    sums in the same order, so every gradient is bit-identical."""
    import os
    from processing.pipeline_torch import ParametrizedProcessing
    state = syn.perturbed_state(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))
    raw = syn.smooth_scene(*shape, "drone", seed=43).cuda()
    g = isp_oracle.cotangent((shape[0], 3, shape[1], shape[2]), "ramp").cuda()
    res = []
    for mode in ("0", "1"):
        os.environ["R2L_ISP_BN_TAIL_COMPLETE"] = mode
        try:
            mod = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], batch_norm_output=True)
            mod.load_state_dict(state, strict=False)
            mod = mod.cuda().train()
            x = raw.clone().requires_grad_(True)
            mod(x).backward(g)
            res.append([x.grad.cpu().numpy()] + [p.grad.cpu().numpy() for p in mod.parameters()])
        finally:
            os.environ.pop("R2L_ISP_BN_TAIL_COMPLETE", None)
    for a, b in zip(*res):
        assert np.isfinite(a).all() and np.array_equal(a, b)


def test_num_batches_tracked_is_advanced_by_the_forward_kernel():
    """nn.BatchNorm2d's counter: one per train-mode forward (eager and CUDA-graph replay alike), none in eval mode; the
    cumulative-average mode (momentum=None) keeps torch's semantics through the host path."""
    from processing.pipeline_torch import ParametrizedProcessing
    from raw2logit_b200.graphs import GraphedStep
    mod = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], batch_norm_output=True).cuda().train()
    raw = syn.smooth_scene(4, 64, 96, "drone", seed=5).cuda()
    ref_bn = torch.nn.BatchNorm2d(3, affine=False).cuda().train()
    plain = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], batch_norm_output=False).cuda()
    for _ in range(3):
        mod(raw)
        ref_bn(plain(raw))
    assert int(mod.batch_norm.num_batches_tracked) == 3 == int(ref_bn.num_batches_tracked)
    assert torch.allclose(mod.batch_norm.running_mean, ref_bn.running_mean, atol=1e-6)
    assert torch.allclose(mod.batch_norm.running_var, ref_bn.running_var, atol=1e-6)
    mod.eval()
    mod(raw)
    assert int(mod.batch_norm.num_batches_tracked) == 3
    mod.train()
    step = GraphedStep(mod, raw, torch.full((4, 3, 64, 96), 1e-3, device="cuda"))
    n0 = int(mod.batch_norm.num_batches_tracked)
    step.replay()
    step.replay()
    torch.cuda.synchronize()
    assert int(mod.batch_norm.num_batches_tracked) == n0 + 2
    mod.batch_norm.momentum, ref_bn.momentum = None, None           # cumulative moving average
    n1 = int(mod.batch_norm.num_batches_tracked)
    ref_bn.num_batches_tracked.fill_(n1)
    ref_bn.running_mean.copy_(mod.batch_norm.running_mean)
    ref_bn.running_var.copy_(mod.batch_norm.running_var)
    mod(raw)
    ref_bn(plain(raw))
    assert int(mod.batch_norm.num_batches_tracked) == n1 + 1
    assert torch.allclose(mod.batch_norm.running_mean, ref_bn.running_mean, atol=1e-6)


def test_graphed_step_after_an_eager_step_whose_graph_is_still_alive():
    """GraphedStep must capture even when an earlier eager forward / backward on the default stream is still referenced
    (its gradient accumulators live on that stream): the captured step differentiates parameter aliases, and its
    gradients equal the eager ones bit for bit; the module's own .grad tensors are not touched."""
    from processing.pipeline_torch import ParametrizedProcessing
    from raw2logit_b200.graphs import GraphedStep
    mod = ParametrizedProcessing(syn.CAMERA_PRESETS["drone"], batch_norm_output=False).cuda()
    raw = syn.smooth_scene(4, 64, 96, "drone", seed=9).cuda()
    g = isp_oracle.cotangent((4, 3, 64, 96), "ramp").cuda()
    out = mod(raw)
    out.backward(g, retain_graph=True)                       # `out` (and its graph) stay alive below
    eager = torch.cat([p.grad.reshape(-1) for p in mod.parameters()]).clone()
    kept = [p.grad for p in mod.parameters()]
    step = GraphedStep(mod, raw, g)
    step.replay()
    torch.cuda.synchronize()
    assert torch.equal(step.flat_grads, eager)
    assert torch.equal(step.out, out)
    assert all(p.grad is k for p, k in zip(mod.parameters(), kept))
