"""Logic check of the CUDA CTA code on the CPU (no GPU in the build container).

tests/emu compiles raw2logit_b200/csrc/isp_core.cuh -- the source the sm_100a kernels are made of -- with g++ and
runs the threads of each CTA sequentially.  Here its results are compared with the reference's own outputs
(tests/golden): indexing, border rules, the collapsed demosaic->YUV tables, the hand-derived adjoints and the
statistics -> 132-gradient finish are all exercised.  The emulation is never used by the product.
"""
import numpy as np
import pytest
import torch

from oracle import isp_oracle
from raw2logit_b200 import synthetic as syn
from tests.emu import emu
from tests.golden_util import GoldenCase, case_names, maxabs

FUSED_CASES = [n for n in case_names() if not GoldenCase(n).track_stages and GoldenCase(n).bn is None]
# inputs whose pre-clip values sit on the clip thresholds (d/dx x^(1/2.2) = 241 at 1e-5): gate vs the fp64 truth
ILL_CONDITIONED = {"noise_g2_pert", "impulses", "car_crop"}


@pytest.mark.parametrize("name", FUSED_CASES)
def test_emulated_forward_matches_reference(name):
    c = GoldenCase(name)
    add = None if c.additive is None else c.additive.numpy()[0]
    out = emu.forward(c.raw.numpy(), c.state, additive=add)
    assert not np.isnan(out).any()
    err32 = maxabs(out, c.f32["out"])
    assert err32 <= 1e-5, (name, err32)                       # north-star forward tolerance
    if name in ILL_CONDITIONED:
        floor = maxabs(c.f32["out"], c.f64["out"])
        assert maxabs(out, c.f64["out"]) <= max(2 * floor, 6e-6), name
    else:
        assert err32 <= 2e-6, (name, err32)


@pytest.mark.parametrize("cot", ["mean", "ramp"])
@pytest.mark.parametrize("name", FUSED_CASES)
def test_emulated_backward_matches_reference(name, cot):
    c = GoldenCase(name)
    g = isp_oracle.cotangent(tuple(c.f32["out"].shape), cot).numpy()
    grads = emu.backward(c.raw.numpy(), c.state, g)
    for k, v in grads.items():
        ref64 = c.f64[f"grad.{cot}.{k}"]
        ref32 = c.f32[f"grad.{cot}.{k}"]
        assert not np.isnan(v).any(), (name, k)
        err = maxabs(v.reshape(ref64.shape), ref64)
        floor = maxabs(ref32, ref64)
        assert err <= 1e-4, (name, k, err)                    # north-star parameter-gradient tolerance
        assert err <= max(2 * floor, 2e-6 * max(1.0, float(np.abs(ref64).max()))) or name in ILL_CONDITIONED, \
            (name, k, err, floor)


def test_emulated_backward_without_raw_grad_gives_same_parameter_grads():
    c = GoldenCase("drone_g1_pert")
    g = isp_oracle.cotangent(tuple(c.f32["out"].shape), "ramp").numpy()
    a = emu.backward(c.raw.numpy(), c.state, g, need_raw_grad=True)
    b = emu.backward(c.raw.numpy(), c.state, g, need_raw_grad=False)
    for k in b:
        assert maxabs(a[k], b[k]) <= 1e-6, k


@pytest.mark.parametrize("n_cta", [1, 2, 7])
def test_tile_to_cta_assignment_does_not_change_results(n_cta):
    raw = syn.smooth_scene(3, 70, 130, "drone", seed=11)            # 3 x 3 tiles per image incl. partial tiles
    st = syn.perturbed_state(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))
    want, grads = isp_oracle.forward_backward(raw, st, grad_out="ramp", dtype=torch.float64)
    out = emu.forward(raw.numpy(), st, n_cta=n_cta)
    assert maxabs(out, want) <= 2e-6
    g = isp_oracle.cotangent(tuple(out.shape), "ramp").numpy()
    got = emu.backward(raw.numpy(), st, g, n_cta=n_cta)
    for k, v in got.items():
        ref = grads[k].numpy()
        assert maxabs(v.reshape(ref.shape), ref) <= 2e-6 * max(1.0, float(np.abs(ref).max())), k


def test_uint16_ingest_equals_float_of_quotient():
    raw = syn.smooth_scene(2, 34, 66, "drone", seed=5)
    u16 = syn.to_uint16(raw)
    st = isp_oracle.default_state(syn.CAMERA_PRESETS["drone"])
    as_float = (u16.to(torch.int32).to(torch.float32) / 65535.0)
    a = emu.forward(u16.numpy(), st)
    b = emu.forward(as_float.numpy(), st)
    assert np.array_equal(a, b)


def test_eval_batchnorm_tail_is_an_affine_epilogue():
    c = GoldenCase("bn_eval")
    rm, rv = c.extra["running_mean"].numpy(), c.extra["running_var"].numpy()
    scale = (1.0 / np.sqrt(rv.astype(np.float64) + 1e-5)).astype(np.float32)
    shift = (-rm.astype(np.float64) * scale).astype(np.float32)
    out = emu.forward(c.raw.numpy(), c.state, affine=np.concatenate([scale, shift]))
    assert maxabs(out, c.f32["out"]) <= 5e-5       # output is scaled by 1/sqrt(var) ~ 6: tolerance scales with it


@pytest.mark.parametrize("name", ["bn_train", "bn_train_additive"])
@pytest.mark.parametrize("cot", ["mean", "ramp"])
def test_train_batchnorm_tail_backward(name, cot):
    """dL/do = gs*(G - mean(G) - yhat*mean(G*yhat)) formed inside the backward kernel (15-float grad_tail)."""
    c = GoldenCase(name)
    add = None if c.additive is None else c.additive
    o, _ = isp_oracle.forward(c.raw.double(), isp_oracle.cast_state(c.state, torch.float64),
                              additive=None if add is None else add.double(), dtype=torch.float64)
    mean = o.mean(dim=(0, 2, 3))
    var = o.var(dim=(0, 2, 3), unbiased=False)
    inv = 1.0 / torch.sqrt(var + 1e-5)
    y = (o - mean.view(1, 3, 1, 1)) * inv.view(1, 3, 1, 1)
    assert maxabs(y, c.f64["out"]) <= 1e-9
    g = isp_oracle.cotangent(tuple(y.shape), cot, torch.float64)
    c1 = g.mean(dim=(0, 2, 3))
    c2 = (g * y).mean(dim=(0, 2, 3))
    tail = torch.cat([inv, c1, c2, inv, -mean * inv]).float().numpy()
    grads = emu.backward(c.raw.numpy(), c.state, g.float().numpy(), grad_tail=tail,
                         additive=None if add is None else add.numpy()[0])
    for k, v in grads.items():
        ref = c.f64[f"grad.{cot}.{k}"]
        assert maxabs(v.reshape(ref.shape), ref) <= 2e-5 * max(1.0, float(np.abs(ref).max())), (name, k)


def test_forward_v1_and_v2_agree_and_channel_statistics():
    raw = syn.smooth_scene(3, 70, 130, "drone", seed=12)
    st = syn.perturbed_state(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))
    v1 = emu.forward(raw.numpy(), st, version=1)
    sums = np.zeros(6, dtype=np.float64)
    v2 = emu.forward(raw.numpy(), st, version=2, chan_sums=sums)
    assert maxabs(v1, v2) <= 1e-6
    o = v2.astype(np.float64)
    want = np.concatenate([o.sum(axis=(0, 2, 3)), (o * o).sum(axis=(0, 2, 3))])
    assert np.allclose(sums, want, rtol=1e-5)


@pytest.mark.parametrize("shape,n_cta,need_raw", [((3, 72, 136), 3, True), ((2, 64, 128), 3, True), ((1, 40, 72), 2, True),
                                                  ((3, 96, 200), 5, True), ((2, 8, 8), 1, True), ((2, 37, 8), 3, True),
                                                  ((2, 64, 128), 3, False), ((1, 101, 72), 4, True),
                                                  ((2, 68, 132), 3, True), ((1, 33, 68), 2, True), ((1, 35, 196), 3, True)])
def test_vectorised_backward_multi_tile_shapes(shape, n_cta, need_raw):
    """Third-generation backward (padded-domain phases + fold passes) on multi-tile / partial-tile / odd-batch shapes
    against the fp64 oracle."""
    raw = syn.smooth_scene(*shape, "drone", seed=sum(shape))
    st = syn.perturbed_state(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))
    want, grads = isp_oracle.forward_backward(raw, st, grad_out="ramp", dtype=torch.float64)
    g = isp_oracle.cotangent(tuple(want.shape), "ramp").numpy()
    got = emu.backward(raw.numpy(), st, g, n_cta=n_cta, version=3, need_raw_grad=need_raw)
    for k, v in got.items():
        ref = grads[k].numpy()
        assert maxabs(v.reshape(ref.shape), ref) <= 5e-6 * max(1.0, float(np.abs(ref).max())), (k, shape)


@pytest.mark.parametrize("shape,n_cta", [((3, 72, 136), 3), ((2, 64, 128), 2), ((1, 40, 72), 2), ((3, 96, 200), 5),
                                         ((2, 8, 8), 1), ((2, 37, 8), 3), ((1, 101, 72), 4), ((1, 5, 12), 1)])
def test_forward_v3_multi_tile_shapes(shape, n_cta):
    """Third-generation forward (border rules on the data, no fix-up passes) against the fp64 oracle and against the
    second generation, with the additive layer + affine epilogue and the channel sums of the BatchNorm tail."""
    raw = syn.smooth_scene(*shape, "drone", seed=sum(shape))
    st = syn.perturbed_state(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))
    want, _ = isp_oracle.forward(raw.double(), isp_oracle.cast_state(st, torch.float64), dtype=torch.float64)
    v3 = emu.forward(raw.numpy(), st, n_cta=n_cta, version=3)
    assert maxabs(v3, want) <= 2e-6
    v2 = emu.forward(raw.numpy(), st, n_cta=n_cta, version=2)
    assert maxabs(v3, v2) <= 1e-6
    g = torch.Generator().manual_seed(1)
    add = (0.01 * torch.randn(3, shape[1], shape[2], generator=g)).numpy()
    aff = np.asarray([1.5, 0.5, 2.0, -0.1, 0.2, 0.3], dtype=np.float32)
    t3 = emu.forward(raw.numpy(), st, additive=add, affine=aff, n_cta=n_cta, version=3)
    t2 = emu.forward(raw.numpy(), st, additive=add, affine=aff, n_cta=n_cta, version=2)
    assert maxabs(t3, t2) <= 2e-6
    s3, s2 = np.zeros(6), np.zeros(6)
    o3 = emu.forward(raw.numpy(), st, additive=add, n_cta=n_cta, version=3, chan_sums=s3)
    emu.forward(raw.numpy(), st, additive=add, n_cta=n_cta, version=2, chan_sums=s2)
    o = o3.astype(np.float64)
    assert np.allclose(s3, np.concatenate([o.sum(axis=(0, 2, 3)), (o * o).sum(axis=(0, 2, 3))]), rtol=1e-5)
    assert np.allclose(s3, s2, rtol=1e-5)


@pytest.mark.parametrize("shape,n_cta,need_raw", [((3, 72, 136), 3, True), ((2, 64, 128), 3, False), ((1, 40, 72), 2, True),
                                                  ((2, 8, 8), 1, True), ((1, 101, 72), 4, True), ((2, 68, 132), 3, True),
                                                  ((1, 33, 68), 2, True)])
def test_backward_from_forward_output_matches_oracle(shape, n_cta, need_raw):
    """Backward variant that is handed the forward output (no Gaussian / colour-tail recompute): clip mask and gamma
    derivative are derived from the output; against the fp64 oracle and against the full-recompute kernel."""
    raw = syn.smooth_scene(*shape, "drone", seed=sum(shape) + 1)
    st = syn.perturbed_state(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))
    want, grads = isp_oracle.forward_backward(raw, st, grad_out="ramp", dtype=torch.float64)
    g = isp_oracle.cotangent(tuple(want.shape), "ramp").numpy()
    out = emu.forward(raw.numpy(), st, n_cta=n_cta, version=3)
    got = emu.backward(raw.numpy(), st, g, n_cta=n_cta, need_raw_grad=need_raw, out=out)
    ref = emu.backward(raw.numpy(), st, g, n_cta=n_cta, need_raw_grad=need_raw)
    for k, v in got.items():
        r64 = grads[k].numpy()
        scale = max(1.0, float(np.abs(r64).max()))
        assert maxabs(v.reshape(r64.shape), r64) <= 5e-6 * scale, (k, shape)
        assert maxabs(v, ref[k]) <= 5e-6 * scale, (k, shape)


@pytest.mark.parametrize("name", ["bn_train", "bn_train_additive", "noise_g2_pert", "car_crop", "impulses"])
def test_backward_from_forward_output_golden_cases(name):
    """Same variant on the clip-heavy golden cases and behind the BatchNorm / additive tail (o is recovered from the
    normalised output by the affine inverse, clip thresholds carry a small margin there)."""
    c = GoldenCase(name)
    cot = "ramp"
    if c.bn is None:
        g = isp_oracle.cotangent(tuple(c.f32["out"].shape), cot).numpy()
        add = None if c.additive is None else c.additive.numpy()[0]
        out = emu.forward(c.raw.numpy(), c.state, additive=add)
        got = emu.backward(c.raw.numpy(), c.state, g, out=out)
        ref = emu.backward(c.raw.numpy(), c.state, g)
    else:
        add = None if c.additive is None else c.additive
        o, _ = isp_oracle.forward(c.raw.double(), isp_oracle.cast_state(c.state, torch.float64),
                                  additive=None if add is None else add.double(), dtype=torch.float64)
        mean, var = o.mean(dim=(0, 2, 3)), o.var(dim=(0, 2, 3), unbiased=False)
        inv = 1.0 / torch.sqrt(var + 1e-5)
        y = (o - mean.view(1, 3, 1, 1)) * inv.view(1, 3, 1, 1)
        gt = isp_oracle.cotangent(tuple(y.shape), cot, torch.float64)
        tail = torch.cat([inv, gt.mean(dim=(0, 2, 3)), (gt * y).mean(dim=(0, 2, 3)), inv, -mean * inv]).float().numpy()
        addn = None if add is None else add.numpy()[0]
        g = gt.float().numpy()
        got = emu.backward(c.raw.numpy(), c.state, g, grad_tail=tail, additive=addn, out=y.float().numpy())
        ref = emu.backward(c.raw.numpy(), c.state, g, grad_tail=tail, additive=addn)
    for k, v in got.items():
        r64 = c.f64[f"grad.{cot}.{k}"]
        scale = max(1.0, float(np.abs(r64).max()))
        assert maxabs(v.reshape(r64.shape), r64) <= 1e-4, (name, k)
        assert maxabs(v, ref[k]) <= 2e-5 * scale, (name, k, maxabs(v, ref[k]))


# ---- fourth generation: the forward keeps Y0 / Y1, the backward recomputes nothing ------------------------------
def _luma_oracle(raw, st):
    """Y0 (after RGB->YUV, :194) and Y1 (after the sharpening filter, :195) in fp64, straight from the oracle's ops."""
    import torch.nn.functional as F
    s64 = isp_oracle.cast_state(st, torch.float64)
    m = isp_oracle.mosaic(raw.double(), s64["black_level"], reduce_size=False, dtype=torch.float64)
    d = F.conv2d(F.pad(m, (1, 1, 1, 1), mode="reflect"), s64["debayer.weight"])
    c = isp_oracle._mix(d * s64["white_balance"].reshape(1, 3, 1, 1), s64["colour_correction"])
    y0 = isp_oracle._mix(c, s64["M_RGB_2_YUV"])[:, 0]
    y1 = F.conv2d(y0[:, None], s64["sharpening_filter.weight"], padding=1)[:, 0]
    return y0.numpy(), y1.numpy()


def _luma_alloc(b, h, w):
    return np.full((2, (b + 1) // 2, h, w, 2), np.nan, dtype=np.float32)


V4_SHAPES = [((3, 72, 136), 3, True), ((2, 64, 128), 3, False), ((1, 40, 72), 2, True), ((2, 8, 8), 1, True),
             ((1, 101, 72), 4, True), ((2, 68, 132), 3, True), ((1, 33, 68), 2, True), ((2, 37, 8), 3, True),
             ((3, 96, 200), 5, True)]


@pytest.mark.parametrize("gen", [4, 5])
@pytest.mark.parametrize("shape,n_cta,need_raw", V4_SHAPES)
def test_saved_luma_planes_and_v4_backward(shape, n_cta, need_raw, gen):
    """The forward's saved luma planes equal the oracle's Y0 / Y1 (pair-interleaved layout, odd batches duplicate the
    last image), and the fourth / fifth-generation backward fed with them matches the fp64 oracle and the third
    generation."""
    raw = syn.smooth_scene(*shape, "drone", seed=sum(shape) + 2)
    st = syn.perturbed_state(isp_oracle.default_state(syn.CAMERA_PRESETS["drone"]))
    b, h, w = shape
    luma = _luma_alloc(b, h, w)
    out = emu.forward(raw.numpy(), st, n_cta=n_cta, version=3, luma=luma)
    assert not np.isnan(luma).any()
    y0, y1 = _luma_oracle(raw, st)
    for img in range(b):
        assert maxabs(luma[0, img // 2, :, :, img % 2], y0[img]) <= 2e-6
        assert maxabs(luma[1, img // 2, :, :, img % 2], y1[img]) <= 2e-6
    if b % 2:
        assert np.array_equal(luma[:, -1, :, :, 1], luma[:, -1, :, :, 0])
    want, grads = isp_oracle.forward_backward(raw, st, grad_out="ramp", dtype=torch.float64)
    g = isp_oracle.cotangent(tuple(want.shape), "ramp").numpy()
    got = emu.backward(raw.numpy(), st, g, n_cta=n_cta, need_raw_grad=need_raw, out=out, luma=luma, version=gen)
    ref = emu.backward(raw.numpy(), st, g, n_cta=n_cta, need_raw_grad=need_raw, out=out, version=3)
    for k, v in got.items():
        r64 = grads[k].numpy()
        scale = max(1.0, float(np.abs(r64).max()))
        assert maxabs(v.reshape(r64.shape), r64) <= 5e-6 * scale, (k, shape)
        assert maxabs(v, ref[k]) <= 5e-6 * scale, (k, shape)


@pytest.mark.parametrize("name", ["bn_train", "bn_train_additive", "noise_g2_pert", "car_crop", "impulses", "drone_g1_pert",
                                  "micro_g1_pert"])
@pytest.mark.parametrize("gen", [4, 5])
def test_v4_backward_golden_cases(name, gen):
    """Fourth / fifth generation on the golden cases (clip-heavy inputs, BatchNorm / additive tails, impulses at every CFA
    phase and border)."""
    c = GoldenCase(name)
    cot = "ramp"
    b, h, w = c.raw.shape
    if w % 4 or h < 8 or w < 8:
        pytest.skip("shape is served by the generic kernel")
    luma = _luma_alloc(b, h, w)
    if c.bn is None:
        g = isp_oracle.cotangent(tuple(c.f32["out"].shape), cot).numpy()
        add = None if c.additive is None else c.additive.numpy()[0]
        out = emu.forward(c.raw.numpy(), c.state, additive=add, luma=luma)
        tail = None
        if add is not None:
            tail = np.asarray([1, 1, 1, 0, 0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0], dtype=np.float32)
        got = emu.backward(c.raw.numpy(), c.state, g, out=out, luma=luma, version=gen, grad_tail=tail, additive=add)
    else:
        add = None if c.additive is None else c.additive
        emu.forward(c.raw.numpy(), c.state, additive=None if add is None else add.numpy()[0], luma=luma)
        o, _ = isp_oracle.forward(c.raw.double(), isp_oracle.cast_state(c.state, torch.float64),
                                  additive=None if add is None else add.double(), dtype=torch.float64)
        mean, var = o.mean(dim=(0, 2, 3)), o.var(dim=(0, 2, 3), unbiased=False)
        inv = 1.0 / torch.sqrt(var + 1e-5)
        y = (o - mean.view(1, 3, 1, 1)) * inv.view(1, 3, 1, 1)
        gt = isp_oracle.cotangent(tuple(y.shape), cot, torch.float64)
        tail = torch.cat([inv, gt.mean(dim=(0, 2, 3)), (gt * y).mean(dim=(0, 2, 3)), inv, -mean * inv]).float().numpy()
        addn = None if add is None else add.numpy()[0]
        got = emu.backward(c.raw.numpy(), c.state, gt.float().numpy(), grad_tail=tail, additive=addn,
                           out=y.float().numpy(), luma=luma, version=gen)
    for k, v in got.items():
        r64 = c.f64[f"grad.{cot}.{k}"]
        assert maxabs(v.reshape(r64.shape), r64) <= 1e-4, (name, k)
