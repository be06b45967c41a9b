"""CPU oracle for the differentiable ISP hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module; the product path (``raw2logit_b200``) never does and fails loudly without its CUDA library.

What it is: an independent restatement, in stock PyTorch CPU ops, of the reference's raw->RGB chain
(``/root/reference/processing/pipeline_torch.py``).  It is dtype-generic (float32 = the parity oracle the
tolerances in BASELINE.json are stated against; float64 = "truth" used to put the fp32 noise floor in context).
Gradients come from torch autograd over this restatement, exactly as the reference gets them.

Parity pinning: the reference holds no tests or golden vectors for this path, so the oracle is pinned
against outputs of the reference itself: ``oracle/make_golden.py`` imports the unmodified reference file in the
build container and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this restatement against
those vectors (and, when ``/root/reference`` is present, against the live reference).

Reference lines restated by each function are cited in the docstrings.
"""
import torch
import torch.nn.functional as F

# ---- constants (pipeline_torch.py:13-40).  Numeric calibration constants, necessarily identical. -------------
_RB = [[0.25, 0.5, 0.25], [0.5, 1.0, 0.5], [0.25, 0.5, 0.25]]                      # K_RB :17-19
_G = [[0.0, 0.25, 0.0], [0.25, 1.0, 0.25], [0.0, 0.25, 0.0]]                        # K_G  :13-15
_SHARP = [[0.0, -1.0, 0.0], [-1.0, 5.0, -1.0], [0.0, -1.0, 0.0]]                    # K_SHARP :33-35
# the reference prints K_BLUR to 5 significant digits (:28-32); the printed values are the parameters
_BLUR = [[6.9625e-08, 2.8089e-05, 2.0755e-04, 2.8089e-05, 6.9625e-08],
         [2.8089e-05, 1.1332e-02, 8.3731e-02, 1.1332e-02, 2.8089e-05],
         [2.0755e-04, 8.3731e-02, 6.1869e-01, 8.3731e-02, 2.0755e-04],
         [2.8089e-05, 1.1332e-02, 8.3731e-02, 1.1332e-02, 2.8089e-05],
         [6.9625e-08, 2.8089e-05, 2.0755e-04, 2.8089e-05, 6.9625e-08]]
_RGB2YUV = [[0.299, 0.587, 0.114],
            [-0.14714119, -0.28886916, 0.43601035],
            [0.61497538, -0.51496512, -0.10001026]]                                 # :21-23
_YUV2RGB = [[1.0000000000e+00, -4.1827794561e-09, 1.1398830414e+00],
            [1.0000000000e+00, -3.9464232326e-01, -5.8062183857e-01],
            [1.0000000000e+00, 2.0320618153e+00, -1.2232658220e-09]]                # :24-26
_DEFAULT_CAMERA = ([0.0] * 4, [1.0] * 3, [1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0])         # :36-40

PARAM_KEYS = ("black_level", "white_balance", "colour_correction", "gamma_correct",
              "debayer.weight", "sharpening_filter.weight", "gaussian_blur.weight")
BUFFER_KEYS = ("M_RGB_2_YUV", "M_YUV_2_RGB")


def default_state(camera_parameters=None):
    """Initial parameters/buffers, float32, keyed like the reference ``state_dict`` (pipeline_torch.py:143-173,
    Debayer init :232-237).  Built in fp32 exactly as the reference does (``K/4`` in fp32), cast later if needed."""
    bl, wb, ccm = camera_parameters if camera_parameters is not None else _DEFAULT_CAMERA
    f32 = dict(dtype=torch.float32)
    wd = torch.zeros(3, 3, 3, 3, **f32)
    wd[0, 0] = torch.tensor(_RB, **f32)
    wd[1, 1] = torch.tensor(_G, **f32)
    wd[2, 2] = torch.tensor(_RB, **f32)
    return {
        "black_level": torch.as_tensor(bl, **f32).clone(),
        "white_balance": torch.as_tensor(wb, **f32).reshape(1, 3).clone(),
        "colour_correction": torch.as_tensor(ccm, **f32).reshape(3, 3).clone(),
        "gamma_correct": torch.tensor([2.2], **f32),
        "M_RGB_2_YUV": torch.tensor(_RGB2YUV, **f32),
        "M_YUV_2_RGB": torch.tensor(_YUV2RGB, **f32),
        "debayer.weight": wd,
        "sharpening_filter.weight": torch.tensor(_SHARP, **f32).reshape(1, 1, 3, 3),
        "gaussian_blur.weight": torch.tensor(_BLUR, **f32).reshape(1, 1, 5, 5),
    }


def cast_state(state, dtype, requires_grad=False):
    """fp32 values carried exactly into ``dtype`` (so the fp32 and fp64 oracles use identical parameters)."""
    out = {}
    for k, v in state.items():
        t = v.detach().float().to(dtype).clone()
        if requires_grad and k in PARAM_KEYS:
            t.requires_grad_(True)
        out[k] = t
    return out


def mosaic(raw, black_level=None, reduce_size=True, out_channels=3, dtype=torch.float32):
    """CFA split with black-level subtraction -- ``raw2rgb`` (pipeline_torch.py:240-283).

    The output is allocated in ``dtype`` regardless of the input dtype (reference: ``torch.zeros`` default dtype,
    :261/:272).  RGGB: R=(even,even) G1=(even,odd) G2=(odd,even) B=(odd,odd) (:256-259).
    """
    assert out_channels in (3, 4)
    if black_level is None:
        black_level = [0, 0, 0, 0]
    n, h, w = raw.shape
    planes = [raw[:, 0::2, 0::2] - black_level[0], raw[:, 0::2, 1::2] - black_level[1],
              raw[:, 1::2, 0::2] - black_level[2], raw[:, 1::2, 1::2] - black_level[3]]
    if reduce_size:
        out = torch.zeros(n, out_channels, h // 2, w // 2, dtype=dtype, device=raw.device)
        if out_channels == 3:
            out[:, 0], out[:, 1], out[:, 2] = planes[0], (planes[1] + planes[2]) / 2, planes[3]
        else:
            for c in range(4):
                out[:, c] = planes[c]
        return out
    out = torch.zeros(n, out_channels, h, w, dtype=dtype, device=raw.device)
    slots = [(0, 0, 0), (1, 0, 1), (1 if out_channels == 3 else 2, 1, 0), (out_channels - 1, 1, 1)]
    for plane, (c, oy, ox) in zip(planes, slots):
        out[:, c, oy::2, ox::2] = plane
    return out


def _mix(x, m):
    """out[b,k] = sum_c x[b,c] * m[k,c]  (the 'bchw,kc->bkhw' contractions, pipeline_torch.py:191,194,203)."""
    return torch.einsum("bchw,kc->bkhw", x, m).contiguous()


def forward(raw, state, track_stages=False, additive=None, bn=None, dtype=torch.float32):
    """``ParametrizedProcessing.forward`` (pipeline_torch.py:175-225).  Returns ``(out, stages)``.

    bn: None, or dict(training=bool, running_mean=Tensor, running_var=Tensor, momentum=0.1, eps=1e-5) for the
    ``BatchNorm2d(3, affine=False)`` tail (:168, :216-217).  ``stages`` uses the reference's names/order.
    """
    assert raw.ndim == 3, f"needs dims (B, H, W), got {raw.shape}"      # :176
    stages = {}
    m = mosaic(raw, state["black_level"], reduce_size=False, dtype=dtype)          # :183
    stages["demosaic"] = m
    d = F.conv2d(F.pad(m, (1, 1, 1, 1), mode="reflect"), state["debayer.weight"])  # :187, Debayer :228-237
    w = d * state["white_balance"].reshape(1, 3, 1, 1)                             # :190 (sum over k of size 1)
    c = _mix(w, state["colour_correction"])                                        # :191
    stages["color_correct"] = c
    yuv = _mix(c, state["M_RGB_2_YUV"])                                            # :194
    y1 = F.conv2d(yuv[:, 0:1], state["sharpening_filter.weight"], padding=1)       # :195 zero padding
    yuv = torch.cat([y1, yuv[:, 1:]], dim=1)
    if track_stages:                                                               # :197-200
        rgb_s = _mix(yuv, state["M_YUV_2_RGB"])
        stages["sharpening"] = rgb_s
        yuv = _mix(rgb_s, state["M_RGB_2_YUV"])
    y2 = F.conv2d(F.pad(yuv[:, 0:1], (2, 2, 2, 2), mode="reflect"), state["gaussian_blur.weight"])  # :202
    yuv = torch.cat([y2, yuv[:, 1:]], dim=1)
    r = _mix(yuv, state["M_YUV_2_RGB"])                                            # :203
    stages["gaussian"] = r
    cl = torch.clamp(r, 1e-5, 1)                                                   # :206
    stages["clipped"] = cl
    o = torch.exp((1 / state["gamma_correct"]) * torch.log(cl))                    # :209
    stages["gamma_correct"] = o
    if additive is not None:                                                       # :212-214
        o = o + additive
        stages["noise"] = o
    if bn is not None:                                                             # :216-217
        o = F.batch_norm(o, bn["running_mean"], bn["running_var"], None, None, bn["training"],
                         bn.get("momentum", 0.1), bn.get("eps", 1e-5))
    return o, stages


def cotangent(shape, kind, dtype=torch.float32):
    """Gradient-test cotangents (SURVEY 7.3-2 iii): 'mean' = d(out.mean()), 'ramp' = fixed positive linspace."""
    n = 1
    for s in shape:
        n *= s
    if kind == "mean":
        return torch.full(shape, 1.0 / n, dtype=dtype)
    if kind == "ramp":
        return (torch.linspace(0.25, 1.75, n, dtype=torch.float64).reshape(shape) / n).to(dtype)
    raise ValueError(kind)


def forward_backward(raw, state, grad_out="mean", raw_grad=True, dtype=torch.float32, **fw):
    """Forward + autograd backward.  Returns ``(out, grads)``; grads keyed by PARAM_KEYS (+ 'raw', 'additive')."""
    st = cast_state(state, dtype, requires_grad=True)
    x = raw.to(dtype).clone().requires_grad_(raw_grad)
    add = fw.pop("additive", None)
    if add is not None:
        add = add.to(dtype).clone().requires_grad_(True)
    out, _ = forward(x, st, additive=add, dtype=dtype, **fw)
    g = cotangent(tuple(out.shape), grad_out, dtype) if isinstance(grad_out, str) else grad_out.to(dtype)
    out.backward(g)
    grads = {k: st[k].grad for k in PARAM_KEYS}
    if raw_grad:
        grads["raw"] = x.grad
    if add is not None:
        grads["additive"] = add.grad
    return out.detach(), grads
