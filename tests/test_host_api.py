"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the module surface matches the
reference's (names, state_dict keys, pickling), and nothing silently falls back to the CPU."""
import copy
import os
import pickle
import re

import pytest
import torch

from raw2logit_b200 import _build, _lib
from tests.conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _lib.load()


def test_library_exports_every_symbol_the_header_declares(lib):
    header = open(os.path.join(ROOT, "include", "r2l_isp.h")).read()
    declared = set(re.findall(r"\b(r2l_isp_\w+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.r2l_isp_abi_version() == _lib.ABI_VERSION == 6
    assert b"shape" in lib.r2l_isp_error_string(-1)
    assert lib.r2l_isp_workspace_bytes(64, 256, 256) >= 155 * 4


def test_argument_validation_happens_before_any_cuda_call(lib):
    import ctypes
    p = _lib.IspParams(*([1] * 9))
    assert lib.r2l_isp_forward(None, 7, 65535.0, 1, 8, 8, ctypes.byref(p), None, None, None, None) == -2   # dtype
    assert lib.r2l_isp_forward(None, 0, 65535.0, 1, 2, 8, ctypes.byref(p), None, None, None, None) == -1   # H < 3
    assert lib.r2l_isp_forward(None, 0, 65535.0, 1, 8, 8, ctypes.byref(p), None, None, None, None) == -3   # null raw
    # saved_luma is written with 256-bit stores: a pointer that is only 16-byte aligned is refused (include/r2l_isp.h)
    vp = ctypes.c_void_p
    raw, out = vp(0x10000), vp(0x20000)
    assert lib.r2l_isp_forward(raw, 0, 65535.0, 2, 64, 64, ctypes.byref(p), None, out, vp(0x30010), None) == -7
    assert lib.r2l_isp_luma_supported(raw, 0, 2, 64, 64, out, None) == 1
    assert lib.r2l_isp_luma_supported(raw, 0, 2, 64, 62, out, None) == 0                             # W % 4 != 0
    assert lib.r2l_isp_mosaic(None, 0, 1.0, 1, 7, 8, None, 1, 3, None, None) == -1                   # odd + packed
    assert lib.r2l_isp_mosaic(None, 0, 1.0, 1, 8, 8, None, 1, 5, None, None) == -7                   # channels


def test_module_surface_matches_the_reference():
    from processing import pipeline_torch as pt
    for name in ["ParametrizedProcessing", "RawToRGB", "NNProcessing", "Debayer", "raw2rgb", "append_additive_layer",
                 "K_G", "K_RB", "K_BLUR", "K_SHARP", "M_RGB_2_YUV", "M_YUV_2_RGB", "DEFAULT_CAMERA_PARAMS"]:
        assert hasattr(pt, name), name
    m = pt.ParametrizedProcessing()
    assert list(m.state_dict().keys()) == [
        "black_level", "white_balance", "colour_correction", "gamma_correct", "M_RGB_2_YUV", "M_YUV_2_RGB",
        "debayer.weight", "sharpening_filter.weight", "gaussian_blur.weight", "batch_norm.running_mean",
        "batch_norm.running_var", "batch_norm.num_batches_tracked"]
    assert [n for n, _ in m.named_parameters()] == [
        "black_level", "white_balance", "colour_correction", "gamma_correct", "debayer.weight",
        "sharpening_filter.weight", "gaussian_blur.weight"]
    assert sum(p.numel() for p in m.parameters()) == 132
    assert pt.ParametrizedProcessing(batch_norm_output=False).batch_norm is None
    assert m.additive_layer is None and m.stages is None and m.buffer is None
    pt.append_additive_layer(m)
    assert m.additive_layer.shape == (1, 3, 256, 256)
    # adv_parameters selects a group by substring of the parameter name (model.py:70-75)
    assert [n for n, _ in m.named_parameters() if "gaussian_blur" in n] == ["gaussian_blur.weight"]


def test_state_dict_interop_with_the_oracle_layout_and_pickling():
    from oracle import isp_oracle
    from processing.pipeline_torch import ParametrizedProcessing
    from raw2logit_b200 import synthetic as syn
    cam = syn.CAMERA_PRESETS["microscopy"]
    m = ParametrizedProcessing(cam, batch_norm_output=False)
    st = isp_oracle.default_state(cam)
    for k, v in m.state_dict().items():
        assert torch.equal(v, st[k]), k
    m.load_state_dict(syn.perturbed_state(st), strict=True)
    m2 = pickle.loads(pickle.dumps(copy.deepcopy(m)))
    assert type(m2).__module__ == "processing.pipeline_torch"
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k


@pytest.mark.skipif(not os.path.exists("/root/reference/processing/pipeline_torch.py"), reason="reference not mounted")
def test_reference_state_dict_loads_strictly():
    from oracle import ref_loader
    from processing.pipeline_torch import ParametrizedProcessing
    ref = ref_loader.load_reference()
    theirs = ref.ParametrizedProcessing()
    mine = ParametrizedProcessing()
    mine.load_state_dict(theirs.state_dict(), strict=True)
    theirs.load_state_dict(mine.state_dict(), strict=True)
    assert [n for n, _ in theirs.named_parameters()] == [n for n, _ in mine.named_parameters()]


def test_cpu_input_raises_instead_of_falling_back():
    from processing.pipeline_torch import ParametrizedProcessing, raw2rgb
    with pytest.raises(RuntimeError, match="CUDA-only"):
        ParametrizedProcessing(batch_norm_output=False)(torch.rand(1, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        raw2rgb(torch.rand(1, 8, 8))
    with pytest.raises(NotImplementedError):
        st = ParametrizedProcessing(batch_norm_output=False)
        torch.ops.raw2logit_isp.forward(torch.rand(1, 8, 8), st.black_level, st.white_balance, st.colour_correction,
                                        st.gamma_correct, st.debayer.weight, st.sharpening_filter.weight,
                                        st.gaussian_blur.weight, st.M_RGB_2_YUV, st.M_YUV_2_RGB, None, None, 65535.0)


def test_product_code_never_imports_the_oracle_or_the_emulation():
    pkg = os.path.join(ROOT, "raw2logit_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "tests.emu" not in src and "isp_emu" not in src, f
