"""The synthetic raw generator (raw2logit_b200/synthetic.py) that stands in for the reference's cloud-backed datasets
(dataset.py:24-41): the contract of a dataset item (dataset.py:87: 16-bit integers divided by 2**bits - 1), the camera
presets (dataset.py:209-213, :290-294; pipeline_torch.py:36-40), SURVEY 8d's generators G1, G2, G4.  CPU only."""
import ast
import os

import pytest
import torch

from raw2logit_b200 import synthetic as syn

REF_DATASET = os.path.join(os.environ.get("R2L_REF", "/root/reference"), "dataset.py")


def _class_constants(path, cls):
    """black_level / white_balance / colour_matrix literals of a dataset class, read without importing the file (its
    import chain needs packages that are absent here)."""
    tree = ast.parse(open(path).read())
    out = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for st in node.body:
                if isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Name) and \
                        st.targets[0].id in ("black_level", "white_balance", "colour_matrix"):
                    out[st.targets[0].id] = ast.literal_eval(st.value)
    return out


@pytest.mark.skipif(not os.path.exists(REF_DATASET), reason="reference tree not present")
@pytest.mark.parametrize("preset,cls", [("drone", "DroneDatasetSegmentationFull"),
                                        ("microscopy", "MicroscopyDataset")])
def test_presets_are_the_reference_calibration_constants(preset, cls):
    ref = _class_constants(REF_DATASET, cls)                          # dataset.py:209-213 / :290-294
    bl, wb, ccm = syn.CAMERA_PRESETS[preset]
    assert list(bl) == list(ref["black_level"])
    assert list(wb) == list(ref["white_balance"])
    assert list(ccm) == list(ref["colour_matrix"])


def test_default_preset_is_the_identity_camera():
    bl, wb, ccm = syn.CAMERA_PRESETS["default"]                     # pipeline_torch.py:36-40
    assert bl == [0.0] * 4 and wb == [1.0] * 3 and ccm == [1., 0., 0., 0., 1., 0., 0., 0., 1.]


@pytest.mark.parametrize("preset", ["default", "drone", "microscopy"])
def test_smooth_scene_is_seeded_quantised_and_in_range(preset):
    a = syn.smooth_scene(3, 34, 50, preset, seed=5)
    b = syn.smooth_scene(3, 34, 50, preset, seed=5)
    c = syn.smooth_scene(3, 34, 50, preset, seed=6)
    assert a.dtype == torch.float32 and a.shape == (3, 34, 50) and a.is_contiguous()
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert a.min() >= 0.0 and a.max() <= 1.0
    # every value is k / 65535 for an integer k, computed the way dataset.py:87 does (integer / float divide)
    k = torch.round(a.double() * 65535.0)
    assert torch.equal((k / 65535.0).float(), a)
    # the uint16 view carries exactly those integers, and dividing them reproduces the float mosaic bit for bit
    u = syn.to_uint16(a)
    assert u.dtype == torch.uint16 and torch.equal(u.to(torch.int32).double(), k)
    assert torch.equal((u.to(torch.int32).float() / 65535.0), a)


def test_smooth_scene_is_grey_world_through_the_preset():
    """White-balancing the black-level-corrected mosaic gives the same scene at all four CFA phases."""
    bl, wb, _ = syn.CAMERA_PRESETS["drone"]
    raw = syn.smooth_scene(2, 64, 64, "drone", seed=3, noise=0.0)
    gain = [abs(wb[0]), wb[1], wb[1], abs(wb[2])]
    planes = [(raw[:, py::2, px::2] - bl[2 * py + px]) * gain[2 * py + px] for py in (0, 1) for px in (0, 1)]
    for p in planes:
        assert 0.14 <= p.min() and p.max() <= 0.51                  # scene = 0.15 + 0.35 * [0, 1]
    # neighbouring phases sample a smooth scene one site apart
    assert (planes[0] - planes[3]).abs().max() < 0.03


def test_noise_stress_is_uniform_and_seeded():
    a, b = syn.noise_stress(2, 32, 32, seed=0), syn.noise_stress(2, 32, 32, seed=0)
    assert torch.equal(a, b) and 0.45 < a.mean() < 0.55 and a.min() >= 0 and a.max() <= 1


def test_impulses_hit_every_cfa_phase_corner_and_edge():
    h, w = 12, 16
    pos = syn.impulse_positions(h, w)
    raw = syn.impulses(h, w, pos, value=0.5)
    assert raw.shape == (len(pos), h, w)
    for n, (y, x) in enumerate(pos):
        assert raw[n, y, x] == 0.5 and raw[n].count_nonzero() == 1
    assert {2 * (y & 1) + (x & 1) for y, x in pos[:4]} == {0, 1, 2, 3}
    for corner in [(0, 0), (0, w - 1), (h - 1, 0), (h - 1, w - 1)]:
        assert corner in pos


def test_perturbed_state_moves_only_the_trainable_tensors():
    from oracle import isp_oracle
    st = isp_oracle.default_state(syn.CAMERA_PRESETS["drone"])
    pt = syn.perturbed_state(st, scale=0.01, seed=1)
    assert set(pt) == set(st)
    for k in st:
        if k in syn.TRAINABLE_KEYS:
            assert not torch.equal(pt[k], st[k]) and (pt[k] - st[k]).abs().max() < 0.06
        else:
            assert torch.equal(pt[k], st[k])
    assert pt["debayer.weight"].count_nonzero() == 81              # every cross-channel tap receives a value
    assert torch.equal(syn.perturbed_state(st)["gamma_correct"], pt["gamma_correct"])


def test_targets_have_the_datasets_shapes():
    y = syn.labels(9, num_classes=16)
    assert y.shape == (9,) and y.dtype == torch.int64 and 0 <= y.min() and y.max() < 16
    m = syn.masks(3, 40, 56)
    assert m.shape == (3, 1, 40, 56) and set(m.unique().tolist()) <= {0.0, 1.0}      # dataset.py:144
