"""The numpy-chain restatement (oracle/numpy_oracle.py; timing baseline, parity unpinned for its third-party parts)
against the torch-chain oracle: the two chains are the same arithmetic away from the border (8e-7 in the
interior; borders differ by design -- half-sample vs whole-sample reflection, clip at 0 vs 1e-5)."""
import numpy as np
import torch

from oracle import isp_oracle, numpy_oracle
from raw2logit_b200 import synthetic as syn


def test_numpy_chain_matches_torch_chain_in_the_interior():
    cam = syn.CAMERA_PRESETS["drone"]
    raw = syn.smooth_scene(2, 64, 96, "drone", seed=5)
    want = isp_oracle.forward(raw, isp_oracle.default_state(cam))[0].numpy()
    for b in range(2):
        got = numpy_oracle.process_chw(raw[b].numpy().copy(), cam)
        assert got.shape == (3, 64, 96) and got.dtype == np.float32
        inner = (slice(None), slice(6, -6), slice(6, -6))
        positive = want[b][inner] > 2e-3                      # away from the torch chain's low clip at 1e-5
        diff = np.abs(got[inner] - want[b][inner])[positive]
        assert diff.max() <= 2e-6, diff.max()


def test_numpy_chain_does_not_touch_its_input_and_pool_matches_serial():
    import multiprocessing as mp
    cam = syn.CAMERA_PRESETS["microscopy"]
    raw = (0.25 * syn.smooth_scene(4, 32, 32, "microscopy", seed=2)).numpy()
    keep = raw.copy()
    serial = numpy_oracle.process_batch(list(raw), cam)
    assert np.array_equal(raw, keep)
    with mp.get_context("spawn").Pool(2) as pool:
        pooled = numpy_oracle.process_batch(list(raw), cam, pool)
    for a, b in zip(serial, pooled):
        assert np.array_equal(a, b)
