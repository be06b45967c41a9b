"""numpy-compatible static pipeline (processing/pipeline_numpy.py:70-141): the fused kernel against the
CPU restatement of the numpy chain (oracle/numpy_oracle.py, float64), borders included."""
import numpy as np
import pytest
import torch

from oracle import numpy_oracle
from raw2logit_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("preset,scale,shape", [("drone", 1.0, (3, 64, 96)), ("microscopy", 0.25, (2, 37, 50)),
                                               ("drone", 1.0, (1, 256, 256)), ("drone", 1.0, (2, 5, 7))])
def test_static_pipeline_matches_numpy_chain_everywhere(preset, scale, shape):
    from processing.pipeline_numpy import RawProcessingPipeline, processing
    cam = syn.CAMERA_PRESETS[preset]
    raw = (scale * syn.smooth_scene(*shape, preset, seed=9)).contiguous()
    got = processing(raw.cuda(), *cam, sharpening="sharpening_filter", denoising="gaussian_denoising").cpu().numpy()
    assert got.shape == (shape[0], 3, shape[1], shape[2]) and got.dtype == np.float32
    for b in range(shape[0]):
        want = numpy_oracle.processing(raw[b].numpy().astype(np.float64), *cam).transpose(2, 0, 1)
        # away from the clip at 0 (x ** (1 / 2.2) has an unbounded slope there) fp32 agrees with float64 to 1e-5
        safe = want > 0.02
        assert np.abs(got[b] - want)[safe].max() <= 1e-5, np.abs(got[b] - want)[safe].max()
        assert np.abs(got[b] - want).max() <= 2e-3
    # the class surface of the reference (one image in, (3, H, W) out) and the uint16 ingest
    pipe = RawProcessingPipeline(cam, sharpening="sharpening_filter", denoising="gaussian_denoising")
    one = pipe(raw[0].numpy())
    assert one.shape == (3, shape[1], shape[2]) and np.array_equal(one.cpu().numpy(), got[0])
    u16 = syn.to_uint16(raw)
    got16 = processing(u16.cuda(), *cam, sharpening="sharpening_filter", denoising="gaussian_denoising")
    ref16 = processing((u16.to(torch.int32).to(torch.float32) / 65535.0).cuda(), *cam, sharpening="sharpening_filter",
                       denoising="gaussian_denoising")
    assert torch.equal(got16, ref16)


def test_stage_switches_and_unserved_options():
    from processing.pipeline_numpy import processing
    cam = syn.CAMERA_PRESETS["drone"]
    raw = syn.smooth_scene(1, 32, 32, "drone", seed=3)
    plain = processing(raw.cuda(), *cam, sharpening="none", denoising="gaussian")       # names the reference skips
    want = numpy_oracle.remove_blacklv(raw[0].numpy().astype(np.float64), cam[0])
    want = numpy_oracle.demosaicing_cfa_bayer_bilinear(want) * np.asarray(cam[1])
    want = np.einsum('ijk,lk->ijl', want, np.asarray(cam[2]).reshape(3, 3))
    want = np.clip(want, 0, 1) ** (1 / 2.2)
    assert np.abs(plain[0].cpu().numpy() - want.transpose(2, 0, 1)).max() <= 1e-5
    for kw in ({"debayer": "menon2007"}, {"sharpening": "unsharp_masking"}, {"denoising": "median_denoising"}):
        with pytest.raises(NotImplementedError):
            processing(raw.cuda(), *cam, **kw)
