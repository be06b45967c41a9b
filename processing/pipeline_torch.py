"""Drop-in for the reference's ``processing/pipeline_torch.py`` (same module path, same public names).

``from processing.pipeline_torch import ParametrizedProcessing, RawToRGB, raw2rgb, ...`` keeps working for the
reference's ``model.py`` / ``train.py`` and for pickled models that name this module; everything is implemented
by the CUDA-backed classes in ``raw2logit_b200.pipeline_torch``.
"""
from raw2logit_b200.pipeline_torch import (  # noqa: F401
    DEFAULT_CAMERA_PARAMS, K_BLUR, K_G, K_RB, K_SHARP, M_RGB_2_YUV, M_YUV_2_RGB,
    Debayer, NNProcessing, ParametrizedProcessing, RawToRGB, append_additive_layer, raw2rgb,
)
