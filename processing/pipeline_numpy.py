"""Module path of the reference's ``processing/pipeline_numpy.py``: re-exports the kernel-backed static pipeline."""
from raw2logit_b200.pipeline_numpy import RawProcessingPipeline, processing  # noqa: F401
