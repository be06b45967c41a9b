"""Module path of the reference's ``utils/ssim.py``: re-exports the kernel-backed SSIM (raw2logit_b200/ssim.py)."""
from raw2logit_b200.ssim import SSIM, ssim  # noqa: F401
