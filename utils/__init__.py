"""Drop-in module path of the reference's ``utils`` package for the parts this repo replaces (``utils.ssim``)."""
