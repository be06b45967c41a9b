"""Module path of the reference's ISP package; the implementation lives in raw2logit_b200."""
