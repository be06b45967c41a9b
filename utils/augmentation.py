"""Module path of the reference's ``utils/augmentation.py``: re-exports the kernel-backed version."""
from raw2logit_b200.augmentation import (AddGaussianNoise, ComposeState, RandomRotate90, augmentation_strong,  # noqa: F401
                                         augmentation_weak, dihedral_handoff, get_augmentation, set_global_seed)
