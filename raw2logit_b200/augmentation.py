"""Kernel-backed drop-in for the reference's ``utils/augmentation.py`` -- augmentation and hand-off to the task model in
one pass.

Same names (``RandomRotate90``, ``AddGaussianNoise``, ``set_global_seed``, ``ComposeState``, ``augmentation_weak``,
``augmentation_strong``, ``get_augmentation``) and the same seed-replay contract (``ComposeState.__call__(x,
retain_state, mask_transform)``, utils/augmentation.py:55-67): a mask transformed with ``retain_state`` /
``mask_transform`` sees the same random draws as its image.

What changes: when every active transform is a flip or a quarter turn (``augmentation_weak``, :70-74) and the input is a
CUDA float batch, the random draws are made in the reference's order (``torch.rand(1) < p`` per flip, as torchvision does;
``random.randint(0, 3)`` for the turn, :9-10) and their composition -- one element of the dihedral group -- is applied by
ONE kernel (``csrc/isp_handoff.cu``) that can also emit the layout / dtype the task model wants
(``ComposeState(..., memory_format=torch.channels_last, dtype=torch.bfloat16)``): one read and one write instead of up
to five passes of stock ops.  Differentiable (the adjoint is the same kernel with the inverse map).  Any other transform
(``augmentation_strong``'s rotation by an arbitrary angle, noise, sharpness) runs as the stock op it is.
"""
import random

import numpy as np
import torch
import torchvision.transforms as T

from . import ops  # noqa: F401  (loads the operator library)


class RandomRotate90():  # Note: not the same as T.RandomRotation(90)   (utils/augmentation.py:8-14)
    def __call__(self, x):
        x = x.rot90(random.randint(0, 3), dims=(-1, -2))
        return x

    def __repr__(self):
        return self.__class__.__name__


class AddGaussianNoise():                                                # utils/augmentation.py:17-31
    def __init__(self, std=0.01):
        self.std = std

    def __call__(self, x):
        return x + torch.randn_like(x) * self.std

    def __repr__(self):
        return self.__class__.__name__ + f'(std={self.std})'


def set_global_seed(seed):                                               # utils/augmentation.py:34-37
    torch.random.manual_seed(seed)
    np.random.seed(seed % (2**32 - 1))
    random.seed(seed)


# ---- dihedral index maps: destination (y, x) -> source (y, x) as affine coefficients ---------------------------------
def _compose(h, w, ops_):
    """ops_: list of ('h',), ('v',), ('r', k).  Returns (h_dst, w_dst, [a0, a1, a2, b0, b1, b2]) of the composition applied in
    order (torch semantics: F.hflip / F.vflip, Tensor.rot90(k, dims=(-1, -2)))."""
    maps = []                                            # each: (function dst->src on the tensor it is applied to)
    ch, cw = h, w
    for op in ops_:
        if op[0] == 'h':
            maps.append(lambda y, x, cw=cw: (y, cw - 1 - x))
        elif op[0] == 'v':
            maps.append(lambda y, x, ch=ch: (ch - 1 - y, x))
        else:
            k = op[1] % 4
            if k == 1:        # flip(-2) then transpose: out[a][b] = in[H-1-b][a], shape (W, H)
                maps.append(lambda a, b, ch=ch: (ch - 1 - b, a))
                ch, cw = cw, ch
            elif k == 2:
                maps.append(lambda y, x, ch=ch, cw=cw: (ch - 1 - y, cw - 1 - x))
            elif k == 3:      # flip(-1) then transpose: out[a][b] = in[b][W-1-a], shape (W, H)
                maps.append(lambda a, b, cw=cw: (b, cw - 1 - a))
                ch, cw = cw, ch

    def full(y, x):
        for f in reversed(maps):
            y, x = f(y, x)
        return y, x

    y00, x00 = full(0, 0)
    y10, x10 = full(1, 0)
    y01, x01 = full(0, 1)
    return ch, cw, [y00, y10 - y00, y01 - y00, x00, x10 - x00, x01 - x00]


def _invert(h_dst, w_dst, m):
    """The inverse of a dihedral map (destination <-> source swapped)."""
    a0, a1, a2, b0, b1, b2 = m
    if a2 == 0:                                          # plain: ys = a0 + a1 y, xs = b0 + b2 x
        return [-a0 * a1, a1, 0, -b0 * b2, 0, b2]        # y = a1 (ys - a0), x = b2 (xs - b0)   (a1, b2 = +-1)
    # transposed: ys = a0 + a2 x, xs = b0 + b1 y  ->  y = b1 (xs - b0), x = a2 (ys - a0)
    return [-b0 * b1, 0, b1, -a0 * a2, a2, 0]


class _DihedralHandoff(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, map6, h_dst, w_dst, channels_last, to_bf16):
        ctx.inv = _invert(h_dst, w_dst, map6)
        ctx.src_hw = (x.shape[2], x.shape[3])
        ctx.src_dtype = x.dtype
        return torch.ops.raw2logit_isp.dihedral_copy(x, list(map6), int(h_dst), int(w_dst), bool(channels_last), bool(to_bf16))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        g = torch.ops.raw2logit_isp.dihedral_copy(grad, ctx.inv, ctx.src_hw[0], ctx.src_hw[1], False,
                                                  ctx.src_dtype == torch.bfloat16)
        return g, None, None, None, None, None


def dihedral_handoff(x, hflip=False, vflip=False, quarter_turns=0, memory_format=None, dtype=None):
    """Flip(s) + ``rot90(quarter_turns, dims=(-1, -2))`` of a (B, C, H, W) CUDA batch, emitted in ``memory_format``
    (None / contiguous or ``torch.channels_last``) and ``dtype`` (None / float32 or ``torch.bfloat16``), in one kernel."""
    seq = ([('h',)] if hflip else []) + ([('v',)] if vflip else []) + ([('r', quarter_turns)] if quarter_turns % 4 else [])
    h_dst, w_dst, m = _compose(x.shape[2], x.shape[3], seq)
    return _DihedralHandoff.apply(x, m, h_dst, w_dst, memory_format == torch.channels_last, dtype == torch.bfloat16)


_DIHEDRAL = (T.RandomHorizontalFlip, T.RandomVerticalFlip, RandomRotate90)


class ComposeState(T.Compose):
    """utils/augmentation.py:39-67, plus the hand-off options (``memory_format``, ``dtype``: what the task model wants
    to receive; None = the reference's contiguous float32)."""

    def __init__(self, transforms, memory_format=None, dtype=None):
        self.transforms = []
        self.mask_transforms = []
        for t in transforms:
            apply_for_mask = True
            if isinstance(t, tuple):
                t, apply_for_mask = t
            self.transforms.append(t)
            if apply_for_mask:
                self.mask_transforms.append(t)
        self.seed = None
        self.memory_format = memory_format
        self.dtype = dtype

    def __call__(self, x, retain_state=False, mask_transform=False):
        if self.seed is not None:   # retain previous state
            set_global_seed(self.seed)
        if retain_state:    # save state for next call
            self.seed = self.seed or torch.seed()
            set_global_seed(self.seed)
        else:
            self.seed = None    # reset / ignore state

        transforms = self.transforms if not mask_transform else self.mask_transforms
        fusable = (isinstance(x, torch.Tensor) and x.is_cuda and x.ndim == 4 and x.shape[1] <= 4 and
                   x.dtype in (torch.float32, torch.bfloat16) and len(transforms) > 0 and
                   all(isinstance(t, _DIHEDRAL) for t in transforms))
        if fusable:
            seq = []
            for t in transforms:                       # the reference's draws, in the reference's order
                if isinstance(t, T.RandomHorizontalFlip):
                    if torch.rand(1) < t.p:
                        seq.append(('h',))
                elif isinstance(t, T.RandomVerticalFlip):
                    if torch.rand(1) < t.p:
                        seq.append(('v',))
                else:
                    k = random.randint(0, 3)
                    if k:
                        seq.append(('r', k))
            h_dst, w_dst, m = _compose(x.shape[2], x.shape[3], seq)
            return _DihedralHandoff.apply(x, m, h_dst, w_dst, self.memory_format == torch.channels_last,
                                          self.dtype == torch.bfloat16)
        for t in transforms:
            x = t(x)
        if isinstance(x, torch.Tensor) and (self.memory_format is not None or self.dtype is not None) and x.ndim == 4:
            x = x.to(dtype=self.dtype or x.dtype, memory_format=self.memory_format or torch.contiguous_format)
        return x


augmentation_weak = ComposeState([
    T.RandomHorizontalFlip(),
    T.RandomVerticalFlip(),
    RandomRotate90(),
])


augmentation_strong = ComposeState([
    T.RandomHorizontalFlip(p=0.5),
    T.RandomVerticalFlip(p=0.5),
    T.RandomApply([T.RandomRotation(90)], p=0.5),
    # (transform, apply_to_mask=True)
    (T.RandomApply([AddGaussianNoise(std=0.0005)], p=0.5), False),
    (T.RandomAdjustSharpness(0.5, p=0.5), False),
])


def get_augmentation(type):
    if type == 'none':
        return None
    if type == 'weak':
        return augmentation_weak
    if type == 'strong':
        return augmentation_strong
