"""Synthetic raw-Bayer generators and camera presets.

Replaces the reference's cloud-backed datasets (dataset.py:24-41) with generators that reproduce only the
*contract* of a dataset item: a float32 ``(H, W)`` RGGB mosaic in [0, 1] obtained from 16-bit integers
(``dataset.py:87`` divides by ``2**bits - 1``), plus the two camera-parameter presets the reference ships
(Drone ``dataset.py:209-213``, Microscopy ``dataset.py:290-294``) and the identity default
(``pipeline_torch.py:36-40``).

Everything is generated on the CPU with a seeded ``torch.Generator`` so tests, golden fixtures and the bench see
identical bits on every machine; callers move the result to the GPU.
"""
import math

import torch

# (black_level[4], white_balance[3], colour_matrix[9]) -- the values are camera calibration constants
CAMERA_PRESETS = {
    "default": ([0.0, 0.0, 0.0, 0.0],
                [1.0, 1.0, 1.0],
                [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]),
    "drone": ([0.0625, 0.0626, 0.0625, 0.0626],
              [2.86653646, 1.0, 1.73079425],
              [1.50768983, -0.33571374, -0.17197604, -0.23048614, 1.70698738, -0.47650126,
               -0.03119153, -0.32803956, 1.35923111]),
    "microscopy": ([9.834368023181512e-06] * 4,
                   [-0.6567, 1.9673, 3.5304],
                   [-2.0338, 0.0933, 0.4157, -0.0286, 2.6464, -0.0574, -0.5516, -0.0947, 2.9308]),
}

# scene brightness scale that keeps the pre-clip stage mostly inside (1e-5, 1) for each preset
_SCENE_SCALE = {"default": 1.0, "drone": 1.0, "microscopy": 0.25}


def quantise16(raw):
    """Round to the 16-bit grid the datasets live on (dataset.py:87): round(clamp(v,0,1)*65535)/65535."""
    return torch.round(raw.clamp(0.0, 1.0) * 65535.0) / 65535.0


def to_uint16(raw):
    """Integer view of a quantised mosaic (for the uint16 ingest path)."""
    return torch.round(raw.clamp(0.0, 1.0).double() * 65535.0).to(torch.int32).to(torch.uint16)


def smooth_scene(batch, height, width, preset="drone", seed=1234, noise=0.002):
    """G1: smooth sinusoidal scene seen through a grey-world RGGB sensor, 16-bit quantised, float32 (B,H,W)."""
    g = torch.Generator().manual_seed(seed)
    bl, wb, _ = CAMERA_PRESETS[preset]
    fx = torch.rand(batch, 1, 1, generator=g) * (3.0 / width)
    fy = torch.rand(batch, 1, 1, generator=g) * (3.0 / height)
    ph = torch.rand(batch, 1, 1, generator=g) * (2.0 * math.pi)
    yy = torch.arange(height, dtype=torch.float32).view(1, height, 1)
    xx = torch.arange(width, dtype=torch.float32).view(1, 1, width)
    scene = 0.15 + 0.35 * (0.5 + 0.5 * torch.sin(2.0 * math.pi * (fx * xx + fy * yy) + ph))
    scene = scene * _SCENE_SCALE[preset]
    gain = torch.tensor([1.0 / abs(wb[0]), 1.0 / wb[1], 1.0 / wb[1], 1.0 / abs(wb[2])])
    black = torch.tensor(bl, dtype=torch.float32)
    par = (2 * (torch.arange(height) % 2).view(height, 1) + (torch.arange(width) % 2).view(1, width))
    raw = scene * gain[par] + black[par] + noise * torch.randn(batch, height, width, generator=g)
    return quantise16(raw).float().contiguous()


def noise_stress(batch, height, width, seed=0):
    """G2: U(0,1) noise -- exercises both clip masks heavily (ill-conditioned at the low clip)."""
    g = torch.Generator().manual_seed(seed)
    return quantise16(torch.rand(batch, height, width, generator=g)).float().contiguous()


def impulses(height, width, positions, value=0.5):
    """G4: one image per position with a single non-zero Bayer site (bit-exact CFA / border indexing)."""
    raw = torch.zeros(len(positions), height, width)
    for n, (y, x) in enumerate(positions):
        raw[n, y, x] = value
    return raw


def impulse_positions(height, width):
    """The four CFA phases in the interior plus corners and edge midpoints."""
    cy, cx = (height // 2) & ~1, (width // 2) & ~1
    pos = [(cy, cx), (cy, cx + 1), (cy + 1, cx), (cy + 1, cx + 1),
           (0, 0), (0, width - 1), (height - 1, 0), (height - 1, width - 1),
           (0, cx), (height - 1, cx + 1), (cy, 0), (cy + 1, width - 1),
           (1, 1), (height - 2, width - 2)]
    return [(min(max(y, 0), height - 1), min(max(x, 0), width - 1)) for y, x in pos]


def perturbed_state(state, scale=0.01, seed=1):
    """Perturb every trainable ISP tensor by scale*N(0,1) so cross-channel demosaic taps become non-zero."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in state.items():
        if k in TRAINABLE_KEYS:
            out[k] = (v + scale * torch.randn(v.shape, generator=g, dtype=torch.float32).to(v.dtype)).clone()
        else:
            out[k] = v.clone()
    return out


TRAINABLE_KEYS = ("black_level", "white_balance", "colour_correction", "gamma_correct",
                  "debayer.weight", "sharpening_filter.weight", "gaussian_blur.weight")


def labels(batch, num_classes=16, seed=7):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, num_classes, (batch,), generator=g)


def masks(batch, height, width, seed=7):
    """{0,1} float masks like the Drone segmentation targets (dataset.py:144)."""
    g = torch.Generator().manual_seed(seed)
    cy = torch.rand(batch, 1, 1, generator=g) * height
    cx = torch.rand(batch, 1, 1, generator=g) * width
    r = (0.1 + 0.3 * torch.rand(batch, 1, 1, generator=g)) * min(height, width)
    yy = torch.arange(height, dtype=torch.float32).view(1, height, 1)
    xx = torch.arange(width, dtype=torch.float32).view(1, 1, width)
    return (((yy - cy) ** 2 + (xx - cx) ** 2) < r ** 2).float().unsqueeze(1)
