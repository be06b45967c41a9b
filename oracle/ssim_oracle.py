"""CPU restatement of the reference's SSIM (``/root/reference/utils/ssim.py``) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/`` import it.  Pinned by ``tests/golden/ssim.npz`` (outputs of the unmodified reference file, written by
``oracle/make_golden_ssim.py``) and, when ``/root/reference`` is present, by the live reference."""
import math

import torch
import torch.nn.functional as F


def window(window_size=11, channel=3, dtype=torch.float32):
    """ssim.py:9-17: 1-D Gaussian (sigma 1.5) from Python floats, fp32, normalised; outer product; one copy per channel."""
    g = torch.Tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    w2 = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous().to(dtype)


def ssim(img1, img2, window_size=11, size_average=True):
    """ssim.py:19-39."""
    c = img1.shape[1]
    w = window(window_size, c, img1.dtype)
    pad = window_size // 2
    mu1 = F.conv2d(img1, w, padding=pad, groups=c)
    mu2 = F.conv2d(img2, w, padding=pad, groups=c)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, w, padding=pad, groups=c) - mu1_sq
    s2 = F.conv2d(img2 * img2, w, padding=pad, groups=c) - mu2_sq
    s12 = F.conv2d(img1 * img2, w, padding=pad, groups=c) - mu1_mu2
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1_mu2 + c1) * (2 * s12 + c2)) / ((mu1_sq + mu2_sq + c1) * (s1 + s2 + c2))
    return m.mean() if size_average else m.mean(1).mean(1).mean(1)
