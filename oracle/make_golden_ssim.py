"""Writes tests/golden/ssim.npz from the UNMODIFIED reference ``utils/ssim.py`` (imported from /root/reference in the
build container; the file only needs torch and numpy) -- TEST INFRASTRUCTURE.  usage: python oracle/make_golden_ssim.py"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("R2L_REF", "/root/reference")


def load_reference_ssim():
    spec = importlib.util.spec_from_file_location("_ref_ssim", os.path.join(REF, "utils", "ssim.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cases():
    """(name, img1, img2): processed-image-like pairs (values in (0, 1), the second a perturbed copy), odd / tiny / large."""
    g = torch.Generator().manual_seed(77)
    out = []
    for name, shape, noise in (("rgb_64x96", (2, 3, 64, 96), 0.05), ("rgb_37x50", (3, 3, 37, 50), 0.2),
                               ("gray_8x8", (1, 1, 8, 8), 0.1), ("rgb_130x70", (1, 3, 130, 70), 0.02),
                               ("identical", (2, 3, 40, 40), 0.0)):
        base = torch.rand(shape, generator=g)
        smooth = torch.nn.functional.avg_pool2d(base, 3, stride=1, padding=1)
        img1 = (0.2 + 0.6 * smooth).contiguous()
        img2 = (img1 + noise * (torch.rand(shape, generator=g) - 0.5)).clamp(0, 1).contiguous()
        out.append((name, img1, img2))
    return out


def main():
    ref = load_reference_ssim()
    data = {}
    for name, img1, img2 in cases():
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            a = img1.to(dt).clone().requires_grad_(True)
            b = img2.to(dt).clone().requires_grad_(True)
            mean = ref.SSIM(window_size=11)(a, b)
            mean.backward()
            per = ref.ssim(a.detach(), b.detach(), window_size=11, size_average=False)
            data[f"{name}.{tag}.mean"] = mean.detach().numpy()
            data[f"{name}.{tag}.per_image"] = per.numpy()
            data[f"{name}.{tag}.grad1"] = a.grad.numpy()
            data[f"{name}.{tag}.grad2"] = b.grad.numpy()
        data[f"{name}.img1"] = img1.numpy()
        data[f"{name}.img2"] = img2.numpy()
    path = os.path.join(ROOT, "tests", "golden", "ssim.npz")
    np.savez_compressed(path, **data)
    print(path, os.path.getsize(path), "bytes,", len(data), "arrays")


if __name__ == "__main__":
    sys.exit(main())
