"""ORACLE pinning for the GPU input pipeline — test infrastructure, run in the build container:

    python oracle/make_golden_aug.py

The reference's transform (utils/trainer.py:88-101) is Albumentations 2.0.8 over OpenCV 4.12 (requirements.txt); neither
library is vendored in the reference and Albumentations is not installed here, but OpenCV is, and every geometric /
photometric step of that Compose is ONE OpenCV call.  This script restates the Compose with those calls — cv2.resize
(INTER_LINEAR; INTER_NEAREST for the mask), cv2.warpAffine with the ShiftScaleRotate matrix
(cv2.getRotationMatrix2D(centre, angle, scale) + shift), cv2.flip, the truncating uint8 LUT of
RandomBrightnessContrast, Normalize — on synthetic X-ray-like uint8 images with FIXED parameters, and stores inputs,
parameters and outputs in tests/golden/augment.npz for tests/test_gpu_augment.py."""
import sys
from pathlib import Path

import cv2
import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.synthetic import IMAGENET_MEAN, IMAGENET_STD, xray_batch  # noqa: E402

S = 64            # output size of the fixture (the reference uses 256; the arithmetic is size independent)
HS, WS = 96, 80   # raw image size

CASES = [
    {"flip": False},                                                                       # val transform: resize only
    {"angle": 11.0, "scale": 1.04, "dx": 0.03, "dy": -0.05, "flip": False},
    {"angle": -15.0, "scale": 0.95, "dx": -0.05, "dy": 0.02, "flip": True, "alpha": 1.08, "beta": -0.06},
    {"flip": True, "alpha": 0.91, "beta": 0.1},
    {"angle": 3.3, "scale": 1.0, "dx": 0.0, "dy": 0.0, "flip": False, "alpha": 1.1, "beta": 0.1},
]


def raw_images(n):
    x, t = xray_batch(n, HS, WS, seed=77)
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    img = ((x * std + mean).clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous().numpy()
    img[..., 1] = np.roll(img[..., 1], 3, axis=1)          # make the three channels differ
    img[..., 2] = 255 - img[..., 2]
    msk = (t[:, 0] * 255).to(torch.uint8).numpy()
    return img, msk


def cpu_pipeline(img, msk, s, border):
    r = cv2.resize(img, (S, S), interpolation=cv2.INTER_LINEAR)
    m = cv2.resize(msk, (S, S), interpolation=cv2.INTER_NEAREST)
    if "angle" in s:
        c = (S / 2 - 0.5, S / 2 - 0.5)
        M = cv2.getRotationMatrix2D(c, s["angle"], s["scale"])
        M[0, 2] += s["dx"] * S
        M[1, 2] += s["dy"] * S
        bm = cv2.BORDER_CONSTANT if border == "constant" else cv2.BORDER_REFLECT_101
        r = cv2.warpAffine(r, M, (S, S), flags=cv2.INTER_LINEAR, borderMode=bm, borderValue=0)
        m = cv2.warpAffine(m, M, (S, S), flags=cv2.INTER_NEAREST, borderMode=bm, borderValue=0)
    if s.get("flip"):
        r, m = cv2.flip(r, 1), cv2.flip(m, 1)
    if "alpha" in s:
        lut = np.arange(256, dtype=np.float32) * np.float32(s["alpha"]) + np.float32(s["beta"] * 255.0)
        r = cv2.LUT(r, np.clip(lut, 0, 255).astype(np.uint8))
    mean = np.array(IMAGENET_MEAN, dtype=np.float32) * 255.0
    inv = 1.0 / (np.array(IMAGENET_STD, dtype=np.float32) * 255.0)
    x = ((r.astype(np.float32) - mean) * inv).transpose(2, 0, 1)
    return x.astype(np.float32), (m.astype(np.float32) / 255.0)[None]


def main():
    img, msk = raw_images(len(CASES))
    out = {"img": img, "mask": msk, "size": np.array(S)}
    for border in ("constant", "reflect101"):
        xs, ts = [], []
        for i, s in enumerate(CASES):
            x, t = cpu_pipeline(img[i], msk[i], s, border)
            xs.append(x)
            ts.append(t)
        out[f"x_{border}"] = np.stack(xs)
        out[f"t_{border}"] = np.stack(ts)
    keys = ("angle", "scale", "dx", "dy", "flip", "alpha", "beta")
    out["params"] = np.array([[float(s.get(k, np.nan)) for k in keys] for s in CASES], dtype=np.float64)
    out["param_keys"] = np.array(keys)
    np.savez_compressed(ROOT / "tests" / "golden" / "augment.npz", **out)
    print("wrote tests/golden/augment.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
