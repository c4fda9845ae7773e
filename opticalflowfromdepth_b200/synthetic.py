"""Synthetic RGB-D frames shaped like the reference's datasets, seeded by frame index.

DIML-shaped: an 8-bit disparity image (ground-plane ramp + constant-disparity rectangles + noise, rounded and clipped
to [0,255]) scaled by 63/255 as utils.get_disparity does (utils.py:61-72), turned into depth by
Convert.disparity_to_depth (preprocess.py:257-262).  Images are uint8-valued float32, i.i.d. uniform.
Pure numpy on the host; used by bench.py, tests and the sweep driver — no dataset files, no network.
"""
from __future__ import annotations

import numpy as np


def diml_disparity_u8(rng: np.random.Generator, h: int, w: int, rects: int = 12) -> np.ndarray:
    y, x = np.mgrid[0:h, 0:w]
    disp = 20.0 + 15.0 * y / h + 3.0 * np.sin(2 * np.pi * x / w)
    for _ in range(rects):
        r0, c0 = int(rng.integers(0, max(1, h - 8))), int(rng.integers(0, max(1, w - 8)))
        r1 = r0 + int(rng.integers(8, max(9, h // 2)))
        c1 = c0 + int(rng.integers(8, max(9, w // 2)))
        disp[r0:r1, c0:c1] = rng.uniform(30, 250)
    disp = disp + rng.normal(0, 0.7, disp.shape)
    return np.clip(np.round(disp), 0, 255)


def diml_frame(idx: int, h: int = 480, w: int = 640, dtype=np.float32):
    """(img[3,h,w] float32 in 0..255, raw_depth[1,h,w] dtype) for frame `idx` (seed = idx)."""
    rng = np.random.default_rng(idx)
    img = rng.integers(0, 256, (3, h, w)).astype(np.float32)
    disp = diml_disparity_u8(rng, h, w) * 63.0 / 255.0
    depth = (50.0 / (disp + 0.005)).astype(dtype)[None]
    return img, depth


def redweb_sizes(n: int, seed: int = 1):
    """Mixed resolutions of SURVEY cfg2: 0.3-2 MP log-uniform, aspect U(0.6,1.8), even H and W."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        mp = np.exp(rng.uniform(np.log(0.3e6), np.log(2e6)))
        aspect = rng.uniform(0.6, 1.8)
        w = int(round(np.sqrt(mp * aspect) / 2)) * 2
        h = int(round(np.sqrt(mp / aspect) / 2)) * 2
        out.append((h, w))
    return out


def redweb_frame(idx: int, h: int, w: int, dtype=np.float32):
    """ReDWeb-shaped: relative-depth uint8 map (smooth field + rectangles) -> utils.smooth_closer (utils.py:118-121)."""
    rng = np.random.default_rng(10_000 + idx)
    img = rng.integers(0, 256, (3, h, w)).astype(np.float32)
    y, x = np.mgrid[0:h, 0:w]
    rel = 120 + 60 * np.sin(2 * np.pi * (x / w + rng.uniform())) * np.cos(2 * np.pi * (y / h + rng.uniform()))
    for _ in range(10):
        r0, c0 = int(rng.integers(0, h - 8)), int(rng.integers(0, w - 8))
        r1, c1 = r0 + int(rng.integers(8, h // 2)), c0 + int(rng.integers(8, w // 2))
        rel[r0:r1, c0:c1] = rng.uniform(10, 250)
    rel = np.clip(np.round(rel), 0, 255)
    rel[rel > 240] = 240
    depth = (1.0 / (255.0 - rel)).astype(dtype)[None]
    return img, depth
