"""Seeded synthetic ridge-pattern images of the benchmark shapes.

Host-side numpy generator shared by the tests and bench.py: the same arrays are
fed to the CUDA path and to the CPU oracle.
"""
from __future__ import annotations

import numpy as np

__all__ = ["ridge_image", "ridge_batch", "degraded_image"]


def ridge_image(h: int = 320, w: int = 240, seed: int = 0, period: float | None = 9.0,
                noise_sigma: float = 12.0, jitter: float = 10.0) -> np.ndarray:
    """One uint8 [h, w] print: `phase = 2*pi*(r + 6*sin(2*phi))/period` about a jittered
    core, `I = 60 + 150*(0.5 + 0.5*cos(phase))` inside an ellipse (semi-axes 0.42 w,
    0.46 h), background 235, plus N(0, sigma) noise, clipped."""
    rng = np.random.default_rng(seed)
    if period is None:
        period = float(rng.uniform(7.0, 11.0))
    cx = w / 2.0 + float(rng.uniform(-jitter, jitter))
    cy = h / 2.0 + float(rng.uniform(-jitter, jitter))
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    dx, dy = xx - cx, yy - cy
    r = np.hypot(dx, dy)
    phi = np.arctan2(dy, dx)
    phase = 2.0 * np.pi * (r + 6.0 * np.sin(2.0 * phi)) / period
    img = 60.0 + 150.0 * (0.5 + 0.5 * np.cos(phase))
    inside = ((xx - w / 2.0) / (0.42 * w)) ** 2 + ((yy - h / 2.0) / (0.46 * h)) ** 2 <= 1.0
    img = np.where(inside, img, 235.0)
    img = img + rng.normal(0.0, noise_sigma, size=img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def degraded_image(h: int = 512, w: int = 512, seed: int = 0, period: float = 9.0,
                   noise_sigma: float = 40.0) -> np.ndarray:
    """NIST-shape degraded print: heavy noise, a saturated band (top 15 % of rows -> 255)
    and a few random black blobs."""
    rng = np.random.default_rng(seed + 100003)
    img = ridge_image(h, w, seed=seed, period=period, noise_sigma=noise_sigma).copy()
    img[: int(0.15 * h)] = 255
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(4):
        by, bx = int(rng.integers(0, h)), int(rng.integers(0, w))
        rad = int(rng.integers(h // 40 + 2, h // 16 + 3))
        img[(yy - by) ** 2 + (xx - bx) ** 2 <= rad * rad] = 0
    return img


def ridge_batch(n: int, h: int = 320, w: int = 240, first_seed: int = 0,
                noise_sigma: float = 12.0) -> np.ndarray:
    """uint8 [n, h, w]; image i uses seed first_seed+i, period U[7,11], core jitter +-10 px."""
    out = np.empty((n, h, w), dtype=np.uint8)
    for i in range(n):
        out[i] = ridge_image(h, w, seed=first_seed + i, period=None, noise_sigma=noise_sigma)
    return out
