"""Seeded synthetic ridge-pattern images of the benchmark shapes.

Host-side numpy generator shared by the tests and bench.py: the same arrays are
fed to the CUDA path and to the CPU oracle.
"""
from __future__ import annotations

import numpy as np

__all__ = ["ridge_image", "ridge_batch", "degraded_image", "ridge_image_counter", "philox4x32_10"]


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


# ---------------------------------------------------------------------------------------------------------------------
# NumPy twin of the on-device generator (csrc/k_synth.cu, fpb_synth_ridge): same Philox4x32-10 integers and the same
# float32 formula.  The GPU's sinf/cosf/logf differ from NumPy's in the last ulp, so a pixel can differ by one grey level
# (tests/test_gpu_io.py bounds it); anything that needs the exact device image fetches it with `fetch_input`.
# ---------------------------------------------------------------------------------------------------------------------
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(v, np.uint64) & np.uint64(0xFFFFFFFF) for v in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = np.uint64(k0 & 0xFFFFFFFF), np.uint64(k1 & 0xFFFFFFFF)
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(0xD2511F53) * c0, np.uint64(0xCD9E8D57) * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & m32, p1 & m32, ((p0 >> np.uint64(32)) ^ c3 ^ k1) & m32, p0 & m32
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & m32, (k1 + np.uint64(0xBB67AE85)) & m32
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def _u01(x):
    return ((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def ridge_image_counter(h: int = 320, w: int = 240, seed: int = 0, index: int = 0, period: float = 0.0,
                        noise_sigma: float = 12.0, jitter: float = 10.0) -> np.ndarray:
    """Image `index` of the synthetic stream `seed`, as fpb_synth_ridge generates it on the device (up to one grey level)."""
    f = np.float32
    k0, k1, i0, i1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF, index & 0xFFFFFFFF, (index >> 32) & 0xFFFFFFFF
    pr = philox4x32_10(i0, i1, 0xFFFFFFFF, 0, k0, k1)
    per = f(period) if period > 0 else f(7.0) + f(4.0) * _u01(pr[0])
    cx = f(0.5) * f(w) + f(jitter) * (f(2.0) * _u01(pr[1]) - f(1.0))
    cy = f(0.5) * f(h) + f(jitter) * (f(2.0) * _u01(pr[2]) - f(1.0))
    yy, xx = np.mgrid[0:h, 0:w]
    pair = (yy * ((w + 1) // 2) + (xx >> 1)).astype(np.uint64)
    rn = philox4x32_10(i0, i1, pair, 1, k0, k1)
    rad = np.sqrt(f(-2.0) * np.log(_u01(rn[0])), dtype=np.float32)
    ang = f(6.28318530718) * _u01(rn[1])
    nz = np.where((xx & 1) == 0, rad * np.cos(ang, dtype=np.float32), rad * np.sin(ang, dtype=np.float32)).astype(np.float32)
    dx, dy = xx.astype(np.float32) - cx, yy.astype(np.float32) - cy
    r = np.sqrt(dx * dx + dy * dy, dtype=np.float32)
    phi = np.arctan2(dy, dx, dtype=np.float32)
    v = f(60.0) + f(150.0) * (f(0.5) + f(0.5) * np.cos(f(6.28318530718) * (r + f(6.0) * np.sin(f(2.0) * phi, dtype=np.float32)) / per, dtype=np.float32))
    ex = (xx.astype(np.float32) - f(0.5) * f(w)) / (f(0.42) * f(w)); ey = (yy.astype(np.float32) - f(0.5) * f(h)) / (f(0.46) * f(h))
    v = np.where(ex * ex + ey * ey > f(1.0), f(235.0), v).astype(np.float32)
    v = v + f(noise_sigma) * nz
    return np.clip(v, 0, 255).astype(np.uint8)
