"""oracle/stages.py (numpy restatements of the insides of the OpenCV / SciPy / NumPy calls - the specification the
CUDA kernels were written from) against the libraries themselves.  CPU only."""
import warnings

import cv2
import numpy as np
import pytest
from scipy import ndimage as ndi

from oracle import stages as st
from oracle.skimage_compat import threshold_otsu

RNG = np.random.default_rng(0)


def _img(h, w, kind):
    if kind == 0:
        return RNG.integers(0, 256, (h, w)).astype(np.uint8)
    if kind == 1:
        return np.clip(RNG.normal(128, 30, (h, w)), 0, 255).astype(np.uint8)
    if kind == 2:
        return (np.add.outer(np.arange(h), np.arange(w)) % 256).astype(np.uint8)
    return np.where(RNG.random((h, w)) < 0.7, 235, RNG.integers(0, 256, (h, w))).astype(np.uint8)


def test_percentile_stretch_bit_exact():
    for t in range(60):
        h, w = int(RNG.integers(40, 700)), int(RNG.integers(40, 700))
        img = np.clip(RNG.normal(RNG.uniform(60, 200), RNG.uniform(5, 60), (h, w)), 0, 255).astype(np.uint8)
        f = img.astype(np.float32) / 255.0
        lo = np.percentile(f, 0.5)
        span = np.percentile(f, 99.5) - np.percentile(f, 0.5) + 1e-12
        ref = (np.clip((f - lo) / span, 0.0, 1.0) * 255).astype(np.uint8)
        assert np.array_equal(st.stretch_lut(img)[img], ref)


@pytest.mark.parametrize("clip", [2.0, 2.5])
def test_clahe_bit_exact(clip):
    for t in range(24):
        h, w = (320, 240) if t % 3 == 0 else (int(RNG.integers(40, 400)), int(RNG.integers(40, 400)))
        img = _img(h, w, t % 4)
        assert np.array_equal(st.clahe(img, clip), cv2.createCLAHE(clipLimit=clip, tileGridSize=(8, 8)).apply(img))


def test_nlm_bit_exact():
    for t in range(3):
        h, w = int(RNG.integers(30, 90)), int(RNG.integers(30, 90))
        yy, xx = np.mgrid[0:h, 0:w]
        img = np.clip(128 + 80 * np.cos(xx / 1.5 + yy / 3) + RNG.normal(0, 12, (h, w)), 0, 255).astype(np.uint8)
        ref = cv2.fastNlMeansDenoising(img, None, h=10, templateWindowSize=7, searchWindowSize=21)
        assert np.array_equal(st.nlm(img), ref)
    tab, shift, mult = st.nlm_weight_table()
    assert (tab > 0).sum() == 528 and shift == 6 and mult == 19096 and tab[0] == 19096


def test_fixed_point_gaussians_bit_exact():
    # heights below 12 are excluded: OpenCV 4.13's fixed-point 5x5 path returns a different row 1 for images of
    # 8..11 rows (a ring-buffer quirk of its vertical pass, not reflect-101); the hot path never sees such images
    for t in range(30):
        h, w = int(RNG.integers(12, 300)), int(RNG.integers(12, 300))
        img = _img(h, w, t % 2)
        assert np.array_equal(st.gauss_u8(img, st.GAUSS3_SIGMA06_TAPS), cv2.GaussianBlur(img, (3, 3), 0.6))
        assert np.array_equal(st.gauss_u8(img, st.GAUSS5_SIGMA0_TAPS), cv2.GaussianBlur(img, (5, 5), 0))


def test_cv_otsu_bit_exact():
    for t in range(60):
        h, w = int(RNG.integers(20, 300)), int(RNG.integers(20, 300))
        img = np.clip(np.where(RNG.random((h, w)) < RNG.uniform(.2, .8), RNG.normal(80, 25, (h, w)),
                               RNG.normal(200, 20, (h, w))), 0, 255).astype(np.uint8)
        if t % 9 == 0:
            img[:] = 77
        tv, _ = cv2.threshold(img, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        assert st.otsu_u8(np.bincount(img.ravel(), minlength=256), img.size) == int(tv)


def test_ellipse_and_binary_morphology_bit_exact():
    for k in (3, 15):
        se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
        hw = st.ellipse_half_widths(k)
        for i in range(k):
            assert se[i].sum() == 2 * hw[i] + 1 and se[i, k // 2 - hw[i]] == 1
    se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (15, 15))
    for t in range(10):
        m = (cv2.GaussianBlur(RNG.random((120, 150)).astype(np.float32), (0, 0), 4) > 0.5).astype(np.uint8) * 255
        assert np.array_equal(st.morph_binary(m, 15, "erode"), cv2.erode(m, se))
        assert np.array_equal(st.morph_binary(m, 15, "dilate"), cv2.dilate(m, se))


def test_box_filter_bit_exact():
    for t in range(15):
        h, w = int(RNG.integers(30, 330)), int(RNG.integers(30, 330))
        img = RNG.integers(0, 256, (h, w)).astype(np.float32)
        assert np.array_equal(st.box_mean_f32(img.astype(np.int64), 25), cv2.boxFilter(img, -1, (25, 25)))
        assert np.array_equal(st.box_mean_f32(img.astype(np.int64) ** 2, 25), cv2.boxFilter(img ** 2, -1, (25, 25)))
        sk = (RNG.random((h, w)) < 0.1).astype(np.float32)
        assert np.array_equal(st.box_mean_f32(sk.astype(np.int64), 25), cv2.blur(sk, (25, 25)))


def test_patch_otsu_bit_exact():
    warnings.simplefilter("ignore")
    for t in range(600):
        ph, pw = int(RNG.integers(2, 33)), int(RNG.integers(2, 33))
        if t % 3 == 0:
            p = RNG.integers(0, 256, (ph, pw))
        elif t % 3 == 1:
            p = np.clip(RNG.normal(RNG.uniform(30, 220), RNG.uniform(1, 50), (ph, pw)), 0, 255).astype(int)
        else:
            p = np.where(RNG.random((ph, pw)) < 0.5, RNG.integers(0, 100, (ph, pw)), RNG.integers(150, 256, (ph, pw)))
        r = threshold_otsu(p.astype(np.float32))
        assert np.float32(r) == st.patch_otsu(p) and np.asarray(r).dtype == np.float32


def test_scipy_gaussian_and_sobel_bit_exact():
    for t in range(12):
        h, w = int(RNG.integers(5, 200)), int(RNG.integers(5, 200))
        a = RNG.random((h, w)).astype(np.float32) * (1000 if t % 2 else 1)
        for s in (0.6, 1.5, 2.0, 3.0):
            assert np.array_equal(st.gaussian_filter_f32(a, s), ndi.gaussian_filter(a, sigma=s))
        for ax in (0, 1):
            assert np.array_equal(st.ndi_sobel_f32(a, ax), ndi.sobel(a, axis=ax))
    a = RNG.random((19, 13)).astype(np.float32)          # block grid: 25 taps wrap around 13 columns
    assert np.array_equal(st.gaussian_filter_f32(a, 3.0), ndi.gaussian_filter(a, sigma=3.0))


def test_cv_sobel_and_resize_within_float_tolerance():
    a = (RNG.random((50, 61)) * 255).astype(np.float32)
    for dx in (1, 0):
        ref = cv2.Sobel(a, cv2.CV_32F, dx, 1 - dx, ksize=3)
        np.testing.assert_allclose(st.cv_sobel_f32(a, dx), ref, rtol=0, atol=2.5e-4)     # 1 ulp at |v| ~ 1000
    g = RNG.random((19, 13)).astype(np.float32)
    np.testing.assert_allclose(st.resize_linear_f32(g, 222, 315), cv2.resize(g, (222, 315), interpolation=cv2.INTER_LINEAR),
                               rtol=0, atol=2e-6)
