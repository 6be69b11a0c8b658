"""The FPB_HD routines the CUDA kernels run (hd_geometry.h, hd_scalar.h), compiled by g++ (tests/hostcheck) and
compared with OpenCV / NumPy / SciPy on random inputs.  Test infrastructure: the package never loads this build."""
import ctypes
import warnings

import pytest

import cv2
import numpy as np

from oracle import stages as st
from oracle.skimage_compat import threshold_otsu

RNG = np.random.default_rng(1)
VP = ctypes.c_void_p


def _p(a):
    return a.ctypes.data_as(VP)


def _hull_fill(lib, mask):
    H, W = mask.shape
    out = np.zeros((H, W), np.uint8); bbox = np.zeros(4, np.int32); a2 = ctypes.c_longlong(0)
    hull = np.zeros(2 * (2 * H + 4), np.int32)
    n = lib.hc_largest_hull_fill(_p(mask), H, W, _p(out), _p(bbox), ctypes.byref(a2), _p(hull), 2 * H + 4)
    return n, out, tuple(int(v) for v in bbox), a2.value


def test_segmentation_geometry_matches_opencv(hostcheck):
    """findContours(EXTERNAL) -> max contourArea -> convexHull -> drawContours(fill) -> boundingRect."""
    se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (15, 15))
    checked = 0
    for t in range(450):
        H, W = int(RNG.integers(20, 330)), int(RNG.integers(20, 330))
        if t % 3 == 0:
            m = (cv2.GaussianBlur(RNG.random((H, W)).astype(np.float32), (0, 0), RNG.uniform(2, 12)) > 0.5).astype(np.uint8) * 255
        elif t % 3 == 1:
            m = (cv2.GaussianBlur(RNG.random((H, W)).astype(np.float32), (0, 0), RNG.uniform(3, 10)) > RNG.uniform(0.47, 0.53)).astype(np.uint8) * 255
            m = cv2.morphologyEx(cv2.morphologyEx(m, cv2.MORPH_CLOSE, se), cv2.MORPH_OPEN, se)
        else:
            m = (RNG.random((H, W)) < RNG.uniform(0.2, 0.7)).astype(np.uint8) * 255      # raw noise: thin bridges, holes
        n, out, bbox, a2 = _hull_fill(hostcheck, m)
        contours, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        if not contours:
            assert n == 0
            continue
        areas = sorted(cv2.contourArea(c) for c in contours)
        if len(areas) > 1 and areas[-1] == areas[-2]:
            continue                                    # max() tie-break of the reference is list-order dependent
        big = max(contours, key=cv2.contourArea)
        hull = cv2.convexHull(big)
        ref = np.zeros_like(m); cv2.drawContours(ref, [hull], -1, 255, -1)
        assert a2 == int(round(2 * cv2.contourArea(big)))
        assert bbox == tuple(cv2.boundingRect(hull))
        assert np.array_equal(out, ref)
        checked += 1
    assert checked > 400


def test_scalar_routines_match_libraries(hostcheck):
    warnings.simplefilter("ignore")
    for t in range(120):
        h, w = int(RNG.integers(40, 600)), int(RNG.integers(40, 600))
        img = np.clip(RNG.normal(RNG.uniform(60, 200), RNG.uniform(5, 60), (h, w)), 0, 255).astype(np.uint8)
        hist = np.bincount(img.ravel(), minlength=256).astype(np.uint32)
        lut = np.zeros(256, np.uint8)
        hostcheck.hc_stretch_lut(_p(hist), img.size, _p(lut))
        assert np.array_equal(lut, st.stretch_lut(img))
        tv, _ = cv2.threshold(img, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        assert hostcheck.hc_otsu_u8(_p(hist), img.size) == int(tv)
        p = np.clip(RNG.normal(RNG.uniform(30, 220), RNG.uniform(1, 50), (int(RNG.integers(2, 33)), int(RNG.integers(2, 33)))), 0, 255).astype(int)
        ih = np.bincount(p.ravel(), minlength=256).astype(np.uint32)
        assert hostcheck.hc_patch_otsu(_p(ih)) == np.float32(threshold_otsu(p.astype(np.float32)))


def test_gaussian_weight_tables_equal_scipy(hostcheck):
    from scipy.ndimage._filters import _gaussian_kernel1d
    for s in (0.6, 1.5, 2.0, 3.0):
        w = np.zeros(64)
        r = hostcheck.hc_gauss_weights(ctypes.c_double(s), _p(w))
        assert r == int(4.0 * s + 0.5)
        assert np.array_equal(w[:2 * r + 1], _gaussian_kernel1d(s, 0, r))


def test_thinning_tables_agree():
    """The product's built-in table (capi.cu) and the oracle's are derived independently from Zhang-Suen."""
    import re, os
    from conftest import ROOT
    from oracle.skimage_compat import zhang_suen_table
    tab = zhang_suen_table()
    assert {int(v): int((tab == v).sum()) for v in (1, 2, 3)} == {1: 6, 2: 6, 3: 28}
    src = open(os.path.join(ROOT, "multimodal_biometric_fingerprints_palms_b200", "csrc", "capi.cu")).read()
    assert "n * e * s == 0 && e * s * w == 0" in src and "n * e * w == 0 && n * s * w == 0" in src


def test_word_parallel_zhang_suen_equals_the_table(hostcheck):
    """fpb_zs_delete_mask (the closed form k_thin_extract uses when the built-in table is installed) on all 256 neighbourhoods."""
    from oracle.skimage_compat import zhang_suen_table
    out = np.zeros(256, np.uint8)
    hostcheck.hc_zs_codes(_p(out))
    assert np.array_equal(out, zhang_suen_table().astype(np.uint8))


def test_skimage_table_if_available():
    """R1 of SURVEY.md: when scikit-image is importable, the derived table must reproduce its skeletonize."""
    import pytest
    skm = pytest.importorskip("skimage.morphology")
    from oracle.skimage_compat import skeletonize
    for t in range(20):
        m = cv2.GaussianBlur(RNG.random((96, 96)).astype(np.float32), (0, 0), 2.0) > 0.5
        assert np.array_equal(skm.skeletonize(m), skeletonize(m)), "derived Zhang-Suen table != scikit-image's table"


def test_config_mirror_defaults_and_opt_in_overrides(tmp_path, monkeypatch):
    """config/config_fingerprint.py mirror: same names as the reference module; the numeric YAML sections are dead unless
    FPB200_YAML_OVERRIDES=1, and then only the keys a kernel parameter exists for are mapped."""
    import importlib
    from multimodal_biometric_fingerprints_palms_b200.config import config_fingerprint as cf
    for name in ("cfg", "get_path", "METADATA_DIR", "DATASET_DIR", "SORTED_DATASET_DIR", "PROCESSED_DIR", "FEATURES_DIR",
                 "DEBUG_DIR", "DB_CONFIG", "PREPROCESSING_PARAMS", "BINARIZATION_PARAMS", "ORIENTATION_PARAMS",
                 "GENERAL_PARAMS", "print_config_summary"):
        assert hasattr(cf, name), name
    monkeypatch.delenv("FPB200_YAML_OVERRIDES", raising=False)
    assert cf.active_overrides() == {}                                    # defaults stay the hard-coded values
    assert cf.HARD_CODED["post_params"]["quality_threshold"] == 0.15 and cf.HARD_CODED["rel_threshold"] == 0.1
    y = tmp_path / "config_fingerprint.yml"
    y.write_text("paths:\n  dataset_dir: ./d\norientation:\n  quality_threshold: 0.3\n  margin: 35\n  orient_sigma: 9.0\n"
                 "general:\n  rel_threshold: 0.25\n  block_size: 16\nbinarization:\n  sauv_k: 0.2\n")
    monkeypatch.setenv("FPB200_CONFIG_YAML", str(y))
    monkeypatch.setenv("FPB200_YAML_OVERRIDES", "1")
    cf2 = importlib.reload(cf)
    try:
        assert cf2.active_overrides() == {"post_params": {"quality_threshold": 0.3, "margin": 35}, "rel_threshold": 0.25}
        assert set(cf2.unused_keys()) == {"orientation.orient_sigma", "general.block_size", "binarization.sauv_k"}
        assert cf2.DATASET_DIR.endswith("/d")
        with pytest.raises(ValueError):
            cf2.overrides({"orientation": {"quality_window": 34}})
        assert cf2.overrides({"orientation": {"quality_window": 24}})["post_params"]["quality_window"] == 24   # even: cv2.blur takes it
    finally:
        monkeypatch.delenv("FPB200_CONFIG_YAML")
        monkeypatch.delenv("FPB200_YAML_OVERRIDES")
        importlib.reload(cf)
