"""The C-ABI library loads and exports every symbol include/fpb200.h declares; the product path fails
loudly without a GPU; the product never imports the oracle.  CPU only (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADERS = [os.path.join(ROOT, "include", h) for h in ("fpb200.h", "fpb200_match.h", "fpb200_io.h", "fpb200_unet.h")]
PKG = os.path.join(ROOT, "multimodal_biometric_fingerprints_palms_b200")


def declared_symbols():
    src = "".join(open(h).read() for h in HEADERS)
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fpb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("fpb_create", "fpb_destroy", "fpb_run_host", "fpb_run_device", "fpb_normalize", "fpb_denoise",
                 "fpb_segment", "fpb_binarize", "fpb_orientation", "fpb_smooth", "fpb_thin", "fpb_skeletonize",
                 "fpb_extract_minutiae", "fpb_postprocess", "fpb_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from multimodal_biometric_fingerprints_palms_b200 import _native
    lib = _native.load()
    for s in declared_symbols():
        assert hasattr(lib, s), f"libfpb200.so does not export {s}"
        assert s in _native.SIGNATURES, f"ctypes binding lacks {s}"
    assert lib.fpb_abi_version() == 2


def test_struct_layout_matches_header():
    from multimodal_biometric_fingerprints_palms_b200 import _native
    assert ctypes.sizeof(_native.Minutia) == 48
    assert ctypes.sizeof(_native.PostParams) == 48
    assert ctypes.sizeof(_native.MatchParams) == 48      # fpb_match_params
    assert ctypes.sizeof(_native.MatchResult) == 48      # fpb_match_result


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline, FpbError
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.fingerprint_preprocess import preprocess_fingerprint
    with pytest.raises(FpbError, match="no CPU path"):
        FingerprintPipeline(320, 240)
    with pytest.raises(RuntimeError, match="preprocess_fingerprint failed"):
        preprocess_fingerprint(np.zeros((320, 240), np.uint8))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "import cv2" not in txt or f in ("fingerprint_preprocess.py", "orientation.py", "extract_features.py",
                                                        "run_preprocessing.py", "drivers.py", "selfcheck.py", "inference.py"), \
                    f"{f}: cv2 is for file I/O / debug drawing (and the scikit-image deployment self-check) only"
