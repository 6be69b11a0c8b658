"""CPU: the NumPy model of the tensor-core NLM kernel (tests/emul/nlm_mma_model.py mirrors csrc/k_nlm_mma.cu byte for
byte: tile, im2col operand layout, Gram GEMM, sign test, survivor arithmetic, lane / quadrant / chunk maps) against
cv2.fastNlMeansDenoising - the call of fingerprint_preprocess.py:36.  Pins the formulation's arithmetic without a GPU."""
import os
import sys

import cv2
import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
from nlm_mma_model import nlm_image, weight_table  # noqa: E402


def test_weight_table_matches_opencv_constants():
    t = weight_table()
    assert t[0] == 19096 and t[527] > 0 and t[528] == 0 and (np.diff(t[:528]) <= 0).all()


@pytest.mark.parametrize("kind", ["ridge", "noise", "flat", "edge"])
def test_gram_formulation_equals_opencv(kind):
    from multimodal_biometric_fingerprints_palms_b200.synth import ridge_image
    from oracle import ref_pipeline as rp
    rng = np.random.default_rng(4)
    img = {"ridge": rp.normalize_image(ridge_image(120, 111, seed=3))[40:73, 30:59].copy(),     # 33 x 29: partial blocks on both axes
           "noise": rng.integers(0, 256, (19, 23), dtype=np.uint8),
           "flat": np.full((17, 9), 201, np.uint8),                                              # every pair survives the sign test
           "edge": np.concatenate([np.zeros((16, 5), np.uint8), np.full((16, 6), 255, np.uint8)], 1)}[kind]   # N(q) spans its whole range
    got, st = nlm_image(img)
    want = cv2.fastNlMeansDenoising(img, None, 10, 7, 21)
    assert np.array_equal(got, want), f"{(got != want).sum()} pixels differ"
    assert st["survivors"] > 0
