"""CPU: the NumPy model of k_nlm_sym's formulation (every patch distance once, added to both pixels of the pair; half-plane
offsets over the image extended by ten pixels) against cv2.fastNlMeansDenoising - the call of fingerprint_preprocess.py:36."""
import os
import sys

import cv2
import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
from nlm_sym_model import nlm_sym, weight_table  # noqa: E402


def test_weight_table_is_the_oracles():
    from oracle.stages import nlm_weight_table
    tab, shift, mult = nlm_weight_table()
    t = weight_table()
    assert shift == 6 and mult == 19096 and (t[:528] == tab[:528]).all() and t[528] == 0 and tab[528] == 0


@pytest.mark.parametrize("kind", ["ridge", "noise", "flat", "edge", "tiny"])
def test_symmetric_half_plane_formulation_equals_opencv(kind):
    from multimodal_biometric_fingerprints_palms_b200.synth import ridge_image
    from oracle import ref_pipeline as rp
    rng = np.random.default_rng(7)
    img = {"ridge": rp.normalize_image(ridge_image(120, 111, seed=5))[30:75, 20:61].copy(),
           "noise": rng.integers(0, 256, (31, 27), dtype=np.uint8),
           "flat": np.full((23, 25), 140, np.uint8),
           "edge": np.concatenate([np.zeros((26, 14), np.uint8), np.full((26, 15), 255, np.uint8)], 1),
           "tiny": rng.integers(0, 256, (22, 21), dtype=np.uint8)}[kind]
    want = cv2.fastNlMeansDenoising(img, None, 10, 7, 21)
    if kind == "ridge":
        assert (want != img).sum() > 100        # live weights beyond the centre offset: the q side matters in this case
    assert np.array_equal(nlm_sym(img), want)
