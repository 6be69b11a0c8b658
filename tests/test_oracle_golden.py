"""The oracle against the frozen outputs of the reference's own modules (tests/golden, made by
oracle/make_golden.py in the build container).  CPU only."""
import numpy as np
import pytest

from conftest import assert_same, golden_cases, load_golden
from oracle import ref_pipeline as rp


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_reproduces_reference_stages(name):
    g, lists = load_golden(name)
    assert_same(rp.normalize_image(g["img"]), g["normalized"], "normalize_image")
    assert_same(rp.denoise_image(g["normalized"]), g["denoised"], "denoise_image")
    seg, mask = rp.segment_fingerprint(g["denoised"])
    assert_same(seg, g["segmented"], "segmented")
    assert_same(mask, g["mask"], "mask")
    assert_same(rp.binarize(g["segmented"]), g["binary"], "binarize")
    blk, oimg, rel = rp.compute_orientation_map(g["segmented"], mask=g["mask"])
    # float maps: same libraries, but allow for SIMD-dispatch differences between hosts
    np.testing.assert_allclose(blk, g["orient_blocks"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(oimg, g["orient_img"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(rel, g["reliability"], rtol=1e-5, atol=1e-6)
    assert_same(rp.smooth_fingerprint_skeleton(g["binary"]), g["binary_smooth"], "smooth")
    assert_same(rp.thinning_and_cleaning(g["binary_smooth"], g["orient_img"], g["reliability"]), g["skeleton"], "skeleton")
    assert rp.extract_minutiae(g["skeleton"]) == lists["raw_minutiae"]
    got = rp.postprocess_minutiae([dict(m) for m in lists["raw_minutiae"]], g["skeleton"], g["skeleton"], None)
    assert [(m["x"], m["y"], m["type"]) for m in got] == [(m["x"], m["y"], m["type"]) for m in lists["minutiae"]]
    for a, b in zip(got, lists["minutiae"]):
        for k in ("orientation", "quality", "coherence", "angular_stability"):
            assert abs(a[k] - b[k]) <= 1e-6 * max(1.0, abs(b[k])), (k, a[k], b[k])


def test_pipeline_end_to_end_matches_golden():
    g, lists = load_golden(golden_cases()[0])
    res = rp.enhance_to_minutiae(g["img"])
    assert_same(res["skeleton"], g["skeleton"], "pipeline skeleton")
    assert res["raw_minutiae"] == lists["raw_minutiae"]
