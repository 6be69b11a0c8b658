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


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_reproduces_the_references_file_handoff(name):
    """golden `*_file` entries = the reference's OWN process_image (extract_features.py:74-108) run on the
    <base>_skeleton.jpg that cv2.imwrite produced: the oracle's default flow and the restated codec must equal them."""
    from oracle.jpeg_fdct import jpeg_roundtrip
    g, lists = load_golden(name)
    assert_same(rp.skeleton_file_roundtrip(g["skeleton"]), g["skeleton_file"], "cv2 JPEG hand-off")
    assert_same(jpeg_roundtrip(g["skeleton"]), g["skeleton_file"], "restated quality-95 codec")
    assert rp.extract_minutiae(g["skeleton_file"]) == lists["raw_minutiae_file"]
    got = rp.postprocess_minutiae([dict(m) for m in lists["raw_minutiae_file"]], g["skeleton_file"], g["skeleton_file"], None)
    assert [(m["x"], m["y"], m["type"]) for m in got] == [(m["x"], m["y"], m["type"]) for m in lists["minutiae_file"]]
    for a, b in zip(got, lists["minutiae_file"]):
        for k in ("orientation", "quality", "coherence", "angular_stability"):
            assert abs(a[k] - b[k]) <= 1e-6 * max(1.0, abs(b[k])), (k, a[k], b[k])


def test_pipeline_end_to_end_matches_golden():
    g, lists = load_golden(golden_cases()[0])
    res = rp.enhance_to_minutiae(g["img"])
    assert_same(res["skeleton"], g["skeleton"], "pipeline skeleton")
    assert_same(res["skeleton_file"], g["skeleton_file"], "pipeline skeleton file")
    assert res["raw_minutiae"] == lists["raw_minutiae_file"]
    assert [(m["x"], m["y"]) for m in res["minutiae"]] == [(m["x"], m["y"]) for m in lists["minutiae_file"]]
    mem = rp.enhance_to_minutiae(g["img"], handoff="memory")
    assert mem["raw_minutiae"] == lists["raw_minutiae"]
    assert [(m["x"], m["y"]) for m in mem["minutiae"]] == [(m["x"], m["y"]) for m in lists["minutiae"]]
