"""Parity of every CUDA stage against the CPU oracle, through the C ABI (host-buffer stage entry points).
Each stage is fed the ORACLE's input for that stage, so a failure upstream does not cascade.
Run on the B200 box:  python -m pytest tests -m gpu -q"""
import numpy as np
import pytest

from conftest import angle_diff, assert_same, golden_cases, load_golden
from oracle import ref_pipeline as rp

pytestmark = pytest.mark.gpu

from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline, pipeline_for  # noqa: E402
from multimodal_biometric_fingerprints_palms_b200 import synth  # noqa: E402

EXTRA = [("ridge_rand_320x240_s11", lambda: synth.ridge_image(320, 240, seed=11, period=None)),
         ("ridge_rand_320x240_s12", lambda: synth.ridge_image(320, 240, seed=12, period=None, noise_sigma=20.0)),
         ("degraded_512x512_s5", lambda: synth.degraded_image(512, 512, seed=5)),
         ("ridge_200x184_s4", lambda: synth.ridge_image(200, 184, seed=4, period=8.0))]

_cache = {}


def case(name):
    """img + every oracle intermediate for `name` (golden fixture or oracle-computed)."""
    if name in _cache:
        return _cache[name]
    if name in golden_cases():
        g, lists = load_golden(name)
        img = g["img"]
    else:
        img = dict(EXTRA)[name]()
    d = {"img": img}
    d["normalized"] = rp.normalize_image(img)
    d["nlm"], d["denoised"] = rp.denoise_image_parts(d["normalized"])
    det = {}
    d["segmented"], d["mask"] = rp.segment_fingerprint(d["denoised"], det)
    d["seg_detail"] = det
    d["binary"] = rp.binarize(d["segmented"])
    d["orient_blocks"], d["orient_img"], d["reliability"] = rp.compute_orientation_map(d["segmented"], mask=d["mask"])
    d["binary_smooth"] = rp.smooth_fingerprint_skeleton(d["binary"])
    d["gate"] = rp.thinning_gate(d["binary_smooth"], d["reliability"])
    d["skeleton"] = rp.thin_and_clean(d["gate"])
    d["raw"] = rp.extract_minutiae(d["skeleton"])
    pdet = {}
    d["refined"] = rp.postprocess_minutiae([dict(m) for m in d["raw"]], d["skeleton"], d["skeleton"], None, pdet)
    d["post_detail"] = pdet
    # the reference's file hand-off (run_preprocessing.py:137-140 -> extract_features.py:83-92)
    d["skeleton_file"] = rp.skeleton_file_roundtrip(d["skeleton"])
    d["raw_file"] = rp.extract_minutiae(d["skeleton_file"])
    d["refined_file"] = rp.postprocess_minutiae([dict(m) for m in d["raw_file"]], d["skeleton_file"], d["skeleton_file"], None)
    _cache[name] = d
    return d


ALL = golden_cases() + [n for n, _ in EXTRA]


def pipe(a, exact=False):
    """K1..K3 take whole frames (exact handle); the crop stages run through the 64-pixel shape BUCKET of `pipeline_for`
    with the image's size as a ROI - the path the reference-style per-file flow takes."""
    return pipeline_for(a.shape[0], a.shape[1], exact=exact)


@pytest.mark.parametrize("name", ALL)
def test_k1_normalize_bit_exact(name):
    d = case(name)
    assert_same(pipe(d["img"], True).normalize(d["img"])[0], d["normalized"], "normalize_image", f"k1_{name}")


@pytest.mark.parametrize("name", ALL)
def test_k2_denoise_bit_exact(name):
    d = case(name)
    out, nlm = pipe(d["normalized"], True).denoise(d["normalized"], with_nlm=True)
    assert_same(nlm[0], d["nlm"], "fastNlMeansDenoising", f"k2nlm_{name}")
    assert_same(out[0], d["denoised"], "denoise_image", f"k2_{name}")


@pytest.mark.parametrize("name", ALL)
def test_k3_segment_bit_exact(name):
    d = case(name)
    seg, mask, roi = pipe(d["denoised"], True).segment(d["denoised"])
    x0, y0, w, h = (int(v) for v in roi[0])
    want_roi = d["seg_detail"].get("roi") or (0, 0, d["denoised"].shape[1], d["denoised"].shape[0])
    assert (x0, y0, w, h) == tuple(want_roi), f"roi {(x0, y0, w, h)} != {want_roi}"
    assert_same(mask[0, :h, :w], d["mask"], "hull mask", f"k3mask_{name}")
    assert_same(seg[0, :h, :w], d["segmented"], "segmented", f"k3seg_{name}")


@pytest.mark.parametrize("name", ALL)
def test_k4_binarize_bit_exact(name):
    d = case(name)
    assert_same(pipe(d["segmented"]).binarize(d["segmented"])[0], d["binary"], "binarize", f"k4_{name}")


@pytest.mark.parametrize("name", ALL)
def test_k5_orientation_within_1e4(name):
    d = case(name)
    blk, oimg, rel = pipe(d["segmented"]).orientation(d["segmented"], d["mask"])
    # tolerance of the contract (BASELINE.json north_star): 1e-4 relative; orientation compared on the circle
    tol = 1e-4
    e_blk = angle_diff(blk[0], d["orient_blocks"]).max()
    e_img = angle_diff(oimg[0], d["orient_img"]).max()
    e_rel = np.abs(rel[0] - d["reliability"]).max() / max(1e-12, np.abs(d["reliability"]).max())
    assert e_blk <= tol * np.pi and e_img <= tol * np.pi and e_rel <= tol, (e_blk, e_img, e_rel)


@pytest.mark.parametrize("name", ALL)
def test_k6_smooth_bit_exact(name):
    d = case(name)
    assert_same(pipe(d["binary"]).smooth(d["binary"])[0], d["binary_smooth"], "smooth_fingerprint_skeleton", f"k6_{name}")


@pytest.mark.parametrize("name", ALL)
def test_k7_thinning_and_cleaning_bit_exact(name):
    d = case(name)
    sk, gate = pipe(d["binary_smooth"]).thin(d["binary_smooth"], d["reliability"], with_gate=True)
    assert_same(gate[0], d["gate"].astype(np.uint8) * 255, "mask entering skeletonize", f"k7gate_{name}")
    assert_same(sk[0], d["skeleton"], "thinning_and_cleaning", f"k7_{name}")


@pytest.mark.parametrize("name", ALL)
def test_k7b_skeleton_from_shared_mask_bit_exact(name):
    d = case(name)
    g = d["gate"].astype(np.uint8) * 255
    assert_same(pipe(g).skeletonize(g)[0], d["skeleton"], "skeletonize+cleanup", f"k7b_{name}")


@pytest.mark.parametrize("name", ALL)
def test_k8_raw_minutiae_bit_exact(name):
    d = case(name)
    assert pipe(d["skeleton"]).extract_minutiae(d["skeleton"])[0] == d["raw"]


@pytest.mark.parametrize("name", ALL)
def test_skeleton_jpeg_handoff_bit_exact(name):
    """k_jpeg_roundtrip == cv2.imwrite(quality 95) + cv2.imread on the skeleton, and K8 of the decoded file."""
    d = case(name)
    got = pipe(d["skeleton"]).jpeg_roundtrip(d["skeleton"])[0]
    assert_same(got, d["skeleton_file"], "skeleton through the JPEG file", f"jpeg_{name}")
    assert pipe(got).extract_minutiae(got)[0] == d["raw_file"]


def test_jpeg_roundtrip_on_arbitrary_images_bit_exact():
    """the codec restatement on non-skeleton content: noise, gradients, constant, sizes that are not multiples of 8"""
    rng = np.random.default_rng(5)
    for h, w in ((8, 8), (97, 131), (64, 200), (33, 17), (250, 199)):
        imgs = [rng.integers(0, 256, (h, w)).astype(np.uint8), (np.add.outer(np.arange(h), 2 * np.arange(w)) % 256).astype(np.uint8),
                np.full((h, w), 255, np.uint8), (rng.random((h, w)) < 0.1).astype(np.uint8) * 255]
        p_ = FingerprintPipeline(h, w, max_batch=len(imgs))
        got = p_.jpeg_roundtrip(np.stack(imgs))
        for i, im in enumerate(imgs):
            assert_same(got[i], rp.skeleton_file_roundtrip(im), f"jpeg roundtrip {h}x{w} case {i}", f"jpegrt_{h}x{w}_{i}")


@pytest.mark.parametrize("src", ["memory", "file"])
@pytest.mark.parametrize("name", ALL)
def test_k9_postprocess(name, src):
    """K9 on the clean skeleton (the function called in-process) and on the decoded JPEG (what the reference's CLI feeds
    it: grey ringing in the density and orientation maps)."""
    d = case(name)
    skel, raw, want = (d["skeleton"], d["raw"], d["refined"]) if src == "memory" else (d["skeleton_file"], d["raw_file"], d["refined_file"])
    got = pipe(skel).postprocess(skel, [raw])[0]
    assert [(m["x"], m["y"], m["type"]) for m in got] == [(m["x"], m["y"], m["type"]) for m in want]
    for a, b in zip(got, want):
        assert angle_diff(a["orientation"], b["orientation"]) <= 1e-4 * np.pi
        for k in ("quality", "coherence", "angular_stability"):
            assert abs(a[k] - b[k]) <= 1e-4 * max(1e-3, abs(b[k])), (k, a[k], b[k])
        # angle bin of SURVEY 8(c)(iii): 32 bins over [-pi/2, pi/2), equal unless within 1e-3 rad of an edge
        ba, bb = (int(np.floor((v + np.pi / 2) / (np.pi / 32))) for v in (a["orientation"], b["orientation"]))
        edge = abs(((b["orientation"] + np.pi / 2) / (np.pi / 32)) % 1.0 - 0.5) > 0.5 - 1e-3 / (np.pi / 32)
        assert ba == bb or edge


def test_skeletonize_edge_cases_bit_exact():
    """empty, full, single pixel, border lines, 1-px frame, checkerboard, random - table-driven thinning must
    equal the oracle for any input."""
    rng = np.random.default_rng(0)
    h, w = 67, 93                      # not multiples of 32
    imgs = [np.zeros((h, w), bool), np.ones((h, w), bool)]
    a = np.zeros((h, w), bool); a[10, 20] = True; imgs.append(a)
    a = np.zeros((h, w), bool); a[0, :] = a[-1, :] = a[:, 0] = a[:, -1] = True; imgs.append(a)
    a = np.zeros((h, w), bool); a[5:30, 31:34] = True; a[40:43, 5:90] = True; imgs.append(a)
    imgs.append((np.indices((h, w)).sum(0) % 2).astype(bool))
    for p in (0.3, 0.5, 0.7, 0.9):
        imgs.append(rng.random((h, w)) < p)
    p_ = FingerprintPipeline(h, w, max_batch=len(imgs))
    got = p_.skeletonize(np.stack(imgs).astype(np.uint8) * 255)
    for i, m in enumerate(imgs):
        assert_same(got[i], rp.thin_and_clean(m), f"skeletonize case {i}", f"thin_edge_{i}")
        assert p_.extract_minutiae(got[i:i + 1])[0] == rp.extract_minutiae(got[i])


def test_custom_thinning_table_is_data():
    rng = np.random.default_rng(1)
    from oracle.skimage_compat import zhang_suen_table
    tab = zhang_suen_table().copy()
    tab[10] = 3; tab[40] = 3; tab[160] = 3; tab[130] = 3        # add the four "staircase" deletions
    m = rng.random((64, 64)) < 0.6
    p_ = FingerprintPipeline(64, 64)
    p_.set_thin_table(tab)
    assert_same(p_.skeletonize(m.astype(np.uint8) * 255)[0], rp.thin_and_clean(m, tab), "custom table")


def test_k7_component_filters_on_random_masks_bit_exact():
    """remove_small_objects(64) / remove_small_holes(80) / gate / thinning on arbitrary binary inputs (blobs, salt and
    pepper, stripes, empty, full): stresses the run-based component labelling in shared memory."""
    import cv2
    rng = np.random.default_rng(7)
    h, w = 150, 173
    masks = []
    for s in (1.0, 2.0, 4.0):
        masks.append(cv2.GaussianBlur(rng.random((h, w)).astype(np.float32), (0, 0), s) > 0.5)
    for p in (0.02, 0.2, 0.5, 0.8, 0.98):
        masks.append(rng.random((h, w)) < p)
    stripes = np.zeros((h, w), bool); stripes[::3] = True; stripes[:, ::17] = True
    masks += [stripes, np.zeros((h, w), bool), np.ones((h, w), bool)]
    checker = (np.indices((h, w)).sum(0) % 2).astype(bool)
    masks.append(checker)                                   # every pixel its own 4-connected component
    rel_maps = [np.ones((h, w), np.float32), cv2.GaussianBlur(rng.random((h, w)).astype(np.float32), (0, 0), 9.0) * 0.4]
    p_ = FingerprintPipeline(h, w, max_batch=len(masks))
    for rel in rel_maps:
        b = np.stack(masks).astype(np.uint8) * 255
        r = np.stack([rel] * len(masks))
        sk, gate = p_.thin(b, r, with_gate=True)
        for i, m in enumerate(masks):
            want_gate = rp.thinning_gate(b[i], rel)
            assert_same(gate[i], want_gate.astype(np.uint8) * 255, f"gate of mask {i}", f"k7rand_gate_{i}")
            assert_same(sk[i], rp.thin_and_clean(want_gate), f"skeleton of mask {i}", f"k7rand_{i}")


def test_k4_binarize_on_textures_bit_exact():
    """binarize on inputs that are not fingerprints (noise, gradients, blobs): Sauvola / patch Otsu / the fused
    component + opening + reconstruction tail must follow the oracle everywhere."""
    import cv2
    rng = np.random.default_rng(8)
    h, w = 141, 199
    imgs = [rng.integers(0, 256, (h, w)).astype(np.uint8),
            (np.add.outer(np.arange(h), np.arange(w)) % 256).astype(np.uint8),
            np.clip(cv2.GaussianBlur(rng.random((h, w)).astype(np.float32), (0, 0), 3.0) * 900 - 320, 0, 255).astype(np.uint8),
            np.full((h, w), 128, np.uint8)]
    p_ = FingerprintPipeline(h, w, max_batch=len(imgs))
    got = p_.binarize(np.stack(imgs))
    import warnings
    for i, im in enumerate(imgs):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert_same(got[i], rp.binarize(im), f"binarize texture {i}", f"k4tex_{i}")


def test_k9_nms_and_redundancy_standalone():
    """nms_adaptive (with its last-writer-wins quirk) and remove_redundant_oriented_adaptive as callables."""
    from multimodal_biometric_fingerprints_palms_b200.features.post_processing import nms_adaptive, remove_redundant_oriented_adaptive
    dens = np.zeros((100, 100), np.float32) + 0.5
    # three collinear points 5 px apart: only the LOWEST-quality one survives in the reference
    trio = [{"x": 40, "y": 50, "type": "ending", "quality": 0.9}, {"x": 45, "y": 50, "type": "ending", "quality": 0.8},
            {"x": 50, "y": 50, "type": "ending", "quality": 0.7}]
    want = rp.nms_adaptive([dict(m) for m in trio], dens, 8.0)
    assert [m["x"] for m in want] == [50]
    assert nms_adaptive([dict(m) for m in trio], dens, 8.0) == want
    rng = np.random.default_rng(11)
    for t in range(25):
        n = int(rng.integers(1, 120))
        dm = rng.random((120, 90)).astype(np.float32)
        ms = [{"x": int(rng.integers(0, 90)), "y": int(rng.integers(0, 120)), "type": "ending" if rng.random() < .5 else "bifurcation",
               "quality": float(rng.random()), "orientation": float(rng.uniform(-np.pi / 2, np.pi / 2))} for _ in range(n)]
        assert nms_adaptive([dict(m) for m in ms], dm, 8.0) == rp.nms_adaptive([dict(m) for m in ms], dm, 8.0)
        assert (remove_redundant_oriented_adaptive([dict(m) for m in ms], dm) ==
                rp.remove_redundant_oriented_adaptive([dict(m) for m in ms], dm))
    assert nms_adaptive([], dens) == [] and remove_redundant_oriented_adaptive([], dens) == []


@pytest.mark.parametrize("rel_thresh", [0.2, 0.05])
def test_k7_non_default_rel_thresh_bit_exact(rel_thresh):
    """thinning_and_cleaning(rel_thresh=...) - config_fingerprint.yml general.rel_threshold (0.2) as the opt-in override."""
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.fingerprint_preprocess import thinning_and_cleaning
    d = case(ALL[0])
    want = rp.thinning_and_cleaning(d["binary_smooth"], None, d["reliability"], rel_thresh)
    got = thinning_and_cleaning(d["binary_smooth"], None, d["reliability"], rel_thresh=rel_thresh)
    assert_same(got, want, f"thinning_and_cleaning(rel_thresh={rel_thresh})")
    assert not np.array_equal(want, d["skeleton"])                # the parameter really changes the result
    back = thinning_and_cleaning(d["binary_smooth"], None, d["reliability"])          # and the default is restored
    assert_same(back, d["skeleton"], "thinning_and_cleaning default after an override")


def test_k9_yaml_orientation_section_as_params():
    """postprocess_minutiae with the reference's shipped YAML values (orientation.*) == the oracle with the same params."""
    from multimodal_biometric_fingerprints_palms_b200.config import config_fingerprint as cf
    from multimodal_biometric_fingerprints_palms_b200.features.post_processing import postprocess_minutiae
    params = cf.overrides(cf._SHIPPED)["post_params"]
    assert params == {"quality_window": 25, "quality_threshold": 0.25, "coherence_threshold": 0.3, "min_distance": 10.0, "margin": 40}
    d = case(ALL[0])
    sk = d["skeleton_file"]
    want = rp.postprocess_minutiae([dict(m) for m in d["raw_file"]], sk, sk, params)
    got = postprocess_minutiae([dict(m) for m in d["raw_file"]], sk, sk, params)
    assert [(m["x"], m["y"], m["type"]) for m in got] == [(m["x"], m["y"], m["type"]) for m in want]
    for a, b in zip(got, want):
        for k in ("orientation", "quality", "coherence", "angular_stability"):
            assert abs(a[k] - b[k]) <= 1e-4 * max(1.0, abs(b[k])), (k, a[k], b[k])
