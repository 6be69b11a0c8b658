"""End-to-end CUDA pipeline (fused K1..K9, batched) against the oracle: agreement figures are REPORTED
(BASELINE.json north_star) and must be essentially perfect; batched results must equal per-image results."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, assert_same, golden_cases, load_golden
from oracle import ref_pipeline as rp

pytestmark = pytest.mark.gpu

from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline  # noqa: E402
from multimodal_biometric_fingerprints_palms_b200 import synth  # noqa: E402


def _match_rate(got, want):
    a = {(m["x"], m["y"], m["type"]) for m in got}
    b = {(m["x"], m["y"], m["type"]) for m in want}
    return len(a & b) / max(1, len(a | b))


def test_fused_batch_matches_oracle_and_reports_agreement():
    n = 12
    imgs = synth.ridge_batch(n, 320, 240, first_seed=100)
    p = FingerprintPipeline(320, 240, max_batch=n)
    p.run(imgs)
    planes = {k: p.fetch(k) for k in ("normalized", "denoised", "mask", "binary", "skeleton")}
    rows = []
    for i in range(n):
        ref = rp.enhance_to_minutiae(imgs[i])
        x0, y0, w, h = p.roi(i)
        same_crop = ref["skeleton"].shape == (h, w)
        row = {"image": i, "crop_equal": bool(same_crop)}
        row["normalized"] = float((planes["normalized"][i] == ref["normalized"]).mean())
        row["denoised"] = float((planes["denoised"][i] == ref["denoised"]).mean())
        if same_crop:
            for k in ("mask", "binary", "skeleton"):
                row[k] = float((planes[k][i, :h, :w] == ref[k]).mean())
        row["raw_match"] = _match_rate(p.raw_minutiae(i), ref["raw_minutiae"])
        row["refined_match"] = _match_rate(p.minutiae(i), ref["minutiae"])
        rows.append(row)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "e2e_agreement.json"), "w") as f:
        json.dump(rows, f, indent=1)
    for r in rows:
        assert r["crop_equal"], r
        assert r["normalized"] == 1.0 and r["denoised"] == 1.0, r
        assert r["mask"] == 1.0, r
        assert r["binary"] >= 0.999 and r["skeleton"] >= 0.999, r
        assert r["raw_match"] >= 0.95 and r["refined_match"] >= 0.9, r


def test_batch_equals_single_image_runs():
    imgs = synth.ridge_batch(5, 320, 240, first_seed=7)
    pb = FingerprintPipeline(320, 240, max_batch=5)
    pb.run(imgs)
    skb = pb.fetch("skeleton")
    lists = [pb.minutiae(i) for i in range(5)]
    rois = [pb.roi(i) for i in range(5)]
    ps = FingerprintPipeline(320, 240, max_batch=1)
    for i in range(5):
        ps.run(imgs[i])
        assert ps.roi(0) == rois[i]
        x0, y0, w, h = rois[i]
        assert_same(ps.fetch("skeleton")[0, :h, :w], skb[i, :h, :w], f"skeleton {i}")
        assert ps.minutiae(0) == lists[i]


def test_reference_api_mirror_single_image():
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.fingerprint_preprocess import preprocess_fingerprint
    from multimodal_biometric_fingerprints_palms_b200.features.extract_features import extract_minutiae
    from multimodal_biometric_fingerprints_palms_b200.features.post_processing import postprocess_minutiae
    g, lists = load_golden(golden_cases()[0])
    res = preprocess_fingerprint(g["img"])
    assert list(res) == ["normalized", "denoised", "segmented", "mask", "binary", "skeleton", "orientation_vis"]
    assert_same(res["normalized"], g["normalized"], "normalized")
    assert_same(res["denoised"], g["denoised"], "denoised")
    assert_same(res["segmented"], g["segmented"], "segmented")
    assert_same(res["mask"], g["mask"], "mask")
    assert res["orientation_vis"].shape == g["segmented"].shape + (3,)
    raw = extract_minutiae(g["skeleton"])
    assert raw == lists["raw_minutiae"]
    refined = postprocess_minutiae(raw, g["skeleton"], g["skeleton"], None)
    assert [(m["x"], m["y"], m["type"]) for m in refined] == [(m["x"], m["y"], m["type"]) for m in lists["minutiae"]]
    assert all(r is m for r in refined for m in raw if (m["x"], m["y"]) == (r["x"], r["y"]) and "quality" in m)
    with pytest.raises(RuntimeError, match="preprocess_fingerprint failed"):
        preprocess_fingerprint(np.zeros((10, 10, 3), np.uint8))
    assert postprocess_minutiae([], g["skeleton"]) == []


PLANES = ("normalized", "denoised", "mask", "binary", "binary_smooth", "gate", "skeleton", "skeleton_file")


def _floats_close(got, want):
    """refined records: orientation on the circle and the three scores within the 1e-4 of the contract"""
    from conftest import angle_diff
    for a, b in zip(got, want):
        if angle_diff(a["orientation"], b["orientation"]) > 1e-4 * np.pi:
            return False
        for k in ("quality", "coherence", "angular_stability"):
            if abs(a[k] - b[k]) > 1e-4 * max(1e-3, abs(b[k])):
                return False
    return True


def _e2e_rows(imgs, handoff="file"):
    """fused batched run vs the oracle image by image; `handoff="file"` = the reference's CLI flow (skeleton through the
    quality-95 JPEG between the two stages), the library's default"""
    n, H, W = imgs.shape
    p = FingerprintPipeline(H, W, max_batch=n, handoff=handoff)
    p.run(imgs)
    names = [k for k in PLANES if handoff == "file" or k != "skeleton_file"]
    planes = {k: p.fetch(k) for k in names}
    rows = []
    for i in range(n):
        ref = rp.enhance_to_minutiae(imgs[i], handoff=handoff)
        ref["gate"] = rp.thinning_gate(ref["binary_smooth"], ref["reliability"]).astype(np.uint8) * 255
        x0, y0, w, h = p.roi(i)
        row = {"image": i, "roi": [x0, y0, w, h], "crop_equal": ref["skeleton"].shape == (h, w), "planes": names}
        row["normalized"] = float((planes["normalized"][i] == ref["normalized"]).mean())
        row["denoised"] = float((planes["denoised"][i] == ref["denoised"]).mean())
        if row["crop_equal"]:
            for k in names[2:]:
                row[k] = float((planes[k][i, :h, :w] == ref[k]).mean())
        row["raw_equal"] = p.raw_minutiae(i) == ref["raw_minutiae"]
        got, want = p.minutiae(i), ref["minutiae"]
        row["refined_equal"] = [(m["x"], m["y"], m["type"]) for m in got] == [(m["x"], m["y"], m["type"]) for m in want]
        row["refined_floats_close"] = row["refined_equal"] and _floats_close(got, want)
        rows.append(row)
    return rows


def _assert_rows_exact(rows):
    for r in rows:
        assert r["crop_equal"], r
        for k in r["planes"]:
            assert r[k] == 1.0, r
        assert r["raw_equal"] and r["refined_equal"] and r["refined_floats_close"], r


def test_in_memory_handoff_option():
    """`handoff="memory"`: K8/K9 on the clean in-memory skeleton (the two reference functions called in one process)"""
    imgs = synth.ridge_batch(3, 320, 240, first_seed=300)
    _assert_rows_exact(_e2e_rows(imgs, handoff="memory"))
    # and the two hand-offs really differ (the JPEG's ringing changes density / coherence): guard against a silent no-op
    a = _e2e_rows(imgs[:1], "file"); b = _e2e_rows(imgs[:1], "memory")
    p = FingerprintPipeline(320, 240, max_batch=1)
    p.run(imgs[0]); x0, y0, w, h = p.roi(0)
    assert (p.fetch("skeleton_file")[0, :h, :w] > 0).sum() > 2 * (p.fetch("skeleton")[0, :h, :w] > 0).sum()
    assert a[0]["raw_equal"] and b[0]["raw_equal"]


def test_golden_file_handoff_matches_the_references_own_json():
    """tests/golden holds what the reference's own process_image wrote for <base>_skeleton.jpg (oracle/make_golden.py):
    the fused default run must reproduce that JSON - membership, order, floats."""
    for name in golden_cases():
        g, lists = load_golden(name)
        H, W = g["img"].shape
        p = FingerprintPipeline(H, W, max_batch=1)
        p.run(g["img"])
        x0, y0, w, h = p.roi(0)
        assert_same(p.fetch("skeleton_file")[0, :h, :w], g["skeleton_file"], f"{name}: skeleton as read from the JPEG")
        assert p.raw_minutiae(0) == lists["raw_minutiae_file"]
        got, want = p.minutiae(0), lists["minutiae_file"]
        assert [(m["x"], m["y"], m["type"]) for m in got] == [(m["x"], m["y"], m["type"]) for m in want], name
        assert _floats_close(got, want), name


def test_config2_degraded_512x512_batch():
    """BASELINE configs[2]: NIST-shape degraded 512x512 inputs (heavy noise, saturated band, blobs)."""
    imgs = np.stack([synth.degraded_image(512, 512, seed=s) for s in (21, 22, 23)])
    _assert_rows_exact(_e2e_rows(imgs))


def test_config4_highres_1024x1024():
    """BASELINE configs[4] shape: 1024x1024, ridge period 18 - also exercises the large-image paths (bit images in
    global scratch, per-pixel union-find) that the 320x240 batch never takes."""
    img = synth.ridge_image(1024, 1024, seed=31, period=18.0, noise_sigma=12.0)
    _assert_rows_exact(_e2e_rows(img[None]))


def test_highres_1024x1024_pure_noise():
    """1024x1024 uniform noise through the whole path: the large-image kernels on a dense, irregular skeleton; the raw
    list must arrive complete (no cap: fpb_raw_capacity is sized from H*W and overflow is an error)."""
    import warnings
    img = np.random.default_rng(77).integers(0, 256, (1024, 1024)).astype(np.uint8)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _assert_rows_exact(_e2e_rows(img[None]))


def test_raw_minutiae_overflow_is_an_error_not_a_truncation():
    """more crossing-number minutiae than the handle can hold: the reference keeps them all, so the library refuses"""
    from multimodal_biometric_fingerprints_palms_b200 import FpbError
    h = w = 96
    sk = np.zeros((h, w), np.uint8)
    sk[1:-1:2, 1:-2:3] = 255; sk[1:-1:2, 2:-1:3] = 255          # two-pixel dashes: every pixel is a ridge ending
    want = rp.extract_minutiae(sk)
    p = FingerprintPipeline(h, w, max_batch=1)
    assert p.raw_capacity == 2048 and len(want) > p.raw_capacity
    with pytest.raises(FpbError, match="raw minutiae"):
        p.extract_minutiae(sk)
    with pytest.raises(FpbError, match="raw minutiae"):
        p.postprocess(sk, [want])
    big = FingerprintPipeline(192, 192, max_batch=1)             # capacity 4608: the same dashes now fit, complete
    pad = np.zeros((192, 192), np.uint8); pad[:h, :w] = sk
    assert big.extract_minutiae(pad)[0] == rp.extract_minutiae(pad)


def test_crop_shapes_share_bucketed_handles():
    """the reference-style per-file flow (extract_features.process_image on crops of data-dependent size) must not create
    a workspace per crop size: 100 different shapes -> a handful of handles, results identical to exact-size handles"""
    from multimodal_biometric_fingerprints_palms_b200.pipeline import handle_cache_size
    from multimodal_biometric_fingerprints_palms_b200.features.extract_features import extract_minutiae
    from multimodal_biometric_fingerprints_palms_b200.features.post_processing import postprocess_minutiae
    res = rp.preprocess_fingerprint(synth.ridge_image(320, 240, seed=5))
    sk_full = rp.skeleton_file_roundtrip(res["skeleton"])
    rng = np.random.default_rng(9)
    before = handle_cache_size()
    for t in range(100):
        h, w = int(rng.integers(200, sk_full.shape[0] + 1)), int(rng.integers(150, sk_full.shape[1] + 1))
        sk = np.ascontiguousarray(sk_full[:h, :w])
        raw = extract_minutiae(sk)
        assert raw == rp.extract_minutiae(sk)
        got = postprocess_minutiae(raw, sk, sk, None)
        want = rp.postprocess_minutiae([dict(m) for m in rp.extract_minutiae(sk)], sk, sk, None)
        assert [(m["x"], m["y"], m["type"]) for m in got] == [(m["x"], m["y"], m["type"]) for m in want], (h, w)
        assert _floats_close(got, want)
    assert handle_cache_size() - before <= 6, handle_cache_size()


def test_selfcheck_against_a_scikit_image_install():
    """where scikit-image is importable the first handle compares binarize / thinning_and_cleaning with the real package
    and raises on any differing pixel.  Here the package is played by oracle/ref_shim (the restated functions): the
    check must pass with the built-in table and must FAIL LOUDLY when the 'installed' skeletonize uses another table."""
    import sys
    from multimodal_biometric_fingerprints_palms_b200 import selfcheck
    shim = os.path.join(ROOT, "oracle", "ref_shim")
    sys.path.insert(0, shim)
    try:
        for m in [k for k in sys.modules if k == "skimage" or k.startswith("skimage.")]:
            del sys.modules[m]
        rep = selfcheck.check(lambda h, w: FingerprintPipeline(h, w, max_batch=1))
        assert rep == {"binarize_px": 0, "thin_px": 0, "cases": 3}, rep
        from oracle import skimage_compat as sc
        tab = sc.zhang_suen_table().copy(); tab[10] = tab[40] = tab[130] = tab[160] = 3
        orig = sc.skeletonize
        import skimage.morphology as mo
        mo.skeletonize = lambda m: orig(m, tab)                  # an install whose table has the staircase deletions
        try:
            rep2 = selfcheck.check(lambda h, w: FingerprintPipeline(h, w, max_batch=1))
            assert rep2["thin_px"] > 0 and rep2["binarize_px"] == 0, rep2
            selfcheck._state["done"] = False
            os.environ["FPB200_SELFCHECK"] = "1"
            with pytest.raises(selfcheck.SkimageParityError, match="FPB200_THIN_TABLE"):
                FingerprintPipeline(64, 64)
            # ... and the documented remedy (the install's table as data) makes the same check pass
            tf = os.path.join(ROOT, "gpurun_out", "staircase_table.txt")
            os.makedirs(os.path.dirname(tf), exist_ok=True)
            np.savetxt(tf, tab.reshape(16, 16), fmt="%d")
            os.environ["FPB200_THIN_TABLE"] = tf
            selfcheck._state["done"] = False
            FingerprintPipeline(64, 64)
            assert selfcheck._state["result"] == {"binarize_px": 0, "thin_px": 0, "cases": 3}
        finally:
            mo.skeletonize = orig
            os.environ.pop("FPB200_SELFCHECK", None); os.environ.pop("FPB200_THIN_TABLE", None)
    finally:
        sys.path.remove(shim)
        for m in [k for k in sys.modules if k == "skimage" or k.startswith("skimage.")]:
            del sys.modules[m]
        selfcheck._state["done"] = True


def test_transposed_and_odd_shapes():
    for shape, seed in (((240, 320), 41), ((200, 184), 42), ((333, 251), 43)):
        img = synth.ridge_image(shape[0], shape[1], seed=seed, period=8.0)
        _assert_rows_exact(_e2e_rows(img[None]))


@pytest.mark.parametrize("kind", ["black", "white", "constant", "noise", "half"])
def test_degenerate_inputs_match_oracle(kind):
    """No-contour / constant / pure-noise inputs: the CUDA path must follow the reference's fall-backs
    (fingerprint_preprocess.py:113-118) and never hang."""
    rng = np.random.default_rng(3)
    H, W = 160, 128
    img = {"black": np.zeros((H, W), np.uint8), "white": np.full((H, W), 255, np.uint8),
           "constant": np.full((H, W), 97, np.uint8), "noise": rng.integers(0, 256, (H, W)).astype(np.uint8),
           "half": np.concatenate([np.full((H, W // 2), 40, np.uint8), np.full((H, W - W // 2), 220, np.uint8)], 1)}[kind]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _assert_rows_exact(_e2e_rows(img[None]))


def test_file_drivers_catalog_to_json(tmp_path):
    """run_preprocessing -> extract_features.main on disk (rows D1/D2): same file names / directory mirroring as the
    reference, JSON equal to the oracle's flow through the same JPEG hand-off."""
    import cv2, json as js
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.run_preprocessing import run_preprocessing
    from multimodal_biometric_fingerprints_palms_b200.features import extract_features as ef
    src = tmp_path / "sorted" / "cluster_0"
    src.mkdir(parents=True)
    imgs = {f"{u}_1_{k}": synth.ridge_image(320, 240, seed=900 + 10 * u + k, period=None) for u in (1, 2) for k in (1, 2)}
    for name, im in imgs.items():
        cv2.imwrite(str(src / f"{name}.png"), im)
    out = tmp_path / "processed"
    assert run_preprocessing(str(tmp_path / "sorted"), str(out), max_workers=2) == 4
    ef.main(input_base=str(out / "enhanced"), output_base=str(out / "minutiae"), max_workers=2)
    for name, im in imgs.items():
        sk_path = out / "enhanced" / "cluster_0" / f"{name}_skeleton.jpg"
        assert sk_path.exists() and (out / "enhanced" / "cluster_0" / f"{name}_enhanced.jpg").exists()
        got = js.load(open(out / "minutiae" / "cluster_0" / f"{name}_minutiae.json"))
        skel = cv2.imread(str(sk_path), cv2.IMREAD_GRAYSCALE)                     # the reference's lossy hand-off
        ref = rp.enhance_to_minutiae(im)
        assert np.array_equal(skel > 127, ref["skeleton"] > 127)
        raw = rp.extract_minutiae(skel)
        want = rp.postprocess_minutiae([dict(m) for m in raw], skel, skel, None)
        assert [(m["x"], m["y"], m["type"]) for m in got] == [(m["x"], m["y"], m["type"]) for m in want]
        assert set(got[0]) == {"x", "y", "type", "orientation", "quality", "coherence", "angular_stability"} if got else True
    with pytest.raises(RuntimeError, match="Nessuna immagine"):
        run_preprocessing(str(tmp_path / "empty_dir_that_has_no_images_" ), str(out)) if (tmp_path / "empty_dir_that_has_no_images_").mkdir() is None else None


def test_thread_pool_callers_like_the_reference_driver():
    """The reference calls preprocess_fingerprint / extract from ThreadPoolExecutor workers (run_preprocessing.py:154,
    extract_features.py:130): handles are per thread, results must not depend on the interleaving."""
    from concurrent.futures import ThreadPoolExecutor
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.fingerprint_preprocess import preprocess_fingerprint
    from multimodal_biometric_fingerprints_palms_b200.features.extract_features import extract_minutiae
    from multimodal_biometric_fingerprints_palms_b200.features.post_processing import postprocess_minutiae
    imgs = [synth.ridge_image(320, 240, seed=700 + i, period=None) for i in range(8)]

    def work(img):
        res = preprocess_fingerprint(img)
        raw = extract_minutiae(res["skeleton"])
        return res["skeleton"], postprocess_minutiae(raw, res["skeleton"], res["skeleton"], None)

    serial = [work(im) for im in imgs]
    with ThreadPoolExecutor(max_workers=4) as ex:
        threaded = list(ex.map(work, imgs))
    for (s0, m0), (s1, m1) in zip(serial, threaded):
        assert np.array_equal(s0, s1) and m0 == m1
    ref = rp.enhance_to_minutiae(imgs[0])
    assert np.array_equal(serial[0][0], ref["skeleton"])


def test_argument_errors_are_loud():
    from multimodal_biometric_fingerprints_palms_b200 import FpbError
    p = FingerprintPipeline(64, 48, max_batch=2)
    with pytest.raises(ValueError):
        p.run(np.zeros((3, 64, 48), np.uint8))              # more than max_batch
    with pytest.raises(ValueError):
        p.run(np.zeros((1, 48, 64), np.uint8))              # wrong shape
    with pytest.raises(TypeError):
        p.run(np.zeros((1, 64, 48), np.float32))
    with pytest.raises(FpbError):
        FingerprintPipeline(64, 48, max_batch=1).fetch("skeleton")        # nothing ran yet
    with pytest.raises(FpbError):
        FingerprintPipeline(2048, 2048)                     # beyond the thinning kernel's image size
    with pytest.raises(FpbError):
        FingerprintPipeline(64, 48, device=99)
    p.set_post_params({"max_minutiae": 5, "margin": 10})
    p.set_post_params(None)
    with pytest.raises(FpbError):
        p.set_post_params({"quality_window": 34})            # the density kernel's tile holds windows up to 33 x 33


def test_fused_directory_driver_with_resume(tmp_path):
    import cv2, json as js
    from multimodal_biometric_fingerprints_palms_b200.drivers import run_directory
    src = tmp_path / "in" / "cluster_3"
    src.mkdir(parents=True)
    imgs = {f"7_1_{k}": synth.ridge_image(320, 240, seed=950 + k, period=None) for k in range(3)}
    for name, im in imgs.items():
        cv2.imwrite(str(src / f"{name}.bmp"), im)
    out = tmp_path / "out"
    st = run_directory(str(tmp_path / "in"), str(out), batch=2)
    assert {k: st[k] for k in ("found", "processed", "skipped", "unreadable", "gpu_decoded")} == \
        {"found": 3, "processed": 3, "skipped": 0, "unreadable": 0, "gpu_decoded": 0}          # BMP inputs: read with cv2
    for name, im in imgs.items():
        got = js.load(open(out / "minutiae" / "cluster_3" / f"{name}_minutiae.json"))
        # the oracle run as the reference's TWO stages through cv2.imencode / imdecode, floats included
        ref = rp.preprocess_fingerprint(im)
        ok, buf = cv2.imencode(".jpg", ref["skeleton"])
        skel = cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE)
        want = rp.postprocess_minutiae([dict(m) for m in rp.extract_minutiae(skel)], skel, skel, None)
        assert [(m["x"], m["y"], m["type"]) for m in got] == [(m["x"], m["y"], m["type"]) for m in want]
        assert _floats_close(got, want)
        # and the skeleton file the driver wrote is the file the second stage of the reference would have read
        assert np.array_equal(cv2.imread(str(out / "enhanced" / "cluster_3" / f"{name}_skeleton.jpg"), cv2.IMREAD_GRAYSCALE), skel)
    st2 = run_directory(str(tmp_path / "in"), str(out), batch=2)
    assert st2["processed"] == 0 and st2["skipped"] == 3


def test_full_size_batch_1480_properties():
    """BASELINE configs[1] at full size (1480 x 320x240): size-independent properties instead of 1480 oracle runs -
    position independence across the whole batch (incl. the two-stream split at image 740), equality with single-image
    runs, crossing-number extraction of the skeleton plane == the raw lists, and idempotence of thinning + clean-up."""
    distinct, reps = 37, 40
    base = synth.ridge_batch(distinct, 320, 240, first_seed=4200)
    imgs = np.stack([base[i % distinct] for i in range(distinct * reps)])
    assert len(imgs) == 1480
    p = FingerprintPipeline(320, 240, max_batch=1480)
    p.run(imgs)
    skel = p.fetch("skeleton"); skel_file = p.fetch("skeleton_file")
    rois = [p.roi(i) for i in range(1480)]
    mins = [p.minutiae(i) for i in range(1480)]
    raws = [p.raw_minutiae(i) for i in range(1480)]
    for i in range(distinct, 1480):                       # every copy equals the first copy of the same print
        j = i % distinct
        assert rois[i] == rois[j] and mins[i] == mins[j] and raws[i] == raws[j], i
        assert np.array_equal(skel[i], skel[j]), i
    q = FingerprintPipeline(320, 240, max_batch=1)
    for j in (0, 17, 36):                                 # and a batch of one gives the same answer
        q.run(imgs[j])
        assert q.roi(0) == rois[j] and q.minutiae(0) == mins[j]
        assert np.array_equal(q.fetch("skeleton")[0], skel[j])
    # K8 on the skeleton plane reproduces the raw lists; thinning a finished skeleton changes nothing
    for j in range(distinct):
        x0, y0, w, h = rois[j]
        s = np.ascontiguousarray(skel[j, :h, :w])
        r = FingerprintPipeline(h, w, max_batch=1)
        assert r.extract_minutiae(np.ascontiguousarray(skel_file[j, :h, :w]))[0] == raws[j]
        assert np.array_equal(r.jpeg_roundtrip(s)[0], skel_file[j, :h, :w])
        assert np.array_equal(r.skeletonize(s)[0], s)
        r.close()
    assert sum(len(m) for m in mins[:distinct]) > 5 * distinct


@pytest.mark.parametrize("shape,seed,period", [((131, 97), 61, 7.0), ((257, 129), 62, 9.0), ((401, 163), 63, 10.0),
                                               ((96, 352), 64, 8.0), ((145, 146), 65, 6.0), ((512, 301), 66, 11.0)])
def test_more_odd_shapes_end_to_end(shape, seed, period):
    """widths / heights that are not multiples of 4, 8, 16 or 32 (vector paths, tile edges, block-grid remainders)"""
    img = synth.ridge_image(shape[0], shape[1], seed=seed, period=period)
    _assert_rows_exact(_e2e_rows(img[None]))


def test_fuzz_random_shapes_and_split_batch():
    """tools/fuzz_shapes.py as a test (VERDICT r1 task 3): 16 random shapes through the fused run against the oracle - exact
    planes and minutiae lists - and an odd-shaped batch large enough for the two-stream split against single-image runs."""
    rng = np.random.default_rng(2024)
    for t in range(16):
        h, w = int(rng.integers(96, 430)), int(rng.integers(96, 430))
        img = synth.ridge_image(h, w, seed=1000 + t, period=float(rng.uniform(6, 12)))
        p = FingerprintPipeline(h, w, max_batch=1)
        p.run(img)
        ref = rp.enhance_to_minutiae(img)
        x0, y0, cw, ch = p.roi(0)
        assert ref["skeleton"].shape == (ch, cw), (h, w)
        for k in ("mask", "binary", "binary_smooth", "skeleton"):
            assert_same(p.fetch(k)[0, :ch, :cw], ref[k], f"{k} at {h}x{w}")
        assert p.raw_minutiae(0) == ref["raw_minutiae"], (h, w)
        assert [(m["x"], m["y"], m["type"]) for m in p.minutiae(0)] == [(m["x"], m["y"], m["type"]) for m in ref["minutiae"]], (h, w)
        p.close()
    h, w, n = 203, 137, 70
    imgs = np.stack([synth.ridge_image(h, w, seed=3000 + i % 7, period=8.0) for i in range(n)])
    pb = FingerprintPipeline(h, w, max_batch=n)
    pb.run(imgs)
    sk = pb.fetch("skeleton")
    p1 = FingerprintPipeline(h, w, max_batch=1)
    for i in range(7):
        p1.run(imgs[i])
        one = p1.fetch("skeleton")[0]
        for j in range(i, n, 7):
            assert pb.roi(j) == p1.roi(0) and pb.minutiae(j) == p1.minutiae(0) and np.array_equal(sk[j], one), (i, j)


def test_async_host_entry_on_two_handles_equals_blocking_runs():
    """fpb_run_host_async / fpb_wait: two handles used alternately give the results of blocking runs, batch after batch"""
    batches = [synth.ridge_batch(70, 320, 240, first_seed=200 + 70 * k) for k in range(4)]
    a, b = FingerprintPipeline(320, 240, max_batch=70), FingerprintPipeline(320, 240, max_batch=70)
    got = []
    pair = (a, b)
    pair[0].run_async(batches[0])
    for k in range(1, 4):
        pair[k & 1].run_async(batches[k])
        pair[(k - 1) & 1].wait()
        got.append([pair[(k - 1) & 1].minutiae(i) for i in range(70)])
    pair[1].wait()
    got.append([pair[1].minutiae(i) for i in range(70)])
    ref = FingerprintPipeline(320, 240, max_batch=70)
    for k in range(4):
        ref.run(batches[k])
        assert got[k] == [ref.minutiae(i) for i in range(70)], k
    roi, rc, oc, rec = ref.result_block()
    assert oc.tolist() == [len(ref.minutiae(i)) for i in range(70)] and roi[3].tolist() == list(ref.roi(3))
    assert [int(v) for v in rec[5, :oc[5]]["x"]] == [m["x"] for m in ref.minutiae(5)]
