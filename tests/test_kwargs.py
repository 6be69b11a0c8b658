"""Non-default keyword arguments of compute_orientation_map (orientation.py:9-14) and smooth_fingerprint_skeleton
(fingerprint_preprocess.py:141-144).  tests/golden/kwargs_128x112.npz holds the outputs of the reference's OWN functions
(oracle/make_golden_args.py, build container); the CPU test pins the oracle against them, the GPU tests compare the CUDA
path - called through the reference-named Python functions, i.e. through fpb_orientation_ex / fpb_smooth_ex of the C ABI -
with the goldens: K6 bit-exact, K5 within the 1e-4 contract."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, angle_diff, assert_same

NAME = "kwargs_128x112"


def _load():
    z = np.load(os.path.join(GOLDEN, NAME + ".npz"))
    with open(os.path.join(GOLDEN, NAME + ".json")) as f:
        meta = json.load(f)
    return {k: z[k] for k in z.files}, meta


G, META = _load()
ORIENT = [tuple(c) for c in META["orientation"]]
SMOOTH = [tuple(c) for c in META["smooth"]]


def _orient_kw(case):
    name, bs, ss, inv, sos, use_mask = case
    return name, dict(block_size=bs, smooth_sigma=ss, invert_if_needed=inv, smooth_orientation_sigma=sos,
                      mask=G["mask"] if use_mask else None)


@pytest.mark.parametrize("case", ORIENT, ids=[c[0] for c in ORIENT])
def test_oracle_orientation_kwargs_match_reference(case):
    from oracle import ref_pipeline as rp
    name, kw = _orient_kw(case)
    blk, oimg, rel = rp.compute_orientation_map(G["img"], **kw)
    np.testing.assert_allclose(blk, G[f"orient_{name}_blocks"], rtol=0, atol=1e-5)
    assert angle_diff(oimg, G[f"orient_{name}_img"]).max() <= 1e-5
    np.testing.assert_allclose(rel, G[f"orient_{name}_rel"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", SMOOTH, ids=[c[0] for c in SMOOTH])
def test_oracle_smooth_kwargs_match_reference(case):
    from oracle import ref_pipeline as rp
    name, sg, it, boost = case
    assert_same(rp.smooth_fingerprint_skeleton(G["binary"], sigma=sg, diffusion_iter=it, contrast_boost=boost),
                G[f"smooth_{name}"], f"smooth {name}")


@pytest.mark.gpu
@pytest.mark.parametrize("case", ORIENT, ids=[c[0] for c in ORIENT])
def test_gpu_orientation_kwargs_within_1e4(case):
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.orientation import compute_orientation_map
    name, kw = _orient_kw(case)
    blk, oimg, rel = compute_orientation_map(G["img"], **kw)
    w_blk, w_img, w_rel = G[f"orient_{name}_blocks"], G[f"orient_{name}_img"], G[f"orient_{name}_rel"]
    assert blk.shape == w_blk.shape and blk.dtype == np.float32 and oimg.shape == w_img.shape and rel.shape == w_rel.shape
    tol = 1e-4
    e_blk = angle_diff(blk, w_blk).max()
    e_img = angle_diff(oimg, w_img).max()
    e_rel = np.abs(rel - w_rel).max() / max(1e-12, np.abs(w_rel).max())
    assert e_blk <= tol * np.pi and e_img <= tol * np.pi and e_rel <= tol, (name, e_blk, e_img, e_rel)


@pytest.mark.gpu
@pytest.mark.parametrize("case", SMOOTH, ids=[c[0] for c in SMOOTH])
def test_gpu_smooth_kwargs_bit_exact(case):
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.fingerprint_preprocess import smooth_fingerprint_skeleton
    name, sg, it, boost = case
    got = smooth_fingerprint_skeleton(G["binary"], sigma=sg, diffusion_iter=it, contrast_boost=boost)
    assert_same(got, G[f"smooth_{name}"], f"smooth {name}", f"k6kw_{name}")


@pytest.mark.gpu
def test_gpu_unfused_smooth_equals_fused_on_defaults():
    """fpb_smooth_ex with the default values (unfused kernel sequence) == fpb_smooth (k_smooth_fused), bit for bit,
    also on a crop smaller than the handle and on a batch."""
    from multimodal_biometric_fingerprints_palms_b200.pipeline import pipeline_for
    b = np.stack([G["binary"], G["binary"][::-1].copy(), G["binary"][:, ::-1].copy()])
    p = pipeline_for(*b.shape[1:], max_batch=3)
    assert_same(p.smooth(b, force_unfused=True), p.smooth(b), "unfused vs fused smooth")
    c = np.ascontiguousarray(b[:, :101, :77])
    assert_same(p.smooth(c, force_unfused=True), p.smooth(c), "unfused vs fused smooth on a crop")


@pytest.mark.gpu
def test_gpu_orientation_ex_defaults_equal_the_hot_path_entry():
    """fpb_orientation_ex(16, 3.0, 1, 3.0) must give fpb_orientation's arrays exactly (same kernels)."""
    from multimodal_biometric_fingerprints_palms_b200 import _native as N
    from multimodal_biometric_fingerprints_palms_b200.pipeline import pipeline_for, _ptr
    img, mask = G["img"], G["mask"]
    p = pipeline_for(*img.shape)
    want = p.orientation(img, mask)
    a, (h, w) = p._batch_roi(img); m = p._batch_roi(mask)[0]
    blocks = np.zeros((1, p.H // 16, p.W // 16), np.float32)
    oimg = np.empty((1, p.H, p.W), np.float32); rel = np.empty_like(oimg)
    rc = N.load().fpb_orientation_ex(p._h, _ptr(a), _ptr(m), 1, 16, 3.0, 1, 3.0, _ptr(blocks), _ptr(oimg), _ptr(rel))
    assert rc == 0
    assert_same(blocks[:, :h // 16, :w // 16], want[0], "blocks")
    assert_same(oimg[:, :h, :w], want[1], "orient_img")
    assert_same(rel[:, :h, :w], want[2], "rel_img")


@pytest.mark.gpu
def test_gpu_orientation_kwargs_errors_follow_the_reference():
    import cv2
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.orientation import compute_orientation_map
    with pytest.raises(ZeroDivisionError):
        compute_orientation_map(G["img"], block_size=0)
    with pytest.raises(cv2.error):
        compute_orientation_map(G["img"], block_size=200)
    with pytest.raises(ValueError):
        compute_orientation_map(G["img"], block_size=-16)


# ---- postprocess_minutiae's `gray` argument (post_processing.py:71, 93) ------------------------------------------------
def _post_inputs(name):
    from conftest import load_golden
    g, lists = load_golden(META["post_case"])
    gray = {"none": None, "segmented": g["segmented"], "skeleton_file": g["skeleton_file"]}[name]
    return g["skeleton"], gray, lists["raw_minutiae"], META["post_gray"][name]


def _check_refined(got, want):
    assert [(m["x"], m["y"], m["type"]) for m in got] == [(m["x"], m["y"], m["type"]) for m in want]
    for a, b in zip(got, want):
        assert angle_diff(a["orientation"], b["orientation"]) <= 1e-4 * np.pi
        for k in ("quality", "coherence", "angular_stability"):
            assert abs(a[k] - b[k]) <= 1e-4 * max(1.0, abs(b[k])), (k, a[k], b[k])


@pytest.mark.parametrize("name", ["none", "segmented", "skeleton_file"])
def test_oracle_postprocess_gray_matches_reference(name):
    from oracle import ref_pipeline as rp
    skel, gray, raw, want = _post_inputs(name)
    _check_refined(rp.postprocess_minutiae([dict(m) for m in raw], skel, gray, None), want)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["none", "segmented", "skeleton_file"])
def test_gpu_postprocess_gray_matches_reference(name):
    from multimodal_biometric_fingerprints_palms_b200.features.post_processing import postprocess_minutiae
    skel, gray, raw, want = _post_inputs(name)
    mine = [dict(m) for m in raw]
    got = postprocess_minutiae(mine, skel, gray, None)
    _check_refined(got, want)
    assert all(any(g is m for m in mine) for g in got)          # the caller's dicts, updated in place


# ---- run_preprocessing(debug=True, small_subset=True) (run_preprocessing.py:49-66, 90-92, 103-108, 143-144) -------------
@pytest.mark.gpu
def test_gpu_run_preprocessing_debug_tree_and_small_subset(tmp_path):
    """debug=True writes debug/<relative dir>/<result key>/<base>.jpg for the seven keys of preprocess_fingerprint's dict,
    with the pixels the reference would write (oracle planes through the same cv2.imwrite); small_subset keeps the
    first ten files."""
    import cv2
    from multimodal_biometric_fingerprints_palms_b200 import synth
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.run_preprocessing import run_preprocessing
    from oracle import ref_pipeline as rp
    src = tmp_path / "sorted" / "cluster_3"
    src.mkdir(parents=True)
    imgs = {f"p{k:02d}": synth.ridge_image(160, 128, seed=40 + k, period=8) for k in range(12)}
    for name, im in imgs.items():
        cv2.imwrite(str(src / f"{name}.png"), im)
    out = tmp_path / "processed"
    assert run_preprocessing(str(tmp_path / "sorted"), str(out), debug=True, small_subset=True, max_workers=2) == 10
    keys = ("normalized", "denoised", "segmented", "mask", "binary", "skeleton", "orientation_vis")
    done = sorted(p.name[:-len("_skeleton.jpg")] for p in (out / "enhanced" / "cluster_3").glob("*_skeleton.jpg"))
    assert len(done) == 10
    for name in done[:3]:
        want = rp.preprocess_fingerprint(imgs[name])
        for key in keys:
            path = out / "debug" / "cluster_3" / key / f"{name}.jpg"
            assert path.exists(), path
            if key == "orientation_vis":
                assert cv2.imread(str(path)).shape == want["segmented"].shape + (3,)
                continue
            ref_path = tmp_path / f"ref_{key}.jpg"
            cv2.imwrite(str(ref_path), want[key])
            assert_same(cv2.imread(str(path), cv2.IMREAD_GRAYSCALE), cv2.imread(str(ref_path), cv2.IMREAD_GRAYSCALE),
                        f"debug image {key} of {name}")


# ---- segment_fingerprint on colour input (fingerprint_preprocess.py:94) -----------------------------------------------------
def test_bgr2gray_weights_equal_cv2_on_every_colour():
    """The fixed-point formula k_bgr2gray uses, against cv2.cvtColor on all 2^24 (B, G, R) triples."""
    import cv2
    v = np.arange(256, dtype=np.uint8)
    B, G_, R = np.meshgrid(v, v, v, indexing="ij")
    full = np.stack([B.ravel(), G_.ravel(), R.ravel()], 1).reshape(4096, 4096, 3)
    f = full.astype(np.int32)
    mine = ((f[..., 0] * 3735 + f[..., 1] * 19235 + f[..., 2] * 9798 + 16384) >> 15).astype(np.uint8)
    assert np.array_equal(mine, cv2.cvtColor(full, cv2.COLOR_BGR2GRAY))


def test_oracle_segment_bgr_matches_reference():
    from oracle import ref_pipeline as rp
    seg, mask = rp.segment_fingerprint(G["bgr"])
    assert_same(seg, G["bgr_segmented"], "segmented (BGR input)")
    assert_same(mask, G["bgr_mask"], "mask (BGR input)")


@pytest.mark.gpu
def test_gpu_segment_bgr_bit_exact():
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.fingerprint_preprocess import segment_fingerprint
    seg, mask = segment_fingerprint(G["bgr"])
    assert_same(seg, G["bgr_segmented"], "segmented (BGR input)", "k3_bgr_seg")
    assert_same(mask, G["bgr_mask"], "mask (BGR input)", "k3_bgr_mask")
    bgra = np.concatenate([G["bgr"], np.full(G["bgr"].shape[:2] + (1,), 77, np.uint8)], axis=-1)
    seg4, mask4 = segment_fingerprint(bgra)                       # alpha is ignored, as cv2.COLOR_BGR2GRAY does
    assert_same(seg4, seg, "segmented (BGRA input)")
    assert_same(mask4, mask, "mask (BGRA input)")


# ---- compute_orientation_map on non-uint8 images (orientation.py:21-28) -----------------------------------------------------
def _float_cases():
    from oracle.make_golden_args import float_cases
    return {name: (fimg, kw) for name, fimg, kw in float_cases(G["img"], G["mask"], G["binary"])}


FLOAT_NAMES = ["unit_f32", "int16_range", "f64_0_255_bs8", "half_at_max", "over_half_at_max"]


@pytest.mark.parametrize("name", FLOAT_NAMES)
def test_oracle_orientation_of_non_uint8_images_matches_reference(name):
    from oracle import ref_pipeline as rp
    fimg, kw = _float_cases()[name]
    blk, oimg, rel = rp.compute_orientation_map(fimg, **kw)
    np.testing.assert_allclose(blk, G[f"float_{name}_blocks"], rtol=0, atol=1e-5)
    assert angle_diff(oimg, G[f"float_{name}_img"]).max() <= 1e-5
    np.testing.assert_allclose(rel, G[f"float_{name}_rel"], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", FLOAT_NAMES)
def test_gpu_orientation_of_non_uint8_images_within_1e4(name):
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.orientation import compute_orientation_map
    fimg, kw = _float_cases()[name]
    blk, oimg, rel = compute_orientation_map(fimg, **kw)
    w_blk, w_img, w_rel = G[f"float_{name}_blocks"], G[f"float_{name}_img"], G[f"float_{name}_rel"]
    assert blk.shape == w_blk.shape and oimg.shape == w_img.shape
    tol = 1e-4
    e_blk = angle_diff(blk, w_blk).max()
    e_img = angle_diff(oimg, w_img).max()
    e_rel = np.abs(rel - w_rel).max() / max(1e-12, np.abs(w_rel).max())
    assert e_blk <= tol * np.pi and e_img <= tol * np.pi and e_rel <= tol, (name, e_blk, e_img, e_rel)


def _float_prep_model(a, allow_invert=True):
    """NumPy statement of k_or_float_prep (k_orient.cu): min / max, optional rescale in float32, and the "max > median"
    test from the count of maxima instead of a selection."""
    f = a.astype(np.float32); mn = f.min(); mx = f.max()
    norm = mx > 1.0 or mn < 0.0
    den = np.float32(np.float32(mx - mn) + np.float32(1e-12))
    v = ((f - mn) / den).astype(np.float32) if norm else f
    top = np.float32((mx - mn) / den) if norm else mx
    cnt = int((v == top).sum()); n = v.size
    rest = v[v != top]
    sec = rest.max() if rest.size else np.float32(-np.inf)
    if not allow_invert or n == 0:
        inv = False
    elif n & 1:
        inv = cnt < (n + 1) // 2
    elif cnt >= n // 2 + 1:
        inv = False
    elif cnt == n // 2:
        inv = bool(top > np.float32(np.float32(sec + top) / np.float32(2)))
    else:
        inv = True
    return (np.float32(1.0) - v) if inv else v


def test_float_front_end_model_equals_the_reference_preprocessing():
    """The decisions the device kernel takes (rescale? invert?) against orientation.py:21-28 as the oracle restates it,
    on the five golden float cases and 60 random small images (ties at the maximum, integer ranges, binary images)."""
    from oracle import ref_pipeline as rp
    rng = np.random.default_rng(1)
    cases = [f for f, _ in _float_cases().values()]
    for k in range(60):
        h, w = rng.integers(3, 12, 2)
        kind = k % 4
        if kind == 0:
            a = rng.random((h, w)).astype(np.float32)
        elif kind == 1:
            a = rng.integers(-5, 6, (h, w)).astype(np.float64)
        elif kind == 2:
            a = (rng.random((h, w)) > rng.random()).astype(np.float32)
        else:
            a = rng.random((h, w)).astype(np.float32)
            a.ravel()[rng.permutation(a.size)[:a.size // 2]] = a.max()
        cases.append(a)
    for a in cases:
        d = {}
        rp.compute_orientation_map(a, block_size=2, detail=d)
        assert np.array_equal(d["f"], _float_prep_model(a))
        d = {}
        rp.compute_orientation_map(a, block_size=2, invert_if_needed=False, detail=d)
        assert np.array_equal(d["f"], _float_prep_model(a, allow_invert=False))


# ---- random keyword combinations against the oracle (which the goldens above pin to the reference) ---------------------------
def _random_print(rng, h, w):
    from multimodal_biometric_fingerprints_palms_b200 import synth
    img = synth.ridge_image(h, w, seed=int(rng.integers(1 << 30)), period=float(rng.uniform(6, 11)))
    yy, xx = np.mgrid[0:h, 0:w]
    inside = ((xx - w / 2.0) / (0.45 * w)) ** 2 + ((yy - h / 2.0) / (0.48 * h)) ** 2 <= 1.0
    return img, inside.astype(np.uint8) * 255, ((img < 128) & inside).astype(np.uint8) * 255


@pytest.mark.gpu
def test_gpu_random_keyword_combinations_against_the_oracle():
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.fingerprint_preprocess import smooth_fingerprint_skeleton
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.orientation import compute_orientation_map
    from oracle import ref_pipeline as rp
    rng = np.random.default_rng(20261019)
    worst = (0.0, 0.0, 0.0)
    for k in range(16):
        h, w = int(rng.integers(70, 230)), int(rng.integers(70, 230))
        img, mask, binary = _random_print(rng, h, w)
        kw = dict(block_size=int(rng.choice([5, 8, 10, 16, 21, 32])), smooth_sigma=float(rng.choice([0.0, 0.8, 1.7, 3.0, 4.5, 6.0])),
                  invert_if_needed=bool(rng.integers(2)), smooth_orientation_sigma=float(rng.choice([0.0, 0.5, 1.3, 3.0, 7.0])),
                  mask=mask if rng.integers(2) else None)
        src = img if k % 3 else (img.astype(np.float32) * np.float32(rng.uniform(0.5, 3.0)) - np.float32(rng.uniform(0, 50)))
        want = rp.compute_orientation_map(src, **kw)
        got = compute_orientation_map(src, **kw)
        e_blk = angle_diff(got[0], want[0]).max()
        e_img = angle_diff(got[1], want[1]).max()
        e_rel = np.abs(got[2] - want[2]).max() / max(1e-12, np.abs(want[2]).max())
        assert e_blk <= 1e-4 * np.pi and e_img <= 1e-4 * np.pi and e_rel <= 1e-4, (k, h, w, kw["block_size"], kw["smooth_sigma"],
                                                                                 kw["smooth_orientation_sigma"], e_blk, e_img, e_rel)
        worst = (max(worst[0], e_blk), max(worst[1], e_img), max(worst[2], e_rel))
        sg, it, boost = float(rng.uniform(0.2, 2.5)), int(rng.integers(0, 6)), float(rng.uniform(0.8, 2.0))
        assert_same(smooth_fingerprint_skeleton(binary, sigma=sg, diffusion_iter=it, contrast_boost=boost),
                    rp.smooth_fingerprint_skeleton(binary, sigma=sg, diffusion_iter=it, contrast_boost=boost),
                    f"smooth({sg:.3f}, {it}, {boost:.3f}) on {h}x{w}", f"k6kw_rand{k}")
    print("worst K5 errors over the random combinations (blocks rad, image rad, rel):", worst)


@pytest.mark.gpu
def test_gpu_preprocess_fingerprint_debug_and_mask_files(tmp_path):
    """debug_dir: the six JPEGs of fingerprint_preprocess.py:205-212 (no mask.jpg); save_mask_dir + img_name: the cropped
    hull mask (:132-134)."""
    import cv2
    from multimodal_biometric_fingerprints_palms_b200 import synth
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.fingerprint_preprocess import preprocess_fingerprint
    img = synth.ridge_image(160, 128, seed=77, period=8)
    dbg, msk = tmp_path / "dbg", tmp_path / "masks"
    res = preprocess_fingerprint(img, debug_dir=str(dbg), save_mask_dir=str(msk), img_name="p.png")
    assert sorted(p.name for p in dbg.iterdir()) == sorted(f"{k}.jpg" for k in ("normalized", "denoised", "segmented", "binary",
                                                                              "skeleton", "orientation_vis"))
    assert_same(cv2.imread(str(msk / "p.png"), cv2.IMREAD_GRAYSCALE), res["mask"], "saved mask")
    assert set(res) == {"normalized", "denoised", "segmented", "mask", "binary", "skeleton", "orientation_vis"}


# ---- postprocess_minutiae's params: even / large quality_window, other radii (post_processing.py:77-83) --------------------
@pytest.mark.gpu
@pytest.mark.parametrize("params", [
    {"quality_window": 26}, {"quality_window": 33, "quality_threshold": 0.1}, {"quality_window": 8, "margin": 20},
    {"quality_window": 2, "quality_threshold": 0.0, "coherence_threshold": 0.1}, {"patch_radius": 7, "max_minutiae": 128, "min_distance": 3.0},
    {"max_minutiae": 0}], ids=["w26", "w33", "w8", "w2", "r7_max128", "max0"])
def test_gpu_postprocess_params_against_the_oracle(params):
    from conftest import load_golden
    from multimodal_biometric_fingerprints_palms_b200.features.post_processing import postprocess_minutiae
    from oracle import ref_pipeline as rp
    g, lists = load_golden(META["post_case"])
    skel, raw = g["skeleton_file"], lists["raw_minutiae_file"]
    want = rp.postprocess_minutiae([dict(m) for m in raw], skel, skel, params)
    got = postprocess_minutiae([dict(m) for m in raw], skel, skel, params)
    _check_refined(got, want)


@pytest.mark.gpu
def test_gpu_postprocess_params_out_of_range_are_loud():
    from multimodal_biometric_fingerprints_palms_b200 import FpbError
    from multimodal_biometric_fingerprints_palms_b200.pipeline import pipeline_for
    p = pipeline_for(64, 64)
    for bad in ({"quality_window": 34}, {"quality_window": 0}, {"max_minutiae": 129}, {"max_minutiae": -1}):
        with pytest.raises(FpbError):
            p.set_post_params(bad)
    p.set_post_params(None)


def test_density_window_model_equals_cv2_blur_for_even_and_odd_windows():
    """k_density's formulation - integer window counts over [x - win // 2, x - win // 2 + win - 1] with BORDER_REFLECT_101,
    float32(count / win^2) - against cv2.blur for the window sizes fpb_set_post_params accepts (1 .. 33, odd or even)."""
    import cv2
    rng = np.random.default_rng(3)
    a = (rng.random((45, 38)) > 0.7).astype(np.float32)
    h, w = a.shape

    def refl(i, n):
        i = np.abs(i)
        return np.where(i >= n, 2 * (n - 1) - i, i)

    for win in (1, 2, 8, 24, 25, 26, 33):
        r = win // 2
        ys, xs = refl(np.arange(-r, h - r + win - 1), h), refl(np.arange(-r, w - r + win - 1), w)
        pad = a[np.ix_(ys, xs)].astype(np.int64)
        c = np.cumsum(np.cumsum(np.pad(pad, ((1, 0), (1, 0))), 0), 1)
        cnt = c[win:win + h, win:win + w] - c[:h, win:win + w] - c[win:win + h, :w] + c[:h, :w]
        mine = (cnt.astype(np.float64) / (win * win)).astype(np.float32)
        assert np.abs(mine - cv2.blur(a, (win, win))).max() <= 3e-8, win
