"""EXTENSION rows G1/G2 (ridge frequency + Gabor bank; NOT in the reference, parity unpinned): the CUDA kernels against
the NumPy statement of the same arithmetic (oracle/gabor_ext.py), fed the block orientations the GPU itself used.

Tolerances: block frequencies 1e-6 absolute (they are ratios of small integers in float32); Gabor response 1e-4 of
the response range (float32 FMA accumulation vs float64 einsum); enhanced u8 within one level, <= 0.1 % of pixels off."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(h, w, seed, period):
    from multimodal_biometric_fingerprints_palms_b200.synth import ridge_image
    img = ridge_image(h, w, seed=seed, period=period)
    yy, xx = np.mgrid[:h, :w]
    mask = ((((xx - w / 2) / (0.45 * w)) ** 2 + ((yy - h / 2) / (0.47 * h)) ** 2) <= 1).astype(np.uint8) * 255
    return img, mask


@pytest.mark.parametrize("h,w,seed,period", [(320, 240, 0, 9.0), (240, 320, 3, 7.0), (203, 187, 5, 11.0)])
def test_frequency_and_gabor_match_numpy_statement(h, w, seed, period):
    from oracle import gabor_ext as ge
    from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline
    img, mask = _case(h, w, seed, period)
    p = FingerprintPipeline(h, w, max_batch=2)
    both = np.stack([img, img[::-1].copy()]); masks = np.stack([mask, mask[::-1].copy()])
    blocks, _, _ = p.orientation(both, masks)
    fb, resp, enh = p.enhance_gabor(both, masks)
    for b in range(2):
        raw = ge.ridge_frequency_raw(both[b], masks[b], blocks[b])
        want_f = ge.fill_frequency(raw)
        np.testing.assert_allclose(fb[b], want_f, rtol=0, atol=1e-6)
        valid = raw[raw > 0]
        assert len(valid) > 0.3 * raw.size                                     # the estimator does find ridges
        assert abs(np.median(1.0 / valid) - period) < 1.5                      # and the right period
        want_r, want_e = ge.gabor_enhance(both[b], masks[b], blocks[b], fb[b])
        scale = np.abs(want_r).max()
        assert scale > 10
        assert np.abs(resp[b] - want_r).max() <= 1e-4 * scale
        d = np.abs(enh[b].astype(int) - want_e.astype(int))
        assert d.max() <= 1 and (d > 0).mean() <= 1e-3
    p.close()


def test_enhanced_plane_in_fused_run_and_result_key(monkeypatch):
    from oracle import gabor_ext as ge
    from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline
    from multimodal_biometric_fingerprints_palms_b200.synth import ridge_image
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.fingerprint_preprocess import preprocess_fingerprint
    img = ridge_image(320, 240, seed=2)
    p = FingerprintPipeline(320, 240, max_batch=1)
    p.run(img)
    base = {k: p.fetch(k) for k in ("skeleton", "binary", "orient_img")}
    base_min = p.minutiae(0)
    p.enable_enhanced({"n_orient": 16})
    p.run(img)
    for k, v in base.items():
        assert np.array_equal(p.fetch(k), v), f"{k} changed when the extension was switched on"
    assert p.minutiae(0) == base_min
    x0, y0, w, h = p.roi(0)
    seg, mask = p.fetch("segmented")[0, :h, :w], p.fetch("mask")[0, :h, :w]
    enh, resp = p.fetch("enhanced")[0, :h, :w], p.fetch("gabor_response")[0, :h, :w]
    fb = p.freq_blocks()[0, :h // 16, :w // 16]
    q = FingerprintPipeline(h, w, max_batch=1)
    blocks, _, _ = q.orientation(np.ascontiguousarray(seg), np.ascontiguousarray(mask))
    want_f = ge.fill_frequency(ge.ridge_frequency_raw(seg, mask, blocks[0]))
    np.testing.assert_allclose(fb, want_f, rtol=0, atol=1e-6)
    want_r, want_e = ge.gabor_enhance(seg, mask, blocks[0], fb)
    assert np.abs(resp - want_r).max() <= 1e-4 * np.abs(want_r).max()
    assert (np.abs(enh.astype(int) - want_e.astype(int)) > 1).sum() == 0
    # ridges stay dark, valleys bright: the enhanced image correlates with the input inside the mask
    on = mask > 0
    assert np.corrcoef(enh[on].astype(float), seg[on].astype(float))[0, 1] > 0.5
    p.disable_enhanced()
    with pytest.raises(Exception):
        FingerprintPipeline(64, 64).fetch("enhanced")
    # result-dict key: absent by default (the reference never produces it), present when opted in
    assert "enhanced" not in preprocess_fingerprint(img)
    monkeypatch.setenv("FPB200_ENHANCED", "1")
    out = preprocess_fingerprint(img)
    assert out["enhanced"].shape == out["segmented"].shape and out["enhanced"].dtype == np.uint8


def test_highres_period18_uses_the_large_filters():
    """configs[4] shape: 1024x1024, period 18 -> sigma 8.1, 43x43 taps"""
    from oracle import gabor_ext as ge
    from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline
    img, mask = _case(1024, 1024, 11, 18.0)
    p = FingerprintPipeline(1024, 1024, max_batch=1)
    blocks, _, _ = p.orientation(img, mask)
    fb, resp, enh = p.enhance_gabor(img, mask)
    raw = ge.ridge_frequency_raw(img, mask, blocks[0])
    np.testing.assert_allclose(fb[0], ge.fill_frequency(raw), rtol=0, atol=1e-6)
    assert abs(np.median(1.0 / raw[raw > 0]) - 18.0) < 2.5
    sub = (slice(384, 640), slice(384, 640))                      # the NumPy statement on a 256x256 window of blocks
    want_r, _ = ge.gabor_enhance(img, mask, blocks[0], fb[0])
    assert np.abs(resp[0][sub] - want_r[sub]).max() <= 1e-4 * np.abs(want_r).max()
    p.close()
