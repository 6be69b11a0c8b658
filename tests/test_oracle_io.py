"""CPU: the host half of the JPEG input path (marker parser + Huffman decoder of libfpb200.so) and the NumPy islow IDCT
against cv2.imdecode(IMREAD_GRAYSCALE) - i.e. against what the reference reads at run_preprocessing.py:41 - and the
native JSON writer against json.dumps(indent=2) (extract_features.py:104-105).  No GPU calls."""
import ctypes as C
import json
import math

import cv2
import numpy as np
import pytest


def jpeg_cases():
    from multimodal_biometric_fingerprints_palms_b200.synth import ridge_image
    rng = np.random.default_rng(0)
    out = []
    for (h, w, q, extra) in [(320, 240, 95, []), (240, 320, 75, []), (333, 251, 90, []), (64, 64, 100, []), (17, 9, 50, []),
                             (320, 240, 95, [cv2.IMWRITE_JPEG_RST_INTERVAL, 7]), (200, 184, 30, [cv2.IMWRITE_JPEG_OPTIMIZE, 1])]:
        img = ridge_image(h, w, seed=h + q) if h >= 64 else rng.integers(0, 256, (h, w), dtype=np.uint8)
        ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q] + extra)
        assert ok
        out.append((f"gray_{h}x{w}_q{q}_{len(extra)}", buf.tobytes(), h, w))
    noise = rng.integers(0, 256, (96, 128), dtype=np.uint8)
    ok, buf = cv2.imencode(".jpg", noise, [cv2.IMWRITE_JPEG_QUALITY, 100]); out.append(("noise_q100", buf.tobytes(), 96, 128))
    col = np.dstack([ridge_image(160, 120, seed=k) for k in range(3)])
    for ss, name in ((cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "420"), (cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "444"),
                     (cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, "422")):
        ok, buf = cv2.imencode(".jpg", col, [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, ss])
        out.append((f"colour_{name}", buf.tobytes(), 160, 120))
    return out


def lib():
    from multimodal_biometric_fingerprints_palms_b200 import _native
    return _native.load()


@pytest.mark.parametrize("case", jpeg_cases(), ids=lambda c: c[0])
def test_entropy_decoder_plus_islow_idct_equals_cv2(case):
    from oracle.jpeg_idct import idct_islow
    name, data, h, w = case
    L = lib()
    wi, hi, ci = C.c_int(), C.c_int(), C.c_int()
    assert L.fpb_jpeg_info(data, len(data), C.byref(wi), C.byref(hi), C.byref(ci)) == 0
    assert (wi.value, hi.value) == (w, h)
    bh, bw = (h + 7) // 8, (w + 7) // 8
    coefs = np.zeros((bh, bw, 64), np.int16); qt = np.zeros(64, np.uint16)
    assert L.fpb_jpeg_coefficients(data, len(data), w, h, coefs.ctypes.data, qt.ctypes.data) == 0
    want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_GRAYSCALE)
    got = idct_islow(coefs, qt, w, h)
    assert np.array_equal(got, want), f"{name}: {(got != want).sum()} pixels differ"


def test_unsupported_and_corrupt_streams_are_refused():
    L = lib()
    img = np.zeros((32, 32), np.uint8)
    ok, prog = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    coefs = np.zeros((4, 4, 64), np.int16); qt = np.zeros(64, np.uint16)
    assert L.fpb_jpeg_coefficients(prog.tobytes(), len(prog), 32, 32, coefs.ctypes.data, qt.ctypes.data) == -11
    ok, base = cv2.imencode(".jpg", img)
    b = base.tobytes()
    assert L.fpb_jpeg_coefficients(b, len(b), 40, 32, coefs.ctypes.data, qt.ctypes.data) == -12       # shape mismatch
    assert L.fpb_jpeg_coefficients(b[:len(b) // 2], len(b) // 2, 32, 32, coefs.ctypes.data, qt.ctypes.data) in (0, -10)
    assert L.fpb_jpeg_coefficients(b"\x89PNG\r\n\x1a\n" + b"0" * 64, 72, 32, 32, coefs.ctypes.data, qt.ctypes.data) == -10
    wi = C.c_int()
    assert L.fpb_jpeg_info(b"nope", 4, C.byref(wi), C.byref(wi), C.byref(wi)) == -10


def test_json_writer_equals_python_json_dump():
    from multimodal_biometric_fingerprints_palms_b200 import _native
    L = lib()
    rng = np.random.default_rng(3)
    specials = [0.0, -0.0, 1.0, -1.5, 1e-4, 9.999e-5, 1e-5, 1e15, 1e16, 1.2345678901234567e16, 123456789012345.6, 0.1, 1 / 3,
                5e-324, 1.7976931348623157e308, 2.5e-7, math.pi / 2, -math.pi / 2, 100.0, 1e22, 1.5e-10]
    vals = specials + list(rng.uniform(-2, 2, 200)) + list(10.0 ** rng.uniform(-12, 20, 100))
    recs, arr = [], (_native.Minutia * len(vals))()
    for i, v in enumerate(vals):
      with np.errstate(over="ignore"):
        o = float(v); q = float(vals[(i + 1) % len(vals)])
        o32 = float(np.float32(o))
        arr[i].x, arr[i].y, arr[i].type = i, 1000 - i, i & 1
        arr[i].orientation, arr[i].quality, arr[i].coherence, arr[i].angular_stability = o, q, o32, -q
        recs.append({"x": i, "y": 1000 - i, "type": "bifurcation" if i & 1 else "ending", "orientation": o, "quality": q,
                     "coherence": o32, "angular_stability": -q})
    for n in (0, 1, len(vals)):
        want = json.dumps(recs[:n], indent=2)
        need = L.fpb_minutiae_json(arr, n, None, 0)
        buf = C.create_string_buffer(need + 1)
        assert L.fpb_minutiae_json(arr, n, buf, need + 1) == need
        assert buf.value.decode() == want


def test_jpeg_decoder_random_streams_equal_cv2():
    """Random sizes (down to 1x1), qualities 1-100, restart intervals, optimised tables, every chroma sampling OpenCV can
    write: host entropy decoder + islow IDCT == cv2.imdecode(IMREAD_GRAYSCALE), bit for bit."""
    from oracle.jpeg_idct import idct_islow
    L = lib()
    rng = np.random.default_rng(7)
    samplings = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
                 cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411]
    for t in range(80):
        h, w = int(rng.integers(1, 200)), int(rng.integers(1, 200))
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        if t % 3 == 1:
            img = cv2.GaussianBlur(img, (0, 0), 2.0)
        colour = rng.random() < 0.3
        src = np.dstack([img, np.roll(img, 3, 0), 255 - img]) if colour else img
        params = [cv2.IMWRITE_JPEG_QUALITY, int(rng.integers(1, 101))]
        if rng.random() < 0.3:
            params += [cv2.IMWRITE_JPEG_RST_INTERVAL, int(rng.integers(1, 20))]
        if rng.random() < 0.3:
            params += [cv2.IMWRITE_JPEG_OPTIMIZE, 1]
        if colour:
            params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, int(rng.choice(samplings))]
        ok, buf = cv2.imencode(".jpg", src, params)
        data = buf.tobytes()
        coefs = np.zeros(((h + 7) // 8, (w + 7) // 8, 64), np.int16); qt = np.zeros(64, np.uint16)
        assert L.fpb_jpeg_coefficients(data, len(data), w, h, coefs.ctypes.data, qt.ctypes.data) == 0, (h, w, params)
        assert np.array_equal(idct_islow(coefs, qt, w, h), cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE)), (h, w, params)


def test_forward_islow_dct_and_quantiser_equal_cv2_encoder():
    """The skeleton hand-off of the reference (cv2.imwrite .jpg at quality 95, run_preprocessing.py:137-140): the NumPy
    statement of the ENCODER's lossy half (oracle/jpeg_fdct.py: edge replication, forward islow DCT, quantiser) must
    produce the very coefficients cv2 writes (read back with the library's host entropy decoder), and the full round
    trip must equal cv2.imencode + cv2.imdecode bit for bit."""
    from oracle.jpeg_fdct import fdct_quantise, jpeg_roundtrip, quality_table
    from oracle import ref_pipeline as rp
    from multimodal_biometric_fingerprints_palms_b200.synth import ridge_image
    L = lib()
    rng = np.random.default_rng(11)
    skel = rp.preprocess_fingerprint(ridge_image(320, 240, seed=3))["skeleton"]
    imgs = [(skel, 95), (rng.integers(0, 256, (97, 131), dtype=np.uint8), 95), ((rng.random((64, 200)) < 0.1).astype(np.uint8) * 255, 95),
            (np.full((33, 17), 255, np.uint8), 95), (np.zeros((8, 8), np.uint8), 95), (ridge_image(200, 184, seed=9), 95),
            (rng.integers(0, 256, (50, 75), dtype=np.uint8), 30), (rng.integers(0, 256, (50, 75), dtype=np.uint8), 100),
            (cv2.GaussianBlur(rng.integers(0, 256, (120, 90), dtype=np.uint8), (0, 0), 1.5), 75)]
    for img, q in imgs:
        ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q])
        data = buf.tobytes()
        h, w = img.shape
        coefs = np.zeros(((h + 7) // 8, (w + 7) // 8, 64), np.int16); qt = np.zeros(64, np.uint16)
        assert L.fpb_jpeg_coefficients(data, len(data), w, h, coefs.ctypes.data, qt.ctypes.data) == 0
        assert np.array_equal(qt, quality_table(q)), q
        mine = fdct_quantise(img, qt)
        assert np.array_equal(mine, coefs), f"{img.shape} q{q}: {(mine != coefs).sum()} coefficients differ"
        assert np.array_equal(jpeg_roundtrip(img, q), cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE))
    # OpenCV's default quality is the 95 the kernel hard-codes
    ok, buf = cv2.imencode(".jpg", skel)
    assert np.array_equal(jpeg_roundtrip(skel, 95), cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE))
