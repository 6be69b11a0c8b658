"""GPU: the JPEG input path (host Huffman + device islow IDCT) against cv2.imdecode - what the reference reads at
run_preprocessing.py:41 - bit for bit; the hot path on the decoded plane against the hot path on cv2's pixels; the
native JSON files against json.dump; and the fused directory driver end to end."""
import json
import os

import cv2
import numpy as np
import pytest

from test_oracle_io import jpeg_cases

pytestmark = pytest.mark.gpu


def test_decode_batch_equals_cv2_for_every_stream_kind():
    from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline
    for name, data, h, w in jpeg_cases():
        p = FingerprintPipeline(h, w, max_batch=3)
        st = p.decode_jpeg([data, data, data], threads=2)
        assert st.tolist() == [0, 0, 0], name
        got = p.fetch_input(3)
        want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_GRAYSCALE)
        for k in range(3):
            assert np.array_equal(got[k], want), f"{name}: {(got[k] != want).sum()} pixels differ"
        p.close()


def test_run_decoded_equals_run_on_cv2_pixels_and_json_files(tmp_path):
    from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline
    from multimodal_biometric_fingerprints_palms_b200.synth import ridge_image
    blobs, imgs = [], []
    for s in range(6):
        ok, buf = cv2.imencode(".jpg", ridge_image(320, 240, seed=40 + s), [cv2.IMWRITE_JPEG_QUALITY, 92])
        blobs.append(buf.tobytes()); imgs.append(cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE))
    ok, prog = cv2.imencode(".jpg", imgs[0], [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    p = FingerprintPipeline(320, 240, max_batch=8)
    st = p.decode_jpeg(blobs + [prog.tobytes()])
    assert st.tolist() == [0] * 6 + [-11]                            # progressive: refused, the caller uses cv2
    p.run_decoded(6)
    got = [(p.roi(i), p.minutiae(i)) for i in range(6)]
    skel = p.fetch("skeleton")
    paths = [str(tmp_path / f"s{i}_minutiae.json") for i in range(6)]
    assert p.write_json(paths, threads=3) == 6
    for i in range(6):
        assert open(paths[i]).read() == json.dumps(got[i][1], indent=2)
    q = FingerprintPipeline(320, 240, max_batch=8)
    q.run(np.stack(imgs))
    assert got == [(q.roi(i), q.minutiae(i)) for i in range(6)]
    assert np.array_equal(skel, q.fetch("skeleton"))
    assert sum(len(m) for _, m in got) > 20


def test_run_directory_mixed_inputs_against_oracle(tmp_path):
    """JPEG (GPU decode), progressive JPEG and PNG (cv2) in cluster dirs -> JSON equal to the oracle's on cv2's pixels."""
    from oracle import ref_pipeline as rp
    from multimodal_biometric_fingerprints_palms_b200.drivers import run_directory
    from multimodal_biometric_fingerprints_palms_b200.synth import ridge_image
    src = tmp_path / "in"
    for c in range(2):
        (src / f"cluster_{c}").mkdir(parents=True)
    names = []
    for s in range(5):
        img = ridge_image(320, 240, seed=70 + s)
        d = src / f"cluster_{s % 2}"
        if s == 3:
            path = d / f"{s:03d}_1.png"; cv2.imwrite(str(path), img)
        elif s == 4:
            path = d / f"{s:03d}_1.jpg"; cv2.imwrite(str(path), img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
        else:
            path = d / f"{s:03d}_1.jpg"; cv2.imwrite(str(path), img, [cv2.IMWRITE_JPEG_QUALITY, 95])
        names.append(path)
    out = tmp_path / "out"
    stats = run_directory(str(src), str(out), batch=4)
    assert stats["processed"] == 5 and stats["gpu_decoded"] == 3 and stats["unreadable"] == 0
    for path in names:
        rel = path.parent.name
        base = path.stem
        got = json.load(open(out / "minutiae" / rel / f"{base}_minutiae.json"))
        ref = rp.enhance_to_minutiae(cv2.imread(str(path), cv2.IMREAD_GRAYSCALE))
        assert [(m["x"], m["y"], m["type"]) for m in got] == [(m["x"], m["y"], m["type"]) for m in ref["minutiae"]]
        for a, b in zip(got, ref["minutiae"]):
            for k in ("orientation", "quality", "coherence", "angular_stability"):
                assert abs(a[k] - b[k]) <= 1e-4 * max(1.0, abs(b[k]))
        sk = cv2.imread(str(out / "enhanced" / rel / f"{base}_skeleton.jpg"), cv2.IMREAD_GRAYSCALE)
        assert np.array_equal(sk > 127, ref["skeleton"] > 127)
    again = run_directory(str(src), str(out), batch=4)                 # resume: nothing left to do
    assert again["processed"] == 0 and again["skipped"] == 5
    # bounded host memory: the same files in windows of two (three windows, handles reused across them) - identical JSON
    out2 = tmp_path / "out_windows"
    st2 = run_directory(str(src), str(out2), batch=2, window=2, write_skeletons=False)
    assert st2["processed"] == 5 and st2["gpu_decoded"] == 3
    for path in names:
        a = open(out / "minutiae" / path.parent.name / f"{path.stem}_minutiae.json").read()
        b = open(out2 / "minutiae" / path.parent.name / f"{path.stem}_minutiae.json").read()
        assert a == b
