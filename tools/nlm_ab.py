"""A/B of the two NLM kernels (tensor-core k_nlm_mma vs integer-ALU k_nlm; FPB_NLM_MMA=1 selects the former):
bit-exactness against cv2.fastNlMeansDenoising on assorted shapes, then the kernel time on the 1480-image batch.
    python tools/nlm_ab.py            # runs itself twice (one process per kernel) and prints one JSON line each"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import cv2
    import numpy as np
    from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline, synth
    rng = np.random.default_rng(0)
    rep = {"kernel": "mma" if os.environ.get("FPB_NLM_MMA") == "1" else "scalar", "cases": []}
    for (h, w, n, kind) in [(320, 240, 6, "ridge"), (64, 48, 3, "noise"), (131, 97, 2, "ridge"), (333, 251, 2, "ridge"), (16, 8, 2, "noise"),
                            (40, 200, 2, "flat"), (512, 512, 2, "degraded"), (17, 9, 1, "noise")]:
        if kind == "ridge":
            imgs = np.stack([synth.ridge_image(h, w, seed=10 + i, period=None) for i in range(n)])
        elif kind == "degraded":
            imgs = np.stack([synth.degraded_image(h, w, seed=20 + i) for i in range(n)])
        elif kind == "flat":
            imgs = np.full((n, h, w), 140, np.uint8); imgs[1, 5:20, 30:90] = 20
        else:
            imgs = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
        p = FingerprintPipeline(h, w, max_batch=n)
        _, nlm = p.denoise(imgs, with_nlm=True)
        bad = sum(int((nlm[i] != cv2.fastNlMeansDenoising(imgs[i], None, 10, 7, 21)).sum()) for i in range(n))
        rep["cases"].append({"shape": [h, w], "n": n, "kind": kind, "mismatching_pixels": bad})
        p.close()
    n = int(os.environ.get("NLM_AB_BATCH", "1480"))
    imgs = synth.ridge_batch(min(n, 64), 320, 240, first_seed=0)
    imgs = np.stack([imgs[i % len(imgs)] for i in range(n)])
    p = FingerprintPipeline(320, 240, max_batch=n)
    p.set_profiling(True)
    times = []
    for _ in range(4):
        p.run(imgs)
        times.append(p.stage_times_ms()["nlm_kernel"])
    rep["nlm_kernel_ms_1480"] = min(times[1:])
    rep["all_runs_ms"] = times
    print(json.dumps(rep))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for env in ({"FPB_NLM_MMA": "1"}, {"FPB_NLM_MMA": "0"}):
            e = dict(os.environ); e.update(env)
            r = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True, timeout=600)
            print(r.stdout.strip() or ("FAILED: " + r.stderr[-2000:]))
