"""A/B of the NLM kernels (default k_nlm_sym; FPB_NLM_V=3 k_nlm3; FPB_NLM_V=1 k_nlm; FPB_NLM_MMA=1 the tensor-core k_nlm_mma):
bit-exactness against cv2.fastNlMeansDenoising on assorted shapes, then the kernel time on the 1480-image batch.
    python tools/nlm_ab.py [sym] [v3] [v1] [mma]   # one process per kernel, one JSON line each (default: sym v3)"""
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
    rep = {"kernel": os.environ.get("NLM_AB_NAME", "?"), "cases": []}
    short = os.environ.get("NLM_AB_SHORT") == "1"      # only the first three shapes, then the timing
    for (h, w, n, kind) in [(320, 240, 6, "ridge"), (64, 48, 3, "noise"), (131, 97, 2, "ridge"), (333, 251, 2, "ridge"), (16, 8, 2, "noise"),
                            (40, 200, 2, "flat"), (512, 512, 2, "degraded"), (17, 9, 1, "noise"), (240, 320, 3, "ridge"),
                            (100, 256, 2, "noise"), (37, 241, 2, "noise"), (1, 1, 1, "noise"), (5, 300, 1, "noise"),
                            (300, 5, 1, "noise"), (96, 480, 2, "noise"), (1024, 1024, 1, "degraded"), (320, 240, 150, "ridge")][:3 if short else None]:
        if kind == "ridge":
            imgs = np.stack([synth.ridge_image(h, w, seed=10 + i, period=None) for i in range(n)])
        elif kind == "degraded":
            imgs = np.stack([synth.degraded_image(h, w, seed=20 + i) for i in range(n)])
        elif kind == "flat":
            imgs = np.full((n, h, w), 140, np.uint8); imgs[1, 5:20, 30:90] = 20
        else:
            imgs = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
        try:
            p = FingerprintPipeline(h, w, max_batch=n)
            _, nlm = p.denoise(imgs, with_nlm=True)
            chk = range(n) if n <= 8 else range(0, n, 13)
            bad = sum(int((nlm[i] != cv2.fastNlMeansDenoising(imgs[i], None, 10, 7, 21)).sum()) for i in chk)
            rep["cases"].append({"shape": [h, w], "n": n, "kind": kind, "mismatching_pixels": bad})
            p.close()
        except Exception as e:      # report and go on: one refused shape must not hide the others
            rep["cases"].append({"shape": [h, w], "n": n, "kind": kind, "error": repr(e)[:200]})
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
        envs = {"sym": {}, "v3": {"FPB_NLM_V": "3"}, "v1": {"FPB_NLM_V": "1"}, "mma": {"FPB_NLM_MMA": "1"}}
        for name in (sys.argv[1:] or ["sym", "v3"]):
            e = dict(os.environ); e.update(envs[name]); e["NLM_AB_NAME"] = name
            r = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True, timeout=600)
            print(r.stdout.strip() or ("FAILED: " + r.stderr[-2000:]))
