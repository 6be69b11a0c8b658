"""Single-image latency (BASELINE configs[0]): one 320x240 print through fpb_run_host (H2D + K1..K9 + D2H), batch 1."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline, synth
img = synth.ridge_image(320, 240, seed=0)
p = FingerprintPipeline(320, 240, max_batch=1)
for _ in range(20):
    p.run(img)
ts = []
for _ in range(200):
    t0 = time.perf_counter(); p.run(img); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print(f"single image fpb_run_host: median {np.median(ts):.3f} ms, p10 {np.percentile(ts,10):.3f}, p90 {np.percentile(ts,90):.3f}; launches per run {p.launch_count // 220}")
for nb in (8, 64):
    q = FingerprintPipeline(320, 240, max_batch=nb)
    b = np.stack([img] * nb)
    for _ in range(5): q.run(b)
    t0 = time.perf_counter()
    for _ in range(20): q.run(b)
    print(f"batch {nb}: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per call")
