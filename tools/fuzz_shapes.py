"""Developer tool: random image shapes through the fused run against the oracle (exact planes and minutiae lists), plus one
odd-shaped batch large enough for the two-stream split.  Run on a GPU box under `timeout`."""
import sys, numpy as np, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline, synth
from oracle import ref_pipeline as rp
rng = np.random.default_rng(2024)
bad = 0
for t in range(24):
    h, w = int(rng.integers(96, 430)), int(rng.integers(96, 430))
    img = synth.ridge_image(h, w, seed=1000 + t, period=float(rng.uniform(6, 12)))
    p = FingerprintPipeline(h, w, max_batch=1)
    p.run(img)
    ref = rp.enhance_to_minutiae(img)
    x0, y0, cw, ch = p.roi(0)
    ok = ref["skeleton"].shape == (ch, cw)
    if ok:
        for k in ("mask", "binary", "binary_smooth", "skeleton"):
            ok &= bool((p.fetch(k)[0, :ch, :cw] == ref[k]).all())
        ok &= p.raw_minutiae(0) == ref["raw_minutiae"]
        ok &= [(m["x"], m["y"], m["type"]) for m in p.minutiae(0)] == [(m["x"], m["y"], m["type"]) for m in ref["minutiae"]]
    print(h, w, "crop", (ch, cw), "OK" if ok else "MISMATCH", flush=True)
    bad += not ok
    p.close()
print("mismatches:", bad)

# an odd-shaped batch through the two-stream split (n >= 64) against single-image runs
h, w, n = 203, 137, 70
imgs = np.stack([synth.ridge_image(h, w, seed=3000 + i % 7, period=8.0) for i in range(n)])
pb = FingerprintPipeline(h, w, max_batch=n)
pb.run(imgs)
sk = pb.fetch("skeleton")
p1 = FingerprintPipeline(h, w, max_batch=1)
bad2 = 0
for i in range(7):
    p1.run(imgs[i])
    for j in range(i, n, 7):
        same = pb.roi(j) == p1.roi(0) and pb.minutiae(j) == p1.minutiae(0) and bool((sk[j] == p1.fetch("skeleton")[0]).all())
        bad2 += not same
print("split-batch mismatches:", bad2)
