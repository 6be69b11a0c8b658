"""Developer tool: per-kernel warm timings inside the fused run (CUDA events after every launch).
   python tools/kernel_times.py [batch [H W]]"""
import sys, re, os, collections
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1480
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (320, 240)
base = synth.ridge_batch(min(n, 32), H, W, first_seed=500)
imgs = np.stack([base[i % len(base)] for i in range(n)])
p = FingerprintPipeline(H, W, max_batch=n)
if os.environ.get("FPB200_ENHANCED") == "1":      # EXTENSION rows G1/G2 timed next to the reference's stages
    p.enable_enhanced()
p.run(imgs); p.run(imgs)
p.kernel_times(True)
acc = collections.OrderedDict()
R = 3
for _ in range(R):
    p.run(imgs)
    for site, ms in p.kernel_times(True):
        acc.setdefault(site, []).append(ms)
src = {}
def kernel_at(site):
    f, l = site.split(":"); l = int(l)
    lines = src.setdefault(f, open(os.path.join(ROOT, "multimodal_biometric_fingerprints_palms_b200", "csrc", f)).read().splitlines())
    for k in range(l - 1, max(l - 6, 0), -1):
        m = re.search(r"(\w+)(<[^>]*>)?<<<", lines[k])
        if m: return m.group(1)
        m = re.search(r"launch_(\w+)<", lines[k])
        if m: return m.group(1)
        m = re.search(r"cudaLaunchKernelEx\(&\w+, (\w+)", lines[k])
        if m: return m.group(1)
    return "?"
tot = collections.OrderedDict(); grand = 0
for site, v in acc.items():
    k = kernel_at(site); ms = sum(v) / R
    t = tot.setdefault(k, [0.0, 0]); t[0] += ms; t[1] += len(v) // R; grand += ms
for k, (ms, c) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:22s} n={c:2d} {ms:8.3f} ms {100*ms/grand:5.1f}%")
print(f"total {grand:.2f} ms for {n} images {H}x{W} -> {n/grand*1e3:.0f} img/s")
