"""Developer tool: run only denoise_image (K2) on a batch - the command line the ncu captures of the NLM kernels use.
   python tools/nlm_only.py [batch [H W [reps]]]"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (320, 240)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
base = np.stack([p for p in synth.ridge_batch(min(n, 32), H, W, first_seed=500)])
p = FingerprintPipeline(H, W, max_batch=n)
norm = p.normalize(np.stack([base[i % len(base)] for i in range(n)]))
for _ in range(reps):
    out = p.denoise(norm)
print("ok", out.shape, int(out.sum()))
