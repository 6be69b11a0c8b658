"""Developer tool: the fused K1..K9 run on a batch - the command line of the ncu captures of the second-tier kernels.
   python tools/run_batch.py [batch [H W [reps]]]"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (320, 240)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
base = synth.ridge_batch(min(n, 32), H, W, first_seed=500)
p = FingerprintPipeline(H, W, max_batch=n)
p.set_profiling(True)                 # single stream: the launch order of the capture is the stage order
imgs = np.stack([base[i % len(base)] for i in range(n)])
for _ in range(reps):
    p.run(imgs)
print("ok", sum(len(p.minutiae(i)) for i in range(n)))
