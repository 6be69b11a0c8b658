"""U-Net++ segmenter inference: accuracy against the torch fp32 oracle and throughput of the tcgen05
engine.   python tools/bench_unet.py [batch [H W]]   -> one JSON line"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_biometric_fingerprints_palms_b200.preprocessing.segmentation.model import NestedUNet   # noqa: E402
from oracle.unetpp_ref import seeded_input, seeded_model                                               # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (256, 256)          # config_segmentation.yml image_size
ora = seeded_model(0)
x = seeded_input(3, n, H, W)
torch.set_num_threads(os.cpu_count() or 1)
t0 = time.perf_counter()
with torch.no_grad():
    want = ora(x[:2]).numpy()
cpu_s = (time.perf_counter() - t0) / 2
m = NestedUNet(max_batch=n)
m.load_state_dict(ora.state_dict())
xa = x.numpy()
got = m(xa)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); m(xa); ts.append(time.perf_counter() - t0)
err = float(np.abs(got[:2] - want).max())
# multiply-adds of the evaluated convolutions (conv4_0 is never evaluated)
f = [64, 128, 256, 512]
macs = 0
for lvl, (cin, cout) in enumerate([(3, f[0]), (f[0], f[1]), (f[1], f[2]), (f[2], f[3])]):
    macs += (H >> lvl) * (W >> lvl) * 9 * (cin * cout + cout * cout)
for lvl, cin, cout in [(0, f[0] + f[1], f[0]), (1, f[1] + f[2], f[1]), (2, f[2] + f[3], f[2]), (0, 2 * f[0] + f[1], f[0]),
                       (1, 2 * f[1] + f[2], f[1]), (0, 3 * f[0] + f[1], f[0])]:
    macs += (H >> lvl) * (W >> lvl) * 9 * (cin * cout + cout * cout)
best = min(ts)
print(json.dumps({"batch": n, "image": [H, W], "images_per_s": n / best, "ms_per_batch": 1e3 * best,
                  "gflop_per_image": 2 * macs / 1e9, "tflops_fp32_equivalent": 2 * macs * n / best / 1e12,
                  "tensor_core_tf32_tflops_issued": 3 * 2 * macs * n / best / 1e12,
                  "max_abs_err_vs_torch_fp32": err, "logit_scale": float(np.abs(want).max()),
                  "torch_cpu_fp32_s_per_image": cpu_s, "cpu_threads": torch.get_num_threads(),
                  "launches_total_tc": m.launches()}))
