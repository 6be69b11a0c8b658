"""What is at stake in the scikit-image thinning-table question (VERDICT r1 task 1b): the oracle's K7..K9 on the same masks
with (a) the built-in table derived from Zhang & Suen (1984) and (b) that table plus the four suspected "staircase" corner
deletions (codes 10, 40, 130, 160 -> 3).  CPU only (oracle).   python tools/thin_table_sensitivity.py > profiles/r02_thin_table_sensitivity.md"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multimodal_biometric_fingerprints_palms_b200 import synth        # noqa: E402
from oracle import ref_pipeline as rp, skimage_compat as sc            # noqa: E402

base = sc.zhang_suen_table()
alt = base.copy(); alt[[10, 40, 130, 160]] = 3
cases = [("polyu 320x240 seed %d" % s, synth.ridge_image(320, 240, seed=s, period=None)) for s in range(12)]
cases += [("degraded 512x512 seed %d" % s, synth.degraded_image(512, 512, seed=s)) for s in range(3)]
print("# Sensitivity of skeleton / minutiae to the thinning table (oracle, CPU)\n")
print("Built-in table: derived Zhang-Suen (40 non-zero entries).  Alternative: the same + codes 10, 40, 130, 160 deletable in both")
print("sub-iterations (the corner deletions scikit-image's literal table is suspected to contain).  Same boolean mask into")
print("`skeletonize` on both sides; then the reference's clean-up, JPEG hand-off, K8 and K9.\n")
print("| input | skeleton px (built-in) | px that differ | raw minutiae built-in / alt | raw in common | refined built-in / alt | refined in common |")
print("|---|---|---|---|---|---|---|")
tot = np.zeros(7)
for name, img in cases:
    r = rp.preprocess_fingerprint(img)
    gate = rp.thinning_gate(r["binary_smooth"], r["reliability"]) if "binary_smooth" in r else None
    if gate is None:
        b = rp.smooth_fingerprint_skeleton(r["binary"])
        _, _, rel = rp.compute_orientation_map(r["segmented"], mask=r["mask"])
        gate = rp.thinning_gate(b, rel)
    out = []
    for tab in (base, alt):
        sk = rp.thin_and_clean(gate, tab)
        f = rp.skeleton_file_roundtrip(sk)
        raw = rp.extract_minutiae(f)
        ref = rp.postprocess_minutiae([dict(m) for m in raw], f, f, None)
        out.append((sk, {(m["x"], m["y"], m["type"]) for m in raw}, {(m["x"], m["y"], m["type"]) for m in ref}))
    (s0, r0, f0), (s1, r1, f1) = out
    row = [int((s0 > 0).sum()), int((s0 != s1).sum()), len(r0), len(r1), len(r0 & r1), len(f0), len(f1), len(f0 & f1)]
    print(f"| {name} | {row[0]} | {row[1]} | {row[2]} / {row[3]} | {row[4]} | {row[5]} / {row[6]} | {row[7]} |")
    tot += np.array([row[0], row[1], row[2], row[3], row[4], row[5] + row[6], 2 * row[7]])
print(f"\nTotals: {int(tot[1])} of {int(tot[0])} skeleton pixels differ ({100 * tot[1] / tot[0]:.2f} %); raw minutiae "
      f"{int(tot[2])} vs {int(tot[3])}, {int(tot[4])} in common; refined minutiae agreement (2 x common / sum of sizes) "
      f"{100 * tot[6] / max(tot[5], 1):.1f} %.")
print("\nThe CUDA kernels take the table as DATA (`fpb_set_thin_table`, `FPB200_THIN_TABLE=<file>`), reproduce either column of this")
print("table bit-for-bit (tests/test_gpu_stages.py::test_custom_thinning_table_is_data), and where scikit-image is installed the")
print("first handle compares against the real package and raises on any difference (selfcheck.py).")
