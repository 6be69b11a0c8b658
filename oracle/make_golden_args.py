"""Golden vectors for the NON-DEFAULT keyword arguments of `compute_orientation_map` (orientation.py:9-14) and
`smooth_fingerprint_skeleton` (fingerprint_preprocess.py:141-144), frozen from the reference's OWN modules.

TEST INFRASTRUCTURE.  Run in the build container only (needs `/root/reference`):

    python -m oracle.make_golden_args       # writes tests/golden/kwargs_128x112.npz + .json

The reference's hot path only ever passes the defaults; these cases pin the other values the public functions accept
(block sizes 8 / 12 / 24, Gaussian radii outside the frozen tables, `invert_if_needed=False`, sigmas SciPy skips, zero
or five diffusion steps).  Each case asserts `oracle.ref_pipeline` == reference bit for bit before it is stored.
"""
from __future__ import annotations

import json
import os

import numpy as np

from oracle.make_golden import GOLDEN, _import_reference, _same

H, W, SEED = 128, 112, 5
POST_CASE = "polyu_320x240_s0"

ORIENT_CASES = [
    # name, block_size, smooth_sigma, invert_if_needed, smooth_orientation_sigma, with mask
    ("bs8_s2_so1", 8, 2.0, True, 1.0, True),
    ("noinvert", 16, 3.0, False, 3.0, True),
    ("bs24_s4_so2_nomask", 24, 4.0, True, 2.0, False),
    ("bs12_s1_so0", 12, 1.0, True, 0.0, True),
    ("bs16_s0_so5", 16, 0.0, True, 5.0, True),
    ("bs20_s2p6_neg", 20, 2.6, True, -1.0, False),
]
SMOOTH_CASES = [
    # name, sigma, diffusion_iter, contrast_boost
    ("s1_i2_b1p5", 1.0, 2, 1.5),
    ("i0", 1.4, 0, 1.25),
    ("s2_i5_b1", 2.0, 5, 1.0),
    ("s0p7_i1_b2", 0.7, 1, 2.0),
]


def make_inputs():
    """A small ridge print, its elliptical foreground mask and a ridge / valley binary image."""
    from multimodal_biometric_fingerprints_palms_b200 import synth
    img = synth.ridge_image(H, W, seed=SEED, period=7)
    yy, xx = np.mgrid[0:H, 0:W]
    inside = ((xx - W / 2.0) / (0.42 * W)) ** 2 + ((yy - H / 2.0) / (0.46 * H)) ** 2 <= 1.0
    mask = inside.astype(np.uint8) * 255
    binary = ((img < 128) & inside).astype(np.uint8) * 255
    return img, mask, binary


FLOAT_CASES = ["unit_f32", "int16_range", "f64_0_255_bs8", "half_at_max", "over_half_at_max"]


def float_cases(img, mask, binary):
    """(name, non-uint8 image, keyword arguments) - rebuilt identically by tests/test_kwargs.py from the stored uint8 inputs."""
    n = img.size
    flat = (binary.ravel() > 0)
    order = np.argsort(~flat, kind="stable")                      # set pixels first, then the others in raster order
    def with_ones(k):
        f = np.zeros(n, np.float32); f[order[:k]] = 1.0
        return f.reshape(img.shape)
    return [
        ("unit_f32", (img / 255.0).astype(np.float32), dict(mask=mask)),
        ("int16_range", img.astype(np.int16) * 4 - 100, dict(mask=mask)),
        ("f64_0_255_bs8", img.astype(np.float64), dict(block_size=8, smooth_sigma=2.0, mask=mask)),
        ("half_at_max", with_ones(n // 2), dict()),               # median 0.5 < max: inverted
        ("over_half_at_max", with_ones(n // 2 + 1), dict()),      # median == max: not inverted
    ]


def main():
    fp, ori, _, pp = _import_reference()
    from oracle import ref_pipeline as rp
    img, mask, binary = make_inputs()
    out = {"img": img, "mask": mask, "binary": binary}
    for name, bs, ss, inv, sos, use_mask in ORIENT_CASES:
        kw = dict(block_size=bs, smooth_sigma=ss, invert_if_needed=inv, smooth_orientation_sigma=sos,
                  mask=mask if use_mask else None)
        r_blk, r_oimg, r_rel = ori.compute_orientation_map(img, **kw)
        o_blk, o_oimg, o_rel = rp.compute_orientation_map(img, **kw)
        _same(o_blk, r_blk, f"{name}: orient_blocks")
        _same(o_oimg, r_oimg, f"{name}: orient_img")
        _same(o_rel, r_rel, f"{name}: rel_img")
        out[f"orient_{name}_blocks"], out[f"orient_{name}_img"], out[f"orient_{name}_rel"] = r_blk, r_oimg, r_rel
        print(f"[golden kwargs] orientation {name}: grid {r_blk.shape}, oracle == reference")
    for name, sg, it, boost in SMOOTH_CASES:
        r = fp.smooth_fingerprint_skeleton(binary, sigma=sg, diffusion_iter=it, contrast_boost=boost)
        _same(rp.smooth_fingerprint_skeleton(binary, sigma=sg, diffusion_iter=it, contrast_boost=boost), r, f"{name}: smooth")
        out[f"smooth_{name}"] = r
        print(f"[golden kwargs] smooth {name}: {int((r > 0).sum())} px set, oracle == reference")
    # compute_orientation_map on non-uint8 images (orientation.py:21-28)
    for name, fimg, kw in float_cases(img, mask, binary):
        r_blk, r_oimg, r_rel = ori.compute_orientation_map(fimg, **kw)
        o_blk, o_oimg, o_rel = rp.compute_orientation_map(fimg, **kw)
        _same(o_blk, r_blk, f"{name}: orient_blocks"); _same(o_oimg, r_oimg, f"{name}: orient_img"); _same(o_rel, r_rel, f"{name}: rel_img")
        out[f"float_{name}_blocks"], out[f"float_{name}_img"], out[f"float_{name}_rel"] = r_blk, r_oimg, r_rel
        print(f"[golden kwargs] orientation of a {fimg.dtype} image ({name}): oracle == reference")
    # segment_fingerprint on a colour image (fingerprint_preprocess.py:94): three differently degraded copies as B, G, R
    from multimodal_biometric_fingerprints_palms_b200 import synth
    bgr = np.stack([synth.ridge_image(160, 144, seed=SEED + 1 + c, period=8) for c in range(3)], axis=-1)
    r_seg, r_mask = fp.segment_fingerprint(bgr)
    o_seg, o_mask = rp.segment_fingerprint(bgr)
    _same(o_seg, r_seg, "bgr: segmented"); _same(o_mask, r_mask, "bgr: mask")
    out["bgr"], out["bgr_segmented"], out["bgr_mask"] = bgr, r_seg, r_mask
    print(f"[golden kwargs] segment_fingerprint(BGR): crop {r_seg.shape}, oracle == reference")
    # postprocess_minutiae's `gray` argument (post_processing.py:71, 93) on the first golden print: None (= sk_bin), the
    # segmented grey image, and the JPEG-decoded skeleton file next to the clean skeleton
    z = np.load(os.path.join(GOLDEN, f"{POST_CASE}.npz"))
    with open(os.path.join(GOLDEN, f"{POST_CASE}.json")) as f:
        raw = json.load(f)["raw_minutiae"]
    skel = z["skeleton"]
    post = {}
    for name, gray in (("none", None), ("segmented", z["segmented"]), ("skeleton_file", z["skeleton_file"])):
        r = pp.postprocess_minutiae([dict(m) for m in raw], skel, gray, None)
        o = rp.postprocess_minutiae([dict(m) for m in raw], skel, gray, None)
        assert o == r, f"postprocess gray={name}: oracle != reference"
        post[name] = r
        print(f"[golden kwargs] postprocess gray={name}: {len(r)} refined, oracle == reference")
    np.savez_compressed(os.path.join(GOLDEN, f"kwargs_{H}x{W}.npz"), **out)
    with open(os.path.join(GOLDEN, f"kwargs_{H}x{W}.json"), "w") as f:
        json.dump({"orientation": [list(c) for c in ORIENT_CASES], "smooth": [list(c) for c in SMOOTH_CASES],
                   "h": H, "w": W, "seed": SEED, "post_case": POST_CASE, "post_gray": post}, f, indent=1)


if __name__ == "__main__":
    main()
