"""Pin the oracle against the reference's OWN modules and freeze golden fixtures.

TEST INFRASTRUCTURE.  Run in the build container only (needs `/root/reference`,
which does not exist on the GPU box):

    python -m oracle.make_golden            # writes tests/golden/*.npz + *.json

What it does
* puts `/root/reference` and `oracle/ref_shim` (a cosmetic `colorama` stand-in and
  a `skimage` shim that routes to `oracle.skimage_compat`; scikit-image is absent
  here) on `sys.path` and imports the reference's
  `src.preprocessing.fingerprint_preprocess`, `src.preprocessing.orientation`,
  `src.features.extract_features`, `src.features.post_processing` UNMODIFIED;
* runs every hot-path function of the reference on the seeded synthetic inputs
  of SURVEY.md section 8(d) and asserts that `oracle.ref_pipeline` reproduces each output
  bit-for-bit (same library calls, so anything else is an oracle bug);
* stores inputs + the reference's outputs as compressed fixtures so that the
  pinning travels to the GPU box (`tests/test_oracle_golden.py`).

The scikit-image functions themselves are NOT pinned by this (they are the
restated ones on both sides) - see oracle/skimage_compat.py.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
GOLDEN = os.path.join(REPO, "tests", "golden")

CASES = [
    # name, h, w, kind, seed
    ("polyu_320x240_s0", 320, 240, "ridge", 0),
    ("polyu_240x320_s1", 240, 320, "ridge", 1),
    ("polyu_320x240_s7", 320, 240, "ridge_rand", 7),
    ("nist_256x256_s3", 256, 256, "degraded", 3),
]


def _import_reference():
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(HERE, "ref_shim"))
    sys.path.insert(0, REF)
    # the reference's driver modules create log directories relative to the cwd at
    # import time (extract_features.py:19-28) - keep that out of the repo
    os.chdir(tempfile.mkdtemp(prefix="ref_import_"))
    from src.preprocessing import fingerprint_preprocess as fp
    from src.preprocessing import orientation as ori
    from src.features import extract_features as ef
    from src.features import post_processing as pp
    return fp, ori, ef, pp


def _make_input(kind, h, w, seed):
    from multimodal_biometric_fingerprints_palms_b200 import synth
    if kind == "ridge":
        return synth.ridge_image(h, w, seed=seed)
    if kind == "ridge_rand":
        return synth.ridge_image(h, w, seed=seed, period=None)
    if kind == "degraded":
        return synth.degraded_image(h, w, seed=seed)
    raise ValueError(kind)


def _same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or a.dtype != b.dtype or not np.array_equal(a, b, equal_nan=True):
        raise AssertionError(f"oracle != reference at {what}: shapes {a.shape}/{b.shape} "
                             f"dtypes {a.dtype}/{b.dtype}")


def main():
    fp, ori, ef, pp = _import_reference()
    from oracle import ref_pipeline as rp

    os.makedirs(GOLDEN, exist_ok=True)
    index = []
    for name, h, w, kind, seed in CASES:
        img = _make_input(kind, h, w, seed)
        # ---- reference, stage by stage -------------------------------------------------
        ref = fp.preprocess_fingerprint(img)
        r_norm = fp.normalize_image(img)
        r_den = fp.denoise_image(r_norm)
        r_seg, r_mask = fp.segment_fingerprint(r_den)
        r_bin = fp.binarize(r_seg)
        r_blk, r_oimg, r_rel = ori.compute_orientation_map(r_seg, block_size=16, smooth_sigma=3.0,
                                                           invert_if_needed=True,
                                                           smooth_orientation_sigma=3.0, mask=r_mask)
        r_smooth = fp.smooth_fingerprint_skeleton(r_bin)
        r_skel = fp.thinning_and_cleaning(r_smooth, r_oimg, r_rel)
        r_raw = ef.extract_minutiae(r_skel)
        r_ref = pp.postprocess_minutiae([dict(m) for m in r_raw], r_skel, r_skel, None)
        # K9's inner call: orientation of the skeleton itself, defaults, no mask
        k_blk, k_oimg, k_rel = ori.compute_orientation_map(r_skel)
        # ---- the reference's CLI hand-off: run_preprocessing.py:137-140 writes <base>_skeleton.jpg, and the
        #      reference's OWN process_image (extract_features.py:74-108) reads it and writes <base>_minutiae.json
        import cv2
        with tempfile.TemporaryDirectory() as td:
            cv2.imwrite(os.path.join(td, "g_skeleton.jpg"), r_skel)
            f_skel = cv2.imread(os.path.join(td, "g_skeleton.jpg"), cv2.IMREAD_GRAYSCALE)
            ef.process_image("g_skeleton.jpg", td, td, None)
            with open(os.path.join(td, "g_minutiae.json")) as f:
                f_ref = json.load(f)
        f_raw = ef.extract_minutiae(f_skel)
        f_blk, f_oimg, f_rel = ori.compute_orientation_map(f_skel)
        for key, val in (("normalized", r_norm), ("denoised", r_den), ("segmented", r_seg),
                         ("mask", r_mask), ("binary", r_bin), ("skeleton", r_skel)):
            _same(ref[key], val, f"{name}: reference self-consistency {key}")

        # ---- oracle must reproduce every stage bit-for-bit ------------------------------
        _same(rp.normalize_image(img), r_norm, f"{name}:normalize_image")
        _same(rp.denoise_image(r_norm), r_den, f"{name}:denoise_image")
        o_seg, o_mask = rp.segment_fingerprint(r_den)
        _same(o_seg, r_seg, f"{name}:segment.segmented")
        _same(o_mask, r_mask, f"{name}:segment.mask")
        _same(rp.binarize(r_seg), r_bin, f"{name}:binarize")
        o_blk, o_oimg, o_rel = rp.compute_orientation_map(r_seg, mask=r_mask)
        _same(o_blk, r_blk, f"{name}:orient_blocks")
        _same(o_oimg, r_oimg, f"{name}:orient_img")
        _same(o_rel, r_rel, f"{name}:rel_img")
        _same(rp.smooth_fingerprint_skeleton(r_bin), r_smooth, f"{name}:smooth")
        _same(rp.thinning_and_cleaning(r_smooth, r_oimg, r_rel), r_skel, f"{name}:skeleton")
        o_raw = rp.extract_minutiae(r_skel)
        assert o_raw == r_raw, f"{name}: raw minutiae differ"
        o_ref = rp.postprocess_minutiae([dict(m) for m in o_raw], r_skel, r_skel, None)
        assert o_ref == r_ref, f"{name}: refined minutiae differ"
        o_all = rp.enhance_to_minutiae(img, handoff="memory")
        assert o_all["minutiae"] == r_ref and o_all["raw_minutiae"] == r_raw
        _same(o_all["skeleton"], r_skel, f"{name}:pipeline skeleton")
        o_file = rp.enhance_to_minutiae(img)                      # default: through the JPEG file, as the CLI
        _same(o_file["skeleton_file"], f_skel, f"{name}:skeleton as read back from the JPEG")
        assert o_file["raw_minutiae"] == f_raw, f"{name}: raw minutiae after the hand-off differ"
        assert o_file["minutiae"] == f_ref, f"{name}: the reference's own <base>_minutiae.json differs"
        from oracle.jpeg_fdct import jpeg_roundtrip
        _same(jpeg_roundtrip(r_skel), f_skel, f"{name}: restated quality-95 codec")

        np.savez_compressed(
            os.path.join(GOLDEN, f"{name}.npz"),
            img=img, normalized=r_norm, denoised=r_den, segmented=r_seg, mask=r_mask,
            binary=r_bin, orient_blocks=r_blk, orient_img=r_oimg, reliability=r_rel,
            binary_smooth=r_smooth, skeleton=r_skel,
            skel_orient_img=k_oimg, skel_coherence=k_rel,
            skeleton_file=f_skel, file_orient_img=f_oimg, file_coherence=f_rel)
        with open(os.path.join(GOLDEN, f"{name}.json"), "w") as f:
            json.dump({"raw_minutiae": r_raw, "minutiae": r_ref, "raw_minutiae_file": f_raw, "minutiae_file": f_ref}, f, indent=1)
        index.append({"name": name, "h": h, "w": w, "kind": kind, "seed": seed,
                      "n_raw": len(r_raw), "n_refined": len(r_ref), "n_refined_file": len(f_ref),
                      "file_vs_memory_refined_common": len({(m["x"], m["y"]) for m in r_ref} & {(m["x"], m["y"]) for m in f_ref}),
                      "file_nonzero_px": int((f_skel > 0).sum()), "memory_nonzero_px": int((r_skel > 0).sum()),
                      "crop": list(r_seg.shape)})
        print(f"[golden] {name}: crop {r_seg.shape}, raw {len(r_raw)}, refined {len(r_ref)} - "
              f"oracle == reference on all stages")

    import cv2, scipy
    with open(os.path.join(GOLDEN, "index.json"), "w") as f:
        json.dump({"cases": index,
                   "made_with": {"numpy": np.__version__, "scipy": scipy.__version__,
                                 "opencv": cv2.__version__,
                                 "skimage": "absent - oracle.skimage_compat on both sides"}},
                  f, indent=1)


if __name__ == "__main__":
    main()
