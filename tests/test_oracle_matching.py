"""CPU: the matcher oracle against the golden vectors minted from the reference's own `src/matching/match.py`
(oracle/make_golden_matching.py), and the host restatement of NumPy's SeedSequence + PCG64 against NumPy."""
import ctypes
import os

import numpy as np

from conftest import GOLDEN


def load_matching_golden():
    z = np.load(os.path.join(GOLDEN, "matching.npz"))
    g = {k: z[k] for k in z.files}
    off = np.concatenate([[0], np.cumsum(g["counts"])])
    g["tpl"] = [g["templates"][off[i]:off[i + 1]] for i in range(len(g["counts"]))]
    return g


def param_kw(ps):
    return dict(dist_thresh=float(ps[0]), orient_thresh_deg=float(ps[1]), use_type=bool(ps[2]), ransac_iter=int(ps[3]),
                min_inliers=int(ps[4]), stop_inlier_ratio=float(ps[5]), cross_check=bool(ps[6]))


def test_oracle_matches_reference_vectors():
    from oracle import ref_matching as rm
    g = load_matching_golden()
    idx = list(range(0, len(g["final_score"]), 5))           # every 5th case keeps the CPU suite short
    for c in idx:
        a, b = g["case_pair"][c]
        r = rm.match_minutiae_pair(g["tpl"][a], g["tpl"][b], **param_kw(g["param_sets"][g["case_param"][c]]))
        assert r["final_score"] == g["final_score"][c], c
        assert r["inlier_ratio"] == g["inlier_ratio"][c], c
        assert len(r["matches"]) == g["n_matches"][c], c
        for k, (i, j, s) in enumerate(r["matches"]):
            assert (i, j) == tuple(g["matches"][c][k]) and s == g["match_scores"][c][k], (c, k)
        assert float(r["theta"]) == g["theta"][c] and np.array_equal(np.asarray(r["t"], float), g["t"][c]), c


def test_seed_uniforms_equal_numpy_default_rng():
    from multimodal_biometric_fingerprints_palms_b200 import _native
    lib = _native.load()
    for seed0, n in ((42, 800), (0, 5), (2 ** 32 - 2, 4), (2 ** 40 + 5, 3)):
        out = np.zeros((n, 2))
        assert lib.fpb_match_seed_uniforms(ctypes.c_uint64(seed0), n, out.ctypes.data) == 0
        ref = np.array([np.random.default_rng(seed0 + i).random(2) for i in range(n)])
        assert np.array_equal(out, ref), seed0


def test_choice_model_is_one_uniform_and_a_cdf_search():
    """`rng.choice(a, p=p)` == a[cdf.searchsorted(rng.random(), 'right')] - the model k_match_prep implements."""
    rng = np.random.default_rng(5)
    for _ in range(300):
        n = int(rng.integers(1, 90))
        w = rng.uniform(0.05, 2.0, n)
        seed = int(rng.integers(0, 5000))
        pick = np.random.default_rng(seed).choice(np.arange(n), p=w / np.sum(w))
        p = w / np.sum(w)
        cdf = p.cumsum()
        cdf /= cdf[-1]
        assert pick == np.searchsorted(cdf, np.random.default_rng(seed).random(), side="right")
