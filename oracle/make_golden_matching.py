"""Pin `oracle.ref_matching` against the reference's OWN matcher and freeze golden fixtures.

TEST INFRASTRUCTURE.  Build container only (needs `/root/reference`):

    python -m oracle.make_golden_matching       # writes tests/golden/matching.npz

* imports `/root/reference/src/matching/match.py` unmodified (with the cosmetic colorama shim);
* the reference consumes its hypotheses with `concurrent.futures.as_completed`, whose order depends on thread
  timing; for the whole-function vectors `as_completed` is replaced IN THE REFERENCE MODULE'S NAMESPACE by the
  identity (futures in submission = seed order) - the per-hypothesis functions (`ransac_worker`,
  `match_with_transform`, `compute_descriptor_weight`) are compared as shipped;
* asserts the oracle reproduces every output bit-for-bit, then stores templates, parameter sets and the
  reference's results.
"""
from __future__ import annotations

import math
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden", "matching.npz")

PARAM_SETS = [   # dist_thresh, orient_thresh_deg, use_type, ransac_iter, min_inliers, stop_inlier_ratio, cross_check
    (10.0, 12.0, 1, 300, 8, 0.25, 1),     # defaults of match_minutiae_pair (match.py:219-231)
    (30.0, 30.0, 1, 300, 6, 0.15, 1),     # compute_frr as called by match_features.py:124-131
    (15.0, 10.0, 1, 300, 12, 0.15, 1),    # compute_far as called by match_features.py:141-149
    (22.0, 38.0, 1, 800, 7, 0.15, 1),     # config_matching.yml
    (30.0, 30.0, 0, 50, 3, 0.15, 0),      # demo-sized, no type gate, no cross-check
]


def build_templates():
    from oracle import ref_matching as rm
    tpl = []
    for user in range(6):
        base = rm.synthetic_template(100 + user, n=40 + 4 * user)
        tpl.append(base)
        tpl.append(rm.perturbed_copy(base, 200 + user, angle_deg=5.0 + user, shift=(6.0 + user, -4.0), jitter=1.0))
        tpl.append(rm.perturbed_copy(base, 300 + user, angle_deg=-9.0, shift=(-7.0, 5.0 + user), jitter=1.6, drop=0.3))
    tpl.append(rm.synthetic_template(900, n=7))                       # fewer than 8 minutiae: early reject 1
    tpl.append(np.zeros((0, 7)))                                      # empty template
    wide = rm.synthetic_template(901, n=30)
    wide[:, :2] *= 3.0                                                # spread differs by > 35: early reject 2
    tpl.append(wide)
    one_type = rm.synthetic_template(902, n=25)
    one_type[:, 2] = 1.0                                              # only bifurcations: empty same-type picks
    tpl.append(one_type)
    tpl.append(rm.synthetic_template(903, n=60))
    return tpl


def build_pairs(n_tpl):
    pairs = []
    for user in range(6):
        b = 3 * user
        pairs += [(b, b + 1), (b, b + 2), (b + 1, b + 2), (b + 1, b)]  # genuine, one reversed
    pairs += [(0, 3), (1, 7), (5, 9), (2, 16), (4, 12), (10, 14)]      # impostors
    pairs += [(0, 0), (18, 1), (1, 18), (19, 2), (2, 19), (20, 3), (3, 20), (21, 4), (4, 21), (22, 5), (22, 22)]
    assert max(max(p) for p in pairs) < n_tpl
    return pairs


def main():
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(HERE, "ref_shim"))
    sys.path.insert(0, REF)
    os.chdir(tempfile.mkdtemp(prefix="ref_import_"))
    from src.matching import match as ref
    from oracle import ref_matching as rm

    tpl = build_templates()
    pairs = build_pairs(len(tpl))

    # ---- per-hypothesis functions as shipped
    checked = 0
    for (a, b) in pairs[:12]:
        A, B = tpl[a], tpl[b]
        if len(A) == 0 or len(B) == 0:
            continue
        wA = np.array([ref.compute_descriptor_weight(m) for m in A])
        wB = np.array([ref.compute_descriptor_weight(m) for m in B])
        assert np.array_equal(wA, np.array([rm.compute_descriptor_weight(m) for m in A]))
        for seed in range(42, 42 + 40):
            r = ref.ransac_worker((A, B, 30.0, math.radians(30.0), 6, True, wA, wB, seed))
            o = rm.ransac_worker(A, B, 30.0, math.radians(30.0), 6, True, wA, wB, seed)
            assert r["score"] == o["score"] and r["inliers"] == o["inliers"], (a, b, seed)
            if "theta" in r:
                assert r["theta"] == o["theta"] and np.array_equal(r["t"], o["t"])
            checked += 1
    print(f"ransac_worker: {checked} hypotheses identical")

    # ---- whole function with hypotheses consumed in seed order
    ref.as_completed = lambda fs: fs
    res = {"final_score": [], "inlier_ratio": [], "theta": [], "t": [], "n_matches": [], "matches": [], "match_scores": [],
           "case_pair": [], "case_param": []}
    maxm = max(len(t) for t in tpl)
    for pi, ps in enumerate(PARAM_SETS):
        for (a, b) in pairs:
            kw = dict(dist_thresh=ps[0], orient_thresh_deg=ps[1], use_type=bool(ps[2]), ransac_iter=ps[3], min_inliers=ps[4],
                      stop_inlier_ratio=ps[5], cross_check=bool(ps[6]))
            r = ref.match_minutiae_pair(tpl[a], tpl[b], thread_workers=1, **kw)
            o = rm.match_minutiae_pair(tpl[a], tpl[b], **kw)
            assert r["final_score"] == o["final_score"] and r["inlier_ratio"] == o["inlier_ratio"], (pi, a, b)
            assert r["matches"] == o["matches"], (pi, a, b)
            assert float(r["theta"]) == float(o["theta"]) and np.array_equal(np.asarray(r["t"]), np.asarray(o["t"]))
            mm = np.full((maxm, 2), -1, np.int32)
            ms = np.zeros(maxm)
            for k, (i, j, s) in enumerate(r["matches"]):
                mm[k] = (i, j)
                ms[k] = s
            res["final_score"].append(r["final_score"]); res["inlier_ratio"].append(r["inlier_ratio"])
            res["theta"].append(float(r["theta"])); res["t"].append(np.asarray(r["t"], float))
            res["n_matches"].append(len(r["matches"])); res["matches"].append(mm); res["match_scores"].append(ms)
            res["case_pair"].append((a, b)); res["case_param"].append(pi)
    n = len(res["final_score"])
    nz = int(np.count_nonzero(res["final_score"]))
    print(f"match_minutiae_pair: {n} cases identical to the oracle, {nz} with a non-zero score")
    np.savez_compressed(OUT, templates=np.vstack(tpl), counts=np.array([len(t) for t in tpl], np.int32),
                        param_sets=np.array(PARAM_SETS, np.float64),
                        **{k: np.array(v) for k, v in res.items()})
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
