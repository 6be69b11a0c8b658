"""GPU parity of the RANSAC matcher (k_match_prep / k_match_pairs through the C ABI) against the golden vectors of the
reference's matcher and against the oracle on fresh templates.

Tolerances: index sets (matches, n_matches, best hypothesis) EXACT; float64 scores / transform within 1e-9 relative
(CUDA's double-precision cos/sin/exp/pow/atan2 are 1-2 ulp from glibc's, and the refinement uses the closed-form
Kabsch angle instead of LAPACK's 2x2 SVD)."""
import math

import numpy as np
import pytest

from test_oracle_matching import load_matching_golden, param_kw

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def close(a, b, what):
    assert math.isclose(float(a), float(b), rel_tol=RTOL, abs_tol=1e-12), f"{what}: {a!r} vs {b!r}"


def test_golden_vectors_all_cases():
    from multimodal_biometric_fingerprints_palms_b200.matching import MinutiaeMatcher
    g = load_matching_golden()
    m = MinutiaeMatcher(len(g["tpl"]), int(g["counts"].max()), 800)
    m.set_templates(g["tpl"])
    for pi, ps in enumerate(g["param_sets"]):
        cases = np.nonzero(g["case_param"] == pi)[0]
        res, mm, ms = m.match(g["case_pair"][cases], True, **param_kw(ps))
        for k, c in enumerate(cases):
            tag = f"param set {pi} pair {tuple(g['case_pair'][c])}"
            assert res["n_matches"][k] == g["n_matches"][c], tag
            n = int(g["n_matches"][c])
            assert np.array_equal(mm[k, :n], g["matches"][c][:n]), tag
            np.testing.assert_allclose(ms[k, :n], g["match_scores"][c][:n], rtol=RTOL, atol=0, err_msg=tag)
            close(res["final_score"][k], g["final_score"][c], tag + " final_score")
            close(res["inlier_ratio"][k], g["inlier_ratio"][c], tag + " inlier_ratio")
            close(res["theta"][k], g["theta"][c], tag + " theta")
            close(res["tx"][k], g["t"][c][0], tag + " tx")
            close(res["ty"][k], g["t"][c][1], tag + " ty")
    m.close()


def test_fresh_templates_against_oracle():
    from oracle import ref_matching as rm
    from multimodal_biometric_fingerprints_palms_b200.matching import match_pairs
    tpl, pairs = [], []
    for u in range(8):
        base = rm.synthetic_template(5000 + u, n=30 + 4 * u)
        tpl += [base, rm.perturbed_copy(base, 6000 + u, angle_deg=3.0 * u - 10, shift=(4.0 - u, 2.0 * u), jitter=0.8 + 0.1 * u)]
        pairs.append((2 * u, 2 * u + 1))
    pairs += [(0, 5), (3, 8), (7, 12), (15, 2)]
    kw = dict(dist_thresh=22.0, orient_thresh_deg=38.0, ransac_iter=400, min_inliers=7, stop_inlier_ratio=0.15)
    got = match_pairs(tpl, pairs, **kw)
    for (a, b), r in zip(pairs, got):
        o = rm.match_minutiae_pair(tpl[a], tpl[b], **kw)
        assert [(i, j) for i, j, _ in r["matches"]] == [(i, j) for i, j, _ in o["matches"]], (a, b)
        close(r["final_score"], o["final_score"], f"{a},{b} final_score")
        close(r["theta"], o["theta"], f"{a},{b} theta")
        np.testing.assert_allclose(r["t"], np.asarray(o["t"], float), rtol=RTOL, atol=1e-9)
    assert sum(r["final_score"] > 0 for r in got[:8]) >= 6          # the genuine pairs do match


def test_match_minutiae_pair_signature_and_edge_cases():
    from oracle import ref_matching as rm
    from multimodal_biometric_fingerprints_palms_b200.matching.match import match_minutiae_pair
    assert match_minutiae_pair(None, [[1, 2, 0, 0.1]]) == {"final_score": 0.0, "inlier_ratio": 0.0, "matches": []}
    a = rm.synthetic_template(1, 40)
    r = match_minutiae_pair(a, np.zeros((0, 7)))
    assert r["final_score"] == 0.0 and r["matches"] == [] and r["theta"] == 0.0 and list(r["t"]) == [0.0, 0.0]
    r = match_minutiae_pair(a.tolist(), a[:7].tolist(), thread_workers=2, debug=True)      # lists, < 8 minutiae
    assert r["final_score"] == 0.0 and r["inlier_ratio"] == 0.0
    r = match_minutiae_pair(a, a)
    o = rm.match_minutiae_pair(a, a)
    assert len(r["matches"]) == len(o["matches"]) == 40
    close(r["final_score"], o["final_score"], "self match")
    with pytest.raises(Exception, match="type column"):
        bad = a.copy(); bad[3, 2] = 2.0
        match_minutiae_pair(bad, a)


def test_large_templates_and_iteration_cap():
    """up to 256 minutiae per template (two KD-tree leaves in sklearn: ties aside the nearest neighbour is the same)."""
    from oracle import ref_matching as rm
    from multimodal_biometric_fingerprints_palms_b200.matching import match_pairs
    a = rm.synthetic_template(77, n=200, size=(900, 700))
    b = rm.perturbed_copy(a, 78, angle_deg=4.0, shift=(5.0, 3.0), jitter=0.7, drop=0.1, extra=0)
    kw = dict(dist_thresh=12.0, orient_thresh_deg=15.0, ransac_iter=120, min_inliers=10, stop_inlier_ratio=0.6)
    r = match_pairs([a, b], [(0, 1)], **kw)[0]
    o = rm.match_minutiae_pair(a, b, **kw)
    assert [(i, j) for i, j, _ in r["matches"]] == [(i, j) for i, j, _ in o["matches"]]
    close(r["final_score"], o["final_score"], "200-minutiae pair")
    with pytest.raises(Exception, match="ransac_iter"):
        from multimodal_biometric_fingerprints_palms_b200.matching import MinutiaeMatcher
        m = MinutiaeMatcher(2, 64, 100)
        m.set_templates([a[:50], b[:50]])
        m.match([(0, 1)], ransac_iter=101)


def test_frr_far_drivers_against_oracle_loop(tmp_path, monkeypatch):
    import random
    from oracle import ref_matching as rm
    from multimodal_biometric_fingerprints_palms_b200.matching.FRR import compute_frr
    from multimodal_biometric_fingerprints_palms_b200.matching.FAR import compute_far, sample_impostor_pairs
    monkeypatch.chdir(tmp_path)
    dataset = {}
    for u in range(5):
        base = rm.synthetic_template(40 + u, n=38 + u)
        dataset[f"{u:03d}"] = [base, rm.perturbed_copy(base, 50 + u, angle_deg=4.0 + u), rm.perturbed_copy(base, 60 + u, angle_deg=-6.0)]
    frr = compute_frr(dataset, dist_thresh=30, orient_thresh_deg=30, use_type=True, ransac_iter=120, min_inliers=6)
    want = []
    for u, s in dataset.items():
        for i in range(3):
            for j in range(i + 1, 3):
                want.append(rm.match_minutiae_pair(s[i], s[j], dist_thresh=30, orient_thresh_deg=30, use_type=True,
                                                   ransac_iter=120, min_inliers=6, stop_inlier_ratio=0.15)["final_score"])
    np.testing.assert_allclose(frr, want, rtol=RTOL, atol=1e-12)
    assert (tmp_path / "logs" / "genuine_match_stats.csv").exists()
    random.seed(3)
    far = compute_far(dataset, dist_thresh=15, orient_thresh_deg=10, use_type=True, ransac_iter=100, min_inliers=12,
                      impostor_sample_size=2)
    random.seed(3)
    want = []
    for u1, u2 in sample_impostor_pairs(list(dataset.keys()), 2):
        for a in dataset[u1]:
            for b in dataset[u2]:
                want.append(rm.match_minutiae_pair(a, b, dist_thresh=15, orient_thresh_deg=10, use_type=True, ransac_iter=100,
                                                   min_inliers=12, stop_inlier_ratio=0.15)["final_score"])
    np.testing.assert_allclose(far, want, rtol=RTOL, atol=1e-12)


def test_match_features_main_end_to_end(tmp_path, monkeypatch):
    """JSON files in the layout extract_features writes -> load_dataset -> FRR/FAR/ROC."""
    import json
    from oracle import ref_matching as rm
    from multimodal_biometric_fingerprints_palms_b200.matching import match_features as mf
    monkeypatch.chdir(tmp_path)
    root = tmp_path / "minutiae" / "cluster_0"
    root.mkdir(parents=True)
    for u in range(4):
        base = rm.synthetic_template(10 + u, n=36)
        for k, t in enumerate([base, rm.perturbed_copy(base, 20 + u)]):
            rec = [{"x": int(r[0]), "y": int(r[1]), "type": "ending" if r[2] == 0 else "bifurcation", "orientation": float(r[3]),
                    "quality": float(r[4]), "coherence": float(r[5]), "angular_stability": float(r[6])} for r in t]
            (root / f"{u:03d}_{k}_minutiae.json").write_text(json.dumps(rec, indent=2))
    ds = mf.load_dataset(str(tmp_path / "minutiae"), max_per_user=2)
    assert sorted(ds) == ["000", "001", "002", "003"] and all(len(v) == 2 and v[0].shape[1] == 7 for v in ds.values())
    out = mf.main(config_path=None, demo=True, minutiae_base=str(tmp_path / "minutiae"), show=False)
    assert len(out["genuine"]) == 4 and len(out["frr"]) == 30 and len(out["far"]) == 30
    assert out["far"][0] == 1.0 and out["frr"][0] == 0.0


def test_large_sweep_is_order_and_batch_independent():
    """3540 pairs (every ordered pair of 60 templates): the result of a pair does not
    depend on where it sits in the pair list or on what else is in the batch; spot checks against the oracle."""
    from oracle import ref_matching as rm
    from multimodal_biometric_fingerprints_palms_b200.matching import MinutiaeMatcher
    tpl = []
    for u in range(20):
        base = rm.synthetic_template(7000 + u, n=35 + u)
        tpl += [base, rm.perturbed_copy(base, 7100 + u, angle_deg=6.0 - u), rm.perturbed_copy(base, 7200 + u, angle_deg=u - 9.0, drop=0.25)]
    pairs = np.array([(a, b) for a in range(60) for b in range(60) if a != b], np.int32)
    m = MinutiaeMatcher(60, 64, 300)
    m.set_templates(tpl)
    kw = dict(dist_thresh=15, orient_thresh_deg=10, ransac_iter=300, min_inliers=12, stop_inlier_ratio=0.15)
    res, mm, ms = m.match(pairs, True, **kw)
    perm = np.random.default_rng(0).permutation(len(pairs))
    res2, mm2, ms2 = m.match(pairs[perm], True, **kw)
    assert np.array_equal(res2, res[perm])
    for k2, k in enumerate(perm):                         # entries past n_matches are unspecified
        n = int(res["n_matches"][k])
        assert np.array_equal(mm2[k2, :n], mm[k, :n]) and np.array_equal(ms2[k2, :n], ms[k, :n])
    res3, _, _ = m.match(pairs[100:164], False, **kw)
    assert np.array_equal(res3, res[100:164])
    genuine = [k for k, (a, b) in enumerate(pairs) if a // 3 == b // 3]
    assert (res["final_score"][genuine] > 0).mean() > 0.5
    impostor = [k for k, (a, b) in enumerate(pairs) if a // 3 != b // 3]
    assert (res["final_score"][impostor] > 0).mean() < 0.05
    for k in genuine[:6] + impostor[:3]:
        a, b = pairs[k]
        o = rm.match_minutiae_pair(tpl[a], tpl[b], **kw)
        assert len(o["matches"]) == res["n_matches"][k]
        close(res["final_score"][k], o["final_score"], f"pair {a},{b}")


def test_random_templates_and_parameters_against_oracle():
    """40 random pairs: template sizes 0..64 (incl. < 8 and empty), random rigid motions, random thresholds / iteration
    counts / stop ratios, with and without the type gate and the cross-check."""
    from oracle import ref_matching as rm
    from multimodal_biometric_fingerprints_palms_b200.matching import match_pairs
    rng = np.random.default_rng(11)
    for t in range(40):
        na = int(rng.choice([0, 5, 8, 9, 17, 33, 48, 64]))
        a = rm.synthetic_template(9000 + t, n=na) if na else np.zeros((0, 7))
        if na >= 8 and rng.random() < 0.7:
            b = rm.perturbed_copy(a, 9100 + t, angle_deg=float(rng.uniform(-25, 25)), shift=(float(rng.uniform(-15, 15)), float(rng.uniform(-15, 15))),
                                  jitter=float(rng.uniform(0.2, 2.5)), drop=float(rng.uniform(0, 0.4)), extra=int(rng.integers(0, 8)))[:64]
        else:
            nb = int(rng.choice([0, 7, 8, 30, 64]))
            b = rm.synthetic_template(9200 + t, n=nb) if nb else np.zeros((0, 7))
        kw = dict(dist_thresh=float(rng.choice([8.0, 10.0, 15.0, 22.0, 30.0])), orient_thresh_deg=float(rng.choice([8.0, 12.0, 30.0, 38.0])),
                  use_type=bool(rng.random() < 0.8), ransac_iter=int(rng.choice([37, 100, 300])), min_inliers=int(rng.choice([3, 6, 8, 12])),
                  stop_inlier_ratio=float(rng.choice([0.1, 0.15, 0.25, 0.9])), cross_check=bool(rng.random() < 0.7))
        r = match_pairs([a, b], [(0, 1)], **kw)[0]
        o = rm.match_minutiae_pair(a, b, **kw)
        tag = f"case {t}: {len(a)} x {len(b)} {kw}"
        assert [(i, j) for i, j, _ in r["matches"]] == [(i, j) for i, j, _ in o["matches"]], tag
        close(r["final_score"], o["final_score"], tag)
        close(r["inlier_ratio"], o["inlier_ratio"], tag)
        close(r["theta"], o.get("theta", 0.0), tag)
