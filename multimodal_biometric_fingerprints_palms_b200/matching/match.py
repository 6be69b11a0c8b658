"""Drop-in for `src/matching/match.py` of the reference: the weighted RANSAC rigid matcher on the GPU.

`match_minutiae_pair` keeps the reference's signature and return dict (match.py:219-275); `MinutiaeMatcher` /
`match_pairs` are the batched forms FRR/FAR use (thousands of pairs per launch).  The arithmetic is
`k_match_prep` / `k_match_pairs` of libfpb200.so - there is no CPU path.  Hypotheses are consumed in seed order
(see include/fpb200_match.h); `thread_workers` and `debug` are accepted and ignored like unused knobs.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence

import numpy as np

from .. import _native
from .._native import FpbError, MatchParams, MatchResult

RESULT_DTYPE = np.dtype([("final_score", "f8"), ("inlier_ratio", "f8"), ("theta", "f8"), ("tx", "f8"), ("ty", "f8"),
                         ("n_matches", "i4"), ("best_iter", "i4")])
assert RESULT_DTYPE.itemsize == C.sizeof(MatchResult)


def _as_template(m) -> np.ndarray:
    """np.array(mins) in the layout of match_features.py:52-62; missing quality columns read as 0.0 (match.py:14-16)."""
    a = np.asarray(m, dtype=np.float64)
    if a.size == 0:
        return np.zeros((0, 7), np.float64)
    if a.ndim != 2 or a.shape[1] < 4:
        raise ValueError(f"a template must be [n, >=4] (x, y, type, orientation, ...), got {a.shape}")
    if a.shape[1] < 7:
        a = np.hstack([a, np.zeros((a.shape[0], 7 - a.shape[1]))])
    return np.ascontiguousarray(a[:, :7])


class MinutiaeMatcher:
    """One matcher handle on one GPU: upload templates once, match any list of index pairs."""

    def __init__(self, max_templates: int, max_minutiae: int = 64, max_iter: int = 800, device: int = 0):
        self._lib = _native.load()
        self._h = C.c_void_p()
        rc = self._lib.fpb_match_create(C.byref(self._h), device, max_templates, max_minutiae, max_iter)
        if rc != 0:
            raise FpbError(self._lib.fpb_match_last_error(None).decode())
        self.max_templates, self.max_iter = max_templates, max_iter
        self.max_minutiae = (max_minutiae + 3) & ~3
        self.n_templates = 0
        self._n_pairs = 0

    def _check(self, rc: int):
        if rc != 0:
            raise FpbError(self._lib.fpb_match_last_error(self._h).decode())

    def close(self):
        if self._h:
            self._lib.fpb_match_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_templates(self, templates: Sequence) -> None:
        tpl = [_as_template(t) for t in templates]
        counts = np.array([len(t) for t in tpl], np.int32)
        flat = np.ascontiguousarray(np.vstack(tpl)) if len(tpl) else np.zeros((0, 7))
        self._check(self._lib.fpb_match_set_templates(self._h, flat.ctypes.data, counts.ctypes.data, len(tpl)))
        self.n_templates = len(tpl)
        self.counts = counts

    @staticmethod
    def params(dist_thresh=10.0, orient_thresh_deg=12.0, use_type=True, ransac_iter=300, min_inliers=8,
               stop_inlier_ratio=0.25, cross_check=True) -> MatchParams:
        return MatchParams(float(dist_thresh), float(orient_thresh_deg), int(bool(use_type)), int(ransac_iter),
                           int(min_inliers), float(stop_inlier_ratio), int(bool(cross_check)))

    def match(self, pairs, want_matches: bool = True, **kw):
        """pairs: [n, 2] template indices.  Returns (results structured array, matches [n, M, 2] int32, scores [n, M])."""
        pairs = np.ascontiguousarray(np.asarray(pairs, np.int32).reshape(-1, 2))
        n = len(pairs)
        res = np.zeros(n, RESULT_DTYPE)
        if n == 0:
            return res, np.zeros((0, self.max_minutiae, 2), np.int32), np.zeros((0, self.max_minutiae))
        p = self.params(**kw)
        mm = np.zeros((n, self.max_minutiae, 2), np.int32) if want_matches else None
        ms = np.zeros((n, self.max_minutiae), np.float64) if want_matches else None
        self._check(self._lib.fpb_match_pairs(self._h, pairs.ctypes.data, n, C.byref(p), res.ctypes.data,
                                              mm.ctypes.data if want_matches else None,
                                              ms.ctypes.data if want_matches else None))
        return res, mm, ms

    # device-resident form (bench.py): upload once, launch many times
    def upload_pairs(self, pairs) -> int:
        pairs = np.ascontiguousarray(np.asarray(pairs, np.int32).reshape(-1, 2))
        self._check(self._lib.fpb_match_upload_pairs(self._h, pairs.ctypes.data, len(pairs)))
        self._n_pairs = len(pairs)
        return len(pairs)

    def run_device(self, **kw) -> None:
        p = self.params(**kw)
        self._check(self._lib.fpb_match_run_device(self._h, C.byref(p)))

    def sync(self) -> None:
        self._check(self._lib.fpb_match_sync(self._h))

    def download(self):
        res = np.zeros(self._n_pairs, RESULT_DTYPE)
        self._check(self._lib.fpb_match_download(self._h, res.ctypes.data, None, None))
        return res

    @property
    def stream(self) -> int:
        return int(self._lib.fpb_match_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._lib.fpb_match_launch_count(self._h))


def _result_dict(r, mm, ms) -> Dict:
    k = int(r["n_matches"])
    return {"final_score": float(r["final_score"]), "inlier_ratio": float(r["inlier_ratio"]),
            "matches": [(int(mm[i, 0]), int(mm[i, 1]), float(ms[i])) for i in range(k)],
            "theta": float(r["theta"]), "t": np.array([float(r["tx"]), float(r["ty"])])}


def match_pairs(templates: Sequence, pairs, device: int = 0, **kw) -> List[Dict]:
    """match_minutiae_pair over many (index_a, index_b) pairs of `templates` in one launch; list of result dicts."""
    tpl = [_as_template(t) for t in templates]
    maxm = max([len(t) for t in tpl] + [1])
    if maxm > 256:
        raise NotImplementedError("templates above 256 minutiae (the reference's extractor keeps 60)")
    m = MinutiaeMatcher(len(tpl), maxm, int(kw.get("ransac_iter", 300)), device)
    try:
        m.set_templates(tpl)
        res, mm, ms = m.match(pairs, True, **kw)
    finally:
        m.close()
    return [_result_dict(res[i], mm[i], ms[i]) for i in range(len(res))]


def match_minutiae_pair(mins_a, mins_b, dist_thresh=10.0, orient_thresh_deg=12.0, use_type=True, ransac_iter=300,
                        min_inliers=8, stop_inlier_ratio=0.25, cross_check=True, thread_workers=4, debug=False) -> Dict:
    """match.py:219-275"""
    if mins_a is None or mins_b is None:
        return {"final_score": 0.0, "inlier_ratio": 0.0, "matches": []}
    return match_pairs([mins_a, mins_b], [(0, 1)], dist_thresh=dist_thresh, orient_thresh_deg=orient_thresh_deg,
                       use_type=use_type, ransac_iter=ransac_iter, min_inliers=min_inliers,
                       stop_inlier_ratio=stop_inlier_ratio, cross_check=cross_check)[0]
