// RANSAC rigid minutiae matcher for batches of template pairs (include/fpb200_match.h).
//
// Restates /root/reference/src/matching/match.py in float64, evaluated literally (-fmad=false):
//   * k_match_prep  - per template: descriptor weights (match.py:10-22), their NumPy pairwise sum, the position
//                     spread of the early reject (:85-88) and, for every hypothesis seed 42+i, the weighted picks
//                     `rng.choice(idxsA, p=wA/sum)` (:93) and `rng.choice(same_type_idx, p=...)` (:100) - a pick is
//                     `cdf.searchsorted(rng.random(), 'right')`, it depends on ONE template and the seed only, so it
//                     is computed once per template instead of once per pair.
//   * k_match_pairs - one CTA per pair, one warp per hypothesis (ransac_worker, :75-127): rigid transform from the
//                     picked pair, brute-force nearest neighbour of every moved A point among the B points in shared
//                     memory (what sklearn's KDTree does for <= 80 points: one leaf, first minimum wins), distance /
//                     type / angle gates, weighted score.  Hypotheses are then consumed in seed order (first one that
//                     reaches stop_inlier_ratio wins, else the first best score, :158-166), and warp 0 refines:
//                     closed-form 2-D Kabsch rotation (= the SVD of :182-190), re-match, spread check (:206-215),
//                     cross-check (:256-260) and final score (:262-266).
// The uniforms of `default_rng(42+i)` come from a host restatement of NumPy's SeedSequence + PCG64 (bottom of file).
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <new>
#include <vector>

#include "../../include/fpb200.h"
#include "../../include/fpb200_match.h"

#define MT_MAXM 256
#define MT_MAXITER 4096
#define MT_NOPICK 0xFFFFu
#define MT_THREADS 256
#define MT_CHUNK 32            /* hypotheses evaluated between two checks of the stop condition */
#define MT_PI 3.141592653589793
#define MT_2PI 6.283185307179586

struct MatchK {
    double dist_thresh, orient_thresh, two_sd2, two_so2, stop_ratio, win;
    int use_type, n_iter, min_inliers, cross_check;
};

struct MatchTemplates {           // device arrays, template t at [t*maxM]
    const double *x, *y, *o, *w;
    const uint8_t* ty;
    const int* n;
    const double* stat;           // [t][3] = sum(w), std x, std y
    const uint16_t *pickA, *pickB;   // [t][maxIter], [t][2][maxIter]
    const uint16_t* perm;            // [t][maxM]: minutia indices in ascending x (ties by index)
};

// ---------------------------------------------------------------------------------------------------------------
// NumPy arithmetic restated
// ---------------------------------------------------------------------------------------------------------------
// np.sum of a contiguous float64 vector: DOUBLE_pairwise_sum (numpy/_core/src/umath/loops_utils.h.src)
__device__ __forceinline__ double np_pw_block(const double* a, int n) {       // n <= 128
    if (n < 8) {
        double r = 0.;
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    }
    double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
        r0 += a[i]; r1 += a[i + 1]; r2 += a[i + 2]; r3 += a[i + 3];
        r4 += a[i + 4]; r5 += a[i + 5]; r6 += a[i + 6]; r7 += a[i + 7];
    }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; ++i) res += a[i];
    return res;
}
__device__ __forceinline__ double np_pw_level1(const double* a, int n) {       // n <= 256: at most one more split
    if (n <= 128) return np_pw_block(a, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_pw_block(a, n2) + np_pw_block(a + n2, n - n2);
}
__device__ __noinline__ double np_pairwise_sum(const double* a, int n) {       // n <= MT_MAXM = 256
    if (n <= 128) return np_pw_block(a, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_pw_level1(a, n2) + np_pw_level1(a + n2, n - n2);
}

// utils.py:23-27  (d + pi) % (2 pi) - pi with Python's sign-of-divisor modulo
__device__ __forceinline__ double mt_angle_diff(double a, double b) {
    const double d = a - b;
    double m = fmod(d + MT_PI, MT_2PI);
    if (m != 0.0) { if (m < 0.0) m += MT_2PI; } else m = 0.0;
    return m - MT_PI;
}

// first k with cdf[k] > u   (ndarray.searchsorted(u, side='right'))
__device__ __forceinline__ int mt_upper_bound(const double* cdf, int n, double u) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
    return lo;
}

// p = w / sum(w); cdf = p.cumsum(); cdf /= cdf[-1]     (Generator.choice with p, numpy/random/_generator.pyx)
__device__ void mt_make_cdf(const double* w, int n, double* cdf) {
    const double s = np_pairwise_sum(w, n);
    double acc = 0.0;
    for (int k = 0; k < n; ++k) { const double p = w[k] / s; acc = (k == 0) ? p : acc + p; cdf[k] = acc; }
    const double last = cdf[n - 1];
    for (int k = 0; k < n; ++k) cdf[k] = cdf[k] / last;
}

// ---------------------------------------------------------------------------------------------------------------
// per-template preparation
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_match_prep(const double* __restrict__ raw, const int* __restrict__ off,
                                                    const int* __restrict__ cnt, int maxM, int maxIter,
                                                    const double* __restrict__ u, double* X, double* Y, double* O, double* Wt,
                                                    uint8_t* TY, double* stat, uint16_t* pickA, uint16_t* pickB, uint16_t* perm) {
    __shared__ double w[MT_MAXM], xs[MT_MAXM], ys[MT_MAXM], cdfA[MT_MAXM], cdfT[2][MT_MAXM], wsub[2][MT_MAXM];
    __shared__ uint16_t idxT[2][MT_MAXM];
    __shared__ uint8_t tys[MT_MAXM];
    __shared__ int nT[2];
    const int t = blockIdx.x, n = cnt[t];
    const double* rows = raw + (size_t)off[t] * 7;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const double* m = rows + (size_t)k * 7;
        const int type = (int)m[2];
        const double bonus = type == 1 ? 1.25 : 1.0;
        const double base = (0.5 * m[4] + 0.3 * m[5]) + 0.2 * m[6];
        double v = bonus * base;                                   // np.clip(v, 0.05, 2.0)
        v = v < 0.05 ? 0.05 : v; v = v > 2.0 ? 2.0 : v;
        w[k] = v; xs[k] = m[0]; ys[k] = m[1]; tys[k] = (uint8_t)type;
        const size_t g = (size_t)t * maxM + k;
        X[g] = m[0]; Y[g] = m[1]; O[g] = m[3]; Wt[g] = v; TY[g] = (uint8_t)type;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) {           // rank sort by x for the windowed neighbour search
        const double xk = xs[k];
        int r = 0;
        for (int j = 0; j < n; ++j) r += (xs[j] < xk) || (xs[j] == xk && j < k);
        perm[(size_t)t * maxM + r] = (uint16_t)k;
    }
    if (n == 0) { if (threadIdx.x == 0) { stat[t * 3] = 0.0; stat[t * 3 + 1] = 0.0; stat[t * 3 + 2] = 0.0; } return; }
    if (threadIdx.x == 0) {
        stat[t * 3] = np_pairwise_sum(w, n);
        mt_make_cdf(w, n, cdfA);
    } else if (threadIdx.x == 32 || threadIdx.x == 64) {           // np.where(mins_b[:,2] == type)[0] and its cdf
        const int ty = threadIdx.x == 32 ? 0 : 1;
        int c = 0;
        for (int k = 0; k < n; ++k) if (tys[k] == ty) { idxT[ty][c] = (uint16_t)k; wsub[ty][c] = w[k]; ++c; }
        nT[ty] = c;
        if (c) mt_make_cdf(wsub[ty], c, cdfT[ty]);
    } else if (threadIdx.x == 96) {                                // ndarray.std(0): sequential axis-0 sums
        double sx = 0.0, sy = 0.0;
        for (int k = 0; k < n; ++k) { sx = k ? sx + xs[k] : xs[k]; sy = k ? sy + ys[k] : ys[k]; }
        const double mx = sx / n, my = sy / n;
        double vx = 0.0, vy = 0.0;
        for (int k = 0; k < n; ++k) {
            const double dx = xs[k] - mx, dy = ys[k] - my;
            vx = k ? vx + dx * dx : dx * dx; vy = k ? vy + dy * dy : dy * dy;
        }
        stat[t * 3 + 1] = sqrt(vx / n); stat[t * 3 + 2] = sqrt(vy / n);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < maxIter; i += blockDim.x) {
        const double u1 = u[2 * i], u2 = u[2 * i + 1];
        int a = mt_upper_bound(cdfA, n, u1);
        if (a >= n) a = n - 1;
        pickA[(size_t)t * maxIter + i] = (uint16_t)a;
        for (int ty = 0; ty < 2; ++ty) {
            uint16_t b = MT_NOPICK;
            if (nT[ty]) { int k = mt_upper_bound(cdfT[ty], nT[ty], u2); if (k >= nT[ty]) k = nT[ty] - 1; b = idxT[ty][k]; }
            pickB[((size_t)t * 2 + ty) * maxIter + i] = b;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// per-pair matching
// ---------------------------------------------------------------------------------------------------------------
struct PairSmem {
    double *xA, *yA, *oA, *wA, *xB, *yB, *oB, *wB;
    double2* sB;                 // B positions in ascending x
    uint16_t* sA;                // A indices in ascending x (order of the counting pass)
    uint16_t* sidx;              // their original indices
    uint8_t *tA, *tB;
    int nA, nB;
};

// One moved A point against B: nearest neighbour + the three gates of match.py:53-62.  Returns "is an inlier".
// KDTree.query(k=1) is only USED when the neighbour is within dist_thresh (match.py:54): a point farther than
// dist_thresh in x alone can be neither that neighbour nor closer than it, so only the x-window (slightly widened:
// extra candidates are harmless) of the x-sorted B points is scanned.  Ties go to the smallest original index, as in
// sklearn's single-leaf scan.
__device__ __forceinline__ bool mt_gate(const PairSmem& s, const MatchK& k, int ia, double theta, double c, double sn,
                                        double tx, double ty, int& bj, double& d, double& ang) {
    const double ax = s.xA[ia], ay = s.yA[ia];
    const double px = (ax * c + ay * (-sn)) + tx, py = (ax * sn + ay * c) + ty;
    double best = INFINITY;
    bj = 0x7fffffff;
    const double xlo = px - k.win, xhi = px + k.win;
    int lo = 0, hi = s.nB;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (s.sB[mid].x < xlo) lo = mid + 1; else hi = mid; }
    for (int j = lo; j < s.nB; ++j) {
        const double2 q = s.sB[j];
        if (q.x > xhi) break;
        const double dx = px - q.x, dy = py - q.y;
        const double rd = dx * dx + dy * dy;
        if (rd <= best) {
            const int id = s.sidx[j];
            if (rd < best || id < bj) { best = rd; bj = id; }
        }
    }
    if (bj == 0x7fffffff) { bj = 0; return false; }
    d = sqrt(best);
    if (d > k.dist_thresh) return false;
    if (k.use_type && s.tA[ia] != s.tB[bj]) return false;
    ang = fabs(mt_angle_diff(s.oA[ia] + theta, s.oB[bj]));
    return !(ang > k.orient_thresh);
}

// match_with_transform (match.py:33-72) by one warp.  Every lane returns the same n and weighted sum.
// sc: per-warp scratch [maxM].  When RECORD, the inlier list goes to rec_* in ia order.
template <bool RECORD>
__device__ __forceinline__ int mt_eval(const PairSmem& s, const MatchK& k, double theta, double c, double sn, double tx,
                                       double ty, double* sc, double& weighted, uint16_t* rec_ia, uint16_t* rec_ib,
                                       double* rec_s) {
    const int lane = threadIdx.x & 31;
    int n = 0;
    weighted = 0.0;
    for (int base = 0; base < s.nA; base += 32) {
        const int ia = base + lane;
        bool ok = false;
        double score = 0.0;
        int bj = 0;
        if (ia < s.nA) {
            double d, ang;
            ok = mt_gate(s, k, ia, theta, c, sn, tx, ty, bj, d, ang);
            if (ok) {
                const double spatial = exp(-(d * d) / k.two_sd2);
                const double of = exp(-(ang * ang) / k.two_so2);
                score = spatial * of * s.wA[ia] * s.wB[bj];
            }
            sc[ia] = score;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        __syncwarp();
        for (unsigned m = mask; m; m &= m - 1) weighted += sc[base + __ffs(m) - 1];     // sum() in list order
        if (RECORD && ok) {
            const int pos = n + __popc(mask & ((1u << lane) - 1u));
            rec_ia[pos] = (uint16_t)ia; rec_ib[pos] = (uint16_t)bj; rec_s[pos] = score;
        }
        n += __popc(mask);
        __syncwarp();
    }
    return n;
}

__global__ void __launch_bounds__(MT_THREADS) k_match_pairs(MatchTemplates T, const int2* __restrict__ pairs, int nPairs,
                                                            int maxM, int maxIter, MatchK k, fpb_match_result* __restrict__ res,
                                                            int2* __restrict__ out_m, double* __restrict__ out_s) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // carve-up (all 8-byte arrays first)
    double* p = reinterpret_cast<double*>(smem_raw);
    PairSmem s;
    s.xA = p; p += maxM; s.yA = p; p += maxM; s.oA = p; p += maxM; s.wA = p; p += maxM;
    s.xB = p; p += maxM; s.yB = p; p += maxM; s.oB = p; p += maxM; s.wB = p; p += maxM;
    double* sc_all = p; p += (MT_THREADS / 32) * maxM;
    double* rec_s = p; p += maxM;
    double* tax = p; p += maxM;
    double* tay = p; p += maxM;
    s.sB = reinterpret_cast<double2*>(p); p += 2 * maxM;
    double* ch = p; p += 5 * MT_CHUNK;       // per-chunk transforms: theta, cos, sin, tx, ty
    double* h_score = p; p += k.n_iter;
    uint16_t* q = reinterpret_cast<uint16_t*>(p);
    uint16_t* h_n = q; q += (k.n_iter + 3) & ~3;
    uint16_t* rec_ia = q; q += maxM;
    uint16_t* rec_ib = q; q += maxM;
    uint16_t* back = q; q += maxM;
    s.sidx = q; q += maxM;
    s.sA = q; q += maxM;
    s.tA = reinterpret_cast<uint8_t*>(q); s.tB = s.tA + maxM;
    __shared__ int s_best;
    __shared__ int ch_cnt[MT_CHUNK];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = MT_THREADS / 32;
    double* sc = sc_all + warp * maxM;

    for (int pr = blockIdx.x; pr < nPairs; pr += gridDim.x) {
        const int a = pairs[pr].x, b = pairs[pr].y;
        const int nA = T.n[a], nB = T.n[b];
        s.nA = nA; s.nB = nB;
        __syncthreads();                                    // previous pair fully consumed
        for (int i = threadIdx.x; i < nA; i += MT_THREADS) {
            const size_t g = (size_t)a * maxM + i;
            s.xA[i] = T.x[g]; s.yA[i] = T.y[g]; s.oA[i] = T.o[g]; s.wA[i] = T.w[g]; s.tA[i] = T.ty[g];
            s.sA[i] = T.perm[g];
        }
        for (int i = threadIdx.x; i < nB; i += MT_THREADS) {
            const size_t g = (size_t)b * maxM + i;
            s.xB[i] = T.x[g]; s.yB[i] = T.y[g]; s.oB[i] = T.o[g]; s.wB[i] = T.w[g]; s.tB[i] = T.ty[g];
            const int src = T.perm[g];
            s.sB[i] = make_double2(T.x[(size_t)b * maxM + src], T.y[(size_t)b * maxM + src]); s.sidx[i] = (uint16_t)src;
        }
        __syncthreads();
        const double sumA = T.stat[a * 3], sumB = T.stat[b * 3];
        const double possible = sumA < sumB ? sumA : sumB;         // min(np.sum(wA), np.sum(wB))
        const int minN = nA < nB ? nA : nB;
        bool any = nA >= 8 && nB >= 8;                              // match.py:81-82 (and :135-136 for empty templates)
        if (any) {                                                  // match.py:85-88
            const double ex = T.stat[a * 3 + 1] - T.stat[b * 3 + 1], ey = T.stat[a * 3 + 2] - T.stat[b * 3 + 2];
            if (sqrt(ex * ex + ey * ey) > 35.0) any = false;
        }
        int best = -1;
        if (any) {
            const uint16_t* pkA = T.pickA + (size_t)a * maxIter;
            const uint16_t* pkB = T.pickB + (size_t)b * 2 * maxIter;
            int n_done = k.n_iter;
            const double stopN0 = k.stop_ratio * (double)minN;
            for (int i0 = 0; i0 < k.n_iter; i0 += MT_CHUNK) {
                const int nh = min(MT_CHUNK, k.n_iter - i0);
                // (a) the rigid transform of every hypothesis of the chunk (estimate_transform_rigid_by_pair)
                if (threadIdx.x < nh) {
                    const int i = i0 + threadIdx.x;
                    const int pA = pkA[i];
                    const unsigned pB = pkB[(size_t)s.tA[pA] * maxIter + i];
                    double theta = NAN, c = 0, sn = 0, tx = 0, ty = 0;
                    if (pB != MT_NOPICK) {
                        theta = mt_angle_diff(s.oB[pB], s.oA[pA]);
                        c = cos(theta); sn = sin(theta);
                        const double rx = s.xA[pA] * c + s.yA[pA] * (-sn), ry = s.xA[pA] * sn + s.yA[pA] * c;
                        tx = s.xB[pB] - rx; ty = s.yB[pB] - ry;
                    }
                    double* h = ch + 5 * threadIdx.x;
                    h[0] = theta; h[1] = c; h[2] = sn; h[3] = tx; h[4] = ty;
                    ch_cnt[threadIdx.x] = 0;
                }
                __syncthreads();
                // (b) inlier COUNT of every hypothesis, all threads over the flat (hypothesis, A point) items; A in
                //     ascending x so that neighbouring lanes scan neighbouring windows of B
                for (int it = threadIdx.x; it < nh * nA; it += MT_THREADS) {
                    const int hi = it / nA, ia = s.sA[it - hi * nA];
                    const double* h = ch + 5 * hi;
                    if (h[0] == h[0]) {
                        int bj; double d, ang;
                        if (mt_gate(s, k, ia, h[0], h[1], h[2], h[3], h[4], bj, d, ang)) atomicAdd(&ch_cnt[hi], 1);
                    }
                }
                __syncthreads();
                // (c) only hypotheses with enough inliers are scored (ordered sum, one warp each)
                bool stop_here = false;
                for (int hi = warp; hi < nh; hi += nwarps) {
                    const int i = i0 + hi;
                    int n = ch_cnt[hi];
                    double score = 0.0;
                    if (n < k.min_inliers) n = 0;
                    else {
                        const double* h = ch + 5 * hi;
                        double weighted;
                        mt_eval<false>(s, k, h[0], h[1], h[2], h[3], h[4], sc, weighted, nullptr, nullptr, nullptr);
                        score = pow(weighted / (possible + 1e-6), 0.75);
                        score = score < 0.0 ? 0.0 : (score > 1.0 ? 1.0 : score);
                    }
                    if (lane == 0) { h_score[i] = score; h_n[i] = (uint16_t)n; }
                    stop_here |= (double)n >= stopN0;
                }
                // match.py:164-166: the loop over hypotheses ends at the first one that reaches the stop ratio
                if (__syncthreads_or(stop_here)) { n_done = i0 + nh; break; }
            }
            __syncthreads();
            if (warp == 0) {                                        // match.py:158-166 in seed order
                const double stopN = k.stop_ratio * (double)minN;
                int istop = 0x7fffffff, ibest = 0x7fffffff;
                double sbest = 0.0;
                for (int i = lane; i < n_done; i += 32) {
                    if ((double)h_n[i] >= stopN && i < istop) istop = i;
                    if (h_score[i] > sbest) { sbest = h_score[i]; ibest = i; }
                }
                for (int o = 16; o; o >>= 1) {
                    const int is2 = __shfl_xor_sync(0xffffffffu, istop, o);
                    const double sb2 = __shfl_xor_sync(0xffffffffu, sbest, o);
                    const int ib2 = __shfl_xor_sync(0xffffffffu, ibest, o);
                    istop = is2 < istop ? is2 : istop;
                    if (sb2 > sbest || (sb2 == sbest && ib2 < ibest)) { sbest = sb2; ibest = ib2; }
                }
                // before the stop index the running best is whatever scored highest so far, but the stopping
                // hypothesis overwrites it; a stop can only come from a hypothesis with inliers, i.e. score > 0
                int bi = -1;
                if (istop != 0x7fffffff) bi = (h_score[istop] > 0.0) ? istop : -1;
                else if (sbest > 0.0) bi = ibest;
                if (lane == 0) s_best = bi;
            }
            __syncthreads();
            best = s_best;
        }
        if (warp != 0) continue;
        // ---- warp 0: refinement, cross-check, final score ----------------------------------------------------
        fpb_match_result r;
        r.final_score = 0.0; r.inlier_ratio = 0.0; r.theta = 0.0; r.tx = 0.0; r.ty = 0.0; r.n_matches = 0; r.best_iter = best;
        int n_final = 0;
        if (best >= 0) {
            const int pA = T.pickA[(size_t)a * maxIter + best];
            const unsigned pB = T.pickB[((size_t)b * 2 + s.tA[pA]) * maxIter + best];
            double theta = mt_angle_diff(s.oB[pB], s.oA[pA]);
            double c = cos(theta), sn = sin(theta);
            double tx = s.xB[pB] - (s.xA[pA] * c + s.yA[pA] * (-sn)), ty = s.yB[pB] - (s.xA[pA] * sn + s.yA[pA] * c);
            double weighted;
            int n = mt_eval<true>(s, k, theta, c, sn, tx, ty, sc, weighted, rec_ia, rec_ib, rec_s);
            __syncwarp();
            // centroids and cross-covariance of the inlier pairs (match.py:176-183), every lane redundantly
            double cax = 0, cay = 0, cbx = 0, cby = 0;
            for (int i = 0; i < n; ++i) {
                const int ia = rec_ia[i], ib = rec_ib[i];
                cax = i ? cax + s.xA[ia] : s.xA[ia]; cay = i ? cay + s.yA[ia] : s.yA[ia];
                cbx = i ? cbx + s.xB[ib] : s.xB[ib]; cby = i ? cby + s.yB[ib] : s.yB[ib];
            }
            cax = cax / n; cay = cay / n; cbx = cbx / n; cby = cby / n;
            double h00 = 0, h01 = 0, h10 = 0, h11 = 0;
            for (int i = 0; i < n; ++i) {
                const int ia = rec_ia[i], ib = rec_ib[i];
                const double ax = s.xA[ia] - cax, ay = s.yA[ia] - cay, bx = s.xB[ib] - cbx, by = s.yB[ib] - cby;
                h00 += ax * bx; h01 += ax * by; h10 += ay * bx; h11 += ay * by;
            }
            // R = V U^T of H = U S V^T with the det fix (match.py:184-190) is the proper rotation maximising
            // trace(R H); in 2-D that is theta = atan2(H01 - H10, H00 + H11)
            theta = atan2(h01 - h10, h00 + h11);
            c = cos(theta); sn = sin(theta);
            tx = cbx - (cax * c + cay * (-sn)); ty = cby - (cax * sn + cay * c);
            __syncwarp();
            n = mt_eval<true>(s, k, theta, c, sn, tx, ty, sc, weighted, rec_ia, rec_ib, rec_s);
            __syncwarp();
            bool rejected = false;
            if (n >= 8) {                                           // match.py:206-215
                double max_ = 0, may = 0, mbx = 0, mby = 0;
                for (int i = 0; i < n; ++i) {
                    const int ia = rec_ia[i], ib = rec_ib[i];
                    max_ = i ? max_ + s.xA[ia] : s.xA[ia]; may = i ? may + s.yA[ia] : s.yA[ia];
                    mbx = i ? mbx + s.xB[ib] : s.xB[ib]; mby = i ? mby + s.yB[ib] : s.yB[ib];
                }
                max_ = max_ / n; may = may / n; mbx = mbx / n; mby = mby / n;
                for (int i = lane; i < n; i += 32) {
                    const int ia = rec_ia[i], ib = rec_ib[i];
                    const double ax = s.xA[ia] - max_, ay = s.yA[ia] - may, bx = s.xB[ib] - mbx, by = s.yB[ib] - mby;
                    tax[i] = sqrt(ax * ax + ay * ay); tay[i] = sqrt(bx * bx + by * by);
                }
                __syncwarp();
                const double dA = np_pairwise_sum(tax, n) / n, dB = np_pairwise_sum(tay, n) / n;
                __syncwarp();
                if (fabs(dA - dB) > 18.0) rejected = true;
            }
            if (!rejected) {
                r.theta = theta; r.tx = tx; r.ty = ty;
                n_final = n;
                if (k.cross_check && n > 0) {                       // match.py:256-260
                    const double nsn = -sn;
                    for (int i = lane; i < nA; i += 32) {
                        tax[i] = (s.xA[i] * c + s.yA[i] * nsn) + tx; tay[i] = (s.xA[i] * sn + s.yA[i] * c) + ty;
                    }
                    __syncwarp();
                    for (int j = lane; j < nB; j += 32) {
                        const double bx = s.xB[j], by = s.yB[j];
                        double bestd = INFINITY; int bi = 0;
                        for (int i = 0; i < nA; ++i) {
                            const double dx = bx - tax[i], dy = by - tay[i];
                            const double rd = dx * dx + dy * dy;
                            if (rd < bestd) { bestd = rd; bi = i; }
                        }
                        back[j] = (uint16_t)bi;
                    }
                    __syncwarp();
                    int kept = 0;
                    for (int base = 0; base < n; base += 32) {      // ordered compaction in place (kept <= base)
                        const int i = base + lane;
                        int ia = 0, ib = 0; double sv = 0.0; bool keep = false;
                        if (i < n) { ia = rec_ia[i]; ib = rec_ib[i]; sv = rec_s[i]; keep = back[ib] == ia; }
                        const unsigned mask = __ballot_sync(0xffffffffu, keep);
                        __syncwarp();
                        if (keep) {
                            const int pos = kept + __popc(mask & ((1u << lane) - 1u));
                            rec_ia[pos] = (uint16_t)ia; rec_ib[pos] = (uint16_t)ib; rec_s[pos] = sv;
                        }
                        kept += __popc(mask);
                        __syncwarp();
                    }
                    n_final = kept;
                }
            }
        }
        // match.py:262-268
        double wsum = 0.0;
        for (int i = 0; i < n_final; ++i) wsum = i ? wsum + rec_s[i] : rec_s[i];
        double fs = pow(wsum / (possible + 1e-6), 0.25);
        if (!(nA > 0 && nB > 0)) fs = 0.0;                          // 0 / 1e-6
        fs = fs < 0.0 ? 0.0 : (fs > 1.0 ? 1.0 : fs);
        r.final_score = fs;
        r.inlier_ratio = (double)n_final / (double)(minN > 1 ? minN : 1);
        r.n_matches = n_final;
        if (lane == 0) res[pr] = r;
        if (out_m)
            for (int i = lane; i < n_final; i += 32) {
                out_m[(size_t)pr * maxM + i] = make_int2(rec_ia[i], rec_ib[i]);
                out_s[(size_t)pr * maxM + i] = rec_s[i];
            }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host: handle + C ABI
// ---------------------------------------------------------------------------------------------------------------
static char g_match_error[512] = "";

struct fpb_matcher {
    int device, maxT, maxM, maxIter, nT, nPairs, capPairs, sms;
    cudaStream_t st;
    long long launches;
    char err[512];
    double *d_u, *d_raw, *d_x, *d_y, *d_o, *d_w, *d_stat, *d_ms;
    uint8_t* d_ty;
    int *d_off, *d_cnt;
    uint16_t *d_pickA, *d_pickB, *d_perm;
    int2 *d_pairs, *d_m;
    fpb_match_result* d_res;
};

static int mfail(fpb_matcher* m, int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(m ? m->err : g_match_error, 512, fmt, ap);
    va_end(ap);
    return code;
}
#define MCU(m, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return mfail(m, FPB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

extern "C" const char* fpb_match_last_error(const fpb_matcher* m) { return m ? m->err : g_match_error; }
extern "C" long long fpb_match_launch_count(const fpb_matcher* m) { return m ? m->launches : 0; }
extern "C" void* fpb_match_stream(fpb_matcher* m) { return m ? (void*)m->st : nullptr; }

extern "C" void fpb_match_destroy(fpb_matcher* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->st) cudaStreamSynchronize(m->st);
    void* dev[] = {m->d_u, m->d_raw, m->d_x, m->d_y, m->d_o, m->d_w, m->d_stat, m->d_ms, m->d_ty, m->d_off, m->d_cnt,
                   m->d_pickA, m->d_pickB, m->d_perm, m->d_pairs, m->d_m, m->d_res};
    for (void* p : dev) if (p) cudaFree(p);
    if (m->st) cudaStreamDestroy(m->st);
    delete m;
}

extern "C" int fpb_match_create(fpb_matcher** out, int device, int max_templates, int max_minutiae, int max_iter) {
    if (!out) return mfail(nullptr, FPB_E_ARG, "fpb_match_create: out is NULL");
    *out = nullptr;
    if (max_templates < 1 || max_minutiae < 1 || max_minutiae > MT_MAXM || max_iter < 1 || max_iter > MT_MAXITER)
        return mfail(nullptr, FPB_E_ARG, "fpb_match_create: need 1 <= max_minutiae <= %d and 1 <= max_iter <= %d", MT_MAXM, MT_MAXITER);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return mfail(nullptr, FPB_E_CUDA, "fpb_match_create: no CUDA device (%s) - this library has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return mfail(nullptr, FPB_E_ARG, "fpb_match_create: device %d out of range (%d devices)", device, ndev);
    fpb_matcher* m = new (std::nothrow) fpb_matcher();
    if (!m) return mfail(nullptr, FPB_E_NOMEM, "fpb_match_create: out of host memory");
    memset(m, 0, sizeof(*m));
    m->device = device; m->maxT = max_templates; m->maxM = (max_minutiae + 3) & ~3; m->maxIter = max_iter;
#define MCUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
        mfail(nullptr, FPB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); fpb_match_destroy(m); return FPB_E_CUDA; } } while (0)
    MCUC(cudaSetDevice(device));
    MCUC(cudaDeviceGetAttribute(&m->sms, cudaDevAttrMultiProcessorCount, device));
    MCUC(cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking));
    const size_t TM = (size_t)m->maxT * m->maxM;
    MCUC(cudaMalloc(&m->d_u, sizeof(double) * 2 * max_iter));
    MCUC(cudaMalloc(&m->d_raw, sizeof(double) * 7 * TM));
    MCUC(cudaMalloc(&m->d_x, sizeof(double) * TM));
    MCUC(cudaMalloc(&m->d_y, sizeof(double) * TM));
    MCUC(cudaMalloc(&m->d_o, sizeof(double) * TM));
    MCUC(cudaMalloc(&m->d_w, sizeof(double) * TM));
    MCUC(cudaMalloc(&m->d_ty, TM));
    MCUC(cudaMalloc(&m->d_stat, sizeof(double) * 3 * m->maxT));
    MCUC(cudaMalloc(&m->d_off, sizeof(int) * m->maxT));
    MCUC(cudaMalloc(&m->d_cnt, sizeof(int) * m->maxT));
    MCUC(cudaMalloc(&m->d_pickA, sizeof(uint16_t) * (size_t)m->maxT * max_iter));
    MCUC(cudaMalloc(&m->d_pickB, sizeof(uint16_t) * (size_t)m->maxT * 2 * max_iter));
    MCUC(cudaMalloc(&m->d_perm, sizeof(uint16_t) * TM));
    std::vector<double> u(2 * (size_t)max_iter);
    fpb_match_seed_uniforms(42, max_iter, u.data());                // base_seed = 42, match.py:144
    MCUC(cudaMemcpy(m->d_u, u.data(), sizeof(double) * u.size(), cudaMemcpyHostToDevice));
    *out = m;
    return FPB_OK;
}

extern "C" int fpb_match_set_templates(fpb_matcher* m, const double* mins, const int32_t* counts, int n) {
    if (!m) return FPB_E_ARG;
    if (!counts || n < 1 || n > m->maxT) return mfail(m, FPB_E_ARG, "fpb_match_set_templates: n=%d outside 1..%d", n, m->maxT);
    std::vector<int> off(n);
    long long tot = 0;
    for (int i = 0; i < n; ++i) {
        if (counts[i] < 0 || counts[i] > m->maxM) return mfail(m, FPB_E_SHAPE, "template %d has %d minutiae (max %d)", i, counts[i], m->maxM);
        off[i] = (int)tot; tot += counts[i];
    }
    if (tot > 0 && !mins) return mfail(m, FPB_E_ARG, "fpb_match_set_templates: mins is NULL");
    for (long long r = 0; r < tot; ++r) {
        const double t = mins[r * 7 + 2];
        if (!(t == 0.0 || t == 1.0)) return mfail(m, FPB_E_ARG, "fpb_match_set_templates: type column must be 0 or 1 (row %lld has %g)", r, t);
    }
    MCU(m, cudaSetDevice(m->device));
    if (tot) MCU(m, cudaMemcpyAsync(m->d_raw, mins, sizeof(double) * 7 * tot, cudaMemcpyHostToDevice, m->st));
    MCU(m, cudaMemcpyAsync(m->d_off, off.data(), sizeof(int) * n, cudaMemcpyHostToDevice, m->st));
    MCU(m, cudaMemcpyAsync(m->d_cnt, counts, sizeof(int) * n, cudaMemcpyHostToDevice, m->st));
    k_match_prep<<<n, 128, 0, m->st>>>(m->d_raw, m->d_off, m->d_cnt, m->maxM, m->maxIter, m->d_u, m->d_x, m->d_y, m->d_o,
                                       m->d_w, m->d_ty, m->d_stat, m->d_pickA, m->d_pickB, m->d_perm);
    m->launches++;
    MCU(m, cudaGetLastError());
    MCU(m, cudaStreamSynchronize(m->st));                           // `off` and the caller's buffers may go away
    m->nT = n;
    return FPB_OK;
}

static int ensure_pairs(fpb_matcher* m, int n_pairs) {
    if (n_pairs <= m->capPairs) return FPB_OK;
    MCU(m, cudaStreamSynchronize(m->st));
    if (m->d_pairs) cudaFree(m->d_pairs);
    if (m->d_res) cudaFree(m->d_res);
    if (m->d_m) cudaFree(m->d_m);
    if (m->d_ms) cudaFree(m->d_ms);
    m->d_pairs = nullptr; m->d_res = nullptr; m->d_m = nullptr; m->d_ms = nullptr; m->capPairs = 0;
    MCU(m, cudaMalloc(&m->d_pairs, sizeof(int2) * (size_t)n_pairs));
    MCU(m, cudaMalloc(&m->d_res, sizeof(fpb_match_result) * (size_t)n_pairs));
    MCU(m, cudaMalloc(&m->d_m, sizeof(int2) * (size_t)n_pairs * m->maxM));
    MCU(m, cudaMalloc(&m->d_ms, sizeof(double) * (size_t)n_pairs * m->maxM));
    m->capPairs = n_pairs;
    return FPB_OK;
}

extern "C" int fpb_match_upload_pairs(fpb_matcher* m, const int32_t* pairs, int n_pairs) {
    if (!m) return FPB_E_ARG;
    if (!pairs || n_pairs < 1) return mfail(m, FPB_E_ARG, "fpb_match_upload_pairs: no pairs");
    if (m->nT == 0) return mfail(m, FPB_E_STATE, "fpb_match_upload_pairs: call fpb_match_set_templates first");
    for (int i = 0; i < 2 * n_pairs; ++i)
        if (pairs[i] < 0 || pairs[i] >= m->nT) return mfail(m, FPB_E_ARG, "pair %d references template %d (have %d)", i / 2, pairs[i], m->nT);
    MCU(m, cudaSetDevice(m->device));
    const int rc = ensure_pairs(m, n_pairs);
    if (rc) return rc;
    MCU(m, cudaMemcpyAsync(m->d_pairs, pairs, sizeof(int2) * (size_t)n_pairs, cudaMemcpyHostToDevice, m->st));
    MCU(m, cudaStreamSynchronize(m->st));
    m->nPairs = n_pairs;
    return FPB_OK;
}

extern "C" int fpb_match_run_device(fpb_matcher* m, const fpb_match_params* p) {
    if (!m) return FPB_E_ARG;
    if (m->nPairs < 1) return mfail(m, FPB_E_STATE, "fpb_match_run_device: no pairs uploaded");
    fpb_match_params d = {10.0, 12.0, 1, 300, 8, 0.25, 1};
    if (p) d = *p;
    if (d.ransac_iter < 1 || d.ransac_iter > m->maxIter)
        return mfail(m, FPB_E_ARG, "ransac_iter %d outside 1..%d (max_iter of this matcher)", d.ransac_iter, m->maxIter);
    MatchK k;
    k.dist_thresh = d.dist_thresh;
    k.orient_thresh = d.orient_thresh_deg * (MT_PI / 180.0);      // math.radians
    const double sd = d.dist_thresh * 0.7, so = k.orient_thresh * 0.7;
    k.two_sd2 = 2 * pow(sd, 2.0); k.two_so2 = 2 * pow(so, 2.0);   // 2 * sigma**2 as Python evaluates it (libm pow)
    k.win = d.dist_thresh * 1.000001 + 1e-9;
    k.stop_ratio = d.stop_inlier_ratio; k.use_type = d.use_type; k.n_iter = d.ransac_iter; k.min_inliers = d.min_inliers;
    k.cross_check = d.cross_check;
    MCU(m, cudaSetDevice(m->device));
    const int M = m->maxM;
    const size_t smem = sizeof(double) * ((size_t)10 * M + (MT_THREADS / 32) * M + 3 * M + 5 * MT_CHUNK + k.n_iter) +
                        sizeof(uint16_t) * (((k.n_iter + 3) & ~3) + 5 * (size_t)M) + 2 * (size_t)M + 16;
    MCU(m, cudaFuncSetAttribute(k_match_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    MCU(m, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_match_pairs, MT_THREADS, smem));
    if (per_sm < 1) return mfail(m, FPB_E_SHAPE, "k_match_pairs does not fit: %zu bytes of shared memory", smem);
    long long grid = (long long)m->sms * per_sm;                   // one wave of resident CTAs, pairs strided over them
    if (grid > m->nPairs) grid = m->nPairs;
    MatchTemplates T = {m->d_x, m->d_y, m->d_o, m->d_w, m->d_ty, m->d_cnt, m->d_stat, m->d_pickA, m->d_pickB, m->d_perm};
    k_match_pairs<<<(unsigned)grid, MT_THREADS, smem, m->st>>>(T, m->d_pairs, m->nPairs, M, m->maxIter, k, m->d_res, m->d_m, m->d_ms);
    m->launches++;
    MCU(m, cudaGetLastError());
    return FPB_OK;
}

extern "C" int fpb_match_sync(fpb_matcher* m) {
    if (!m) return FPB_E_ARG;
    MCU(m, cudaSetDevice(m->device));
    MCU(m, cudaStreamSynchronize(m->st));
    return FPB_OK;
}

extern "C" int fpb_match_download(fpb_matcher* m, fpb_match_result* results, int32_t* matches, double* match_scores) {
    if (!m) return FPB_E_ARG;
    if (m->nPairs < 1 || !results) return mfail(m, FPB_E_STATE, "fpb_match_download: nothing to download");
    MCU(m, cudaSetDevice(m->device));
    const size_t np = (size_t)m->nPairs;
    MCU(m, cudaMemcpyAsync(results, m->d_res, sizeof(fpb_match_result) * np, cudaMemcpyDeviceToHost, m->st));
    if (matches) MCU(m, cudaMemcpyAsync(matches, m->d_m, sizeof(int2) * np * m->maxM, cudaMemcpyDeviceToHost, m->st));
    if (match_scores) MCU(m, cudaMemcpyAsync(match_scores, m->d_ms, sizeof(double) * np * m->maxM, cudaMemcpyDeviceToHost, m->st));
    MCU(m, cudaStreamSynchronize(m->st));
    return FPB_OK;
}

extern "C" int fpb_match_pairs(fpb_matcher* m, const int32_t* pairs, int n_pairs, const fpb_match_params* p,
                               fpb_match_result* results, int32_t* matches, double* match_scores) {
    int rc = fpb_match_upload_pairs(m, pairs, n_pairs);
    if (rc) return rc;
    rc = fpb_match_run_device(m, p);
    if (rc) return rc;
    return fpb_match_download(m, results, matches, match_scores);
}

// ---------------------------------------------------------------------------------------------------------------
// numpy.random.default_rng(seed).random() restated: SeedSequence (numpy/random/bit_generator.pyx) + PCG64
// (XSL-RR 128/64, numpy/random/src/pcg64) + next_double = (u64 >> 11) / 2^53
// ---------------------------------------------------------------------------------------------------------------
namespace {
typedef unsigned __int128 u128;

struct SeedSeq {
    uint32_t pool[4];
    static uint32_t hashmix(uint32_t v, uint32_t& hc) {
        v ^= hc; hc *= 0x931e8875u; v *= hc; v ^= v >> 16; return v;
    }
    static uint32_t mix(uint32_t x, uint32_t y) {
        uint32_t r = 0xca01f9ddu * x - 0x4973f715u * y; r ^= r >> 16; return r;
    }
    explicit SeedSeq(uint64_t seed) {
        uint32_t ent[2]; int ne = 1;
        ent[0] = (uint32_t)seed; ent[1] = (uint32_t)(seed >> 32);
        if (ent[1]) ne = 2;
        uint32_t hc = 0x43b0d7e5u;
        for (int i = 0; i < 4; ++i) pool[i] = hashmix(i < ne ? ent[i] : 0u, hc);
        for (int s = 0; s < 4; ++s)
            for (int d = 0; d < 4; ++d)
                if (s != d) pool[d] = mix(pool[d], hashmix(pool[s], hc));
    }
    void generate(uint32_t* out, int n) const {
        uint32_t hc = 0x8b51f9ddu;
        for (int i = 0; i < n; ++i) {
            uint32_t v = pool[i % 4];
            v ^= hc; hc *= 0x58f38dedu; v *= hc; v ^= v >> 16;
            out[i] = v;
        }
    }
};

struct Pcg64 {
    u128 state, inc;
    static u128 mult() { return ((u128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull; }
    void step() { state = state * mult() + inc; }
    explicit Pcg64(uint64_t seed) {
        uint32_t w[8];
        SeedSeq(seed).generate(w, 8);
        uint64_t s[4];
        for (int i = 0; i < 4; ++i) s[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
        const u128 initstate = ((u128)s[0] << 64) | s[1], initseq = ((u128)s[2] << 64) | s[3];
        state = 0; inc = (initseq << 1) | 1;
        step(); state += initstate; step();
    }
    uint64_t next() {
        step();
        const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state, x = hi ^ lo;
        const unsigned rot = (unsigned)(state >> 122);
        return (x >> rot) | (x << ((64 - rot) & 63));
    }
    double next_double() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};
}  // namespace

extern "C" int fpb_match_seed_uniforms(uint64_t seed0, int n_iter, double* out) {
    if (!out || n_iter < 0) return FPB_E_ARG;
    for (int i = 0; i < n_iter; ++i) {
        Pcg64 g(seed0 + (uint64_t)i);
        out[2 * i] = g.next_double();
        out[2 * i + 1] = g.next_double();
    }
    return FPB_OK;
}
