// Scalar (one thread per image / per tile) arithmetic of the hot path, written as FPB_HD
// functions so that the exact device code can also be compiled by g++ for the no-GPU
// checks in tests/hostcheck (never a product path).  Build flags: nvcc -fmad=false and
// g++ -ffp-contract=off, so every float expression below is evaluated literally.
#pragma once
#include <stdint.h>
#include <math.h>

#ifndef FPB_HD
#ifdef __CUDACC__
#define FPB_HD __host__ __device__ __forceinline__
#else
#define FPB_HD static inline
#endif
#endif

// ---- np.percentile(img.astype(f32)/255, q) from the cumulative 256-bin histogram ----------
// NumPy 2.3 `_quantile`/`_lerp` with a float32 array and a python-float q
// (fingerprint_preprocess.py:20): everything is float32.
FPB_HD float fpb_kth_u8_unit(const unsigned* cum, int k) {
    int v = 0;
    while (v < 255 && cum[v] < (unsigned)(k + 1)) ++v;
    return (float)v / 255.0f;
}

FPB_HD float fpb_percentile_u8_unit(const unsigned* cum, int n, float q) {
    const float qf = q / 100.0f;
    const float vi = (float)(n - 1) * qf;
    const float lo_f = floorf(vi);
    int lo = (int)lo_f, hi = lo + 1;
    if (vi >= (float)(n - 1)) lo = hi = n - 1;
    const float g = vi - lo_f;
    const float a = fpb_kth_u8_unit(cum, lo), b = fpb_kth_u8_unit(cum, hi);
    const float d = b - a;
    if (g >= 0.5f) return b - d * (1.0f - g);
    return a + d * g;
}

// (clip((v/255 - lo)/(hi - lo + 1e-12), 0, 1) * 255).astype(uint8)   (:20-23)
FPB_HD uint8_t fpb_stretch_value(int v, float lo, float hi) {
    const float span = (hi - lo) + 1e-12f;
    float t = ((float)v / 255.0f - lo) / span;
    t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
    return (uint8_t)(t * 255.0f);
}

// ---- cv2 THRESH_OTSU on uint8 (imgproc/thresh.cpp::getThreshVal_Otsu_8u), double arithmetic ----
FPB_HD int fpb_otsu_u8(const unsigned* hist, int n) {
    const double scale = 1.0 / (double)n;
    double mu = 0.0;
    for (int i = 0; i < 256; ++i) mu += (double)i * (double)hist[i];
    mu *= scale;
    double mu1 = 0.0, q1 = 0.0, best = 0.0;
    int best_t = 0;
    const double eps = 1.1920928955078125e-07;   // FLT_EPSILON
    for (int i = 0; i < 256; ++i) {
        const double p = (double)hist[i] * scale;
        mu1 *= q1;
        q1 += p;
        const double q2 = 1.0 - q1;
        const double mn = q1 < q2 ? q1 : q2, mx = q1 > q2 ? q1 : q2;
        if (mn < eps || mx > 1.0 - eps) continue;
        mu1 = (mu1 + (double)i * p) / q1;
        const double mu2 = (mu - q1 * mu1) / q2;
        const double sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
        if (sigma > best) { best = sigma; best_t = i; }
    }
    return best_t;
}

// ---- skimage.filters.threshold_otsu on a float32 patch that holds integers 0..255 ----------
// (np.histogram with 256 bins over [min,max] in float32, float32 counts / cumsums).
// ih: integer histogram of the patch (256 bins); scratch: counts[256], centers[256], tmp[512].
// Returns the threshold; `sub < t` is then evaluated on the integer-valued pixels.
FPB_HD float fpb_patch_otsu(const unsigned* ih, float* counts, float* centers, float* tmp) {
    int a = 0, b = 255;
    while (a < 255 && ih[a] == 0) ++a;
    while (b > 0 && ih[b] == 0) --b;
    if (a >= b) return (float)a;                       // constant patch: skimage returns the value
    const float fa = (float)a, fb = (float)b;
    const float norm = fb - fa;
    const float step = norm / 256.0f;
    // np.linspace(fa, fb, 257, dtype=f32): edges[i] = i*step + fa, edges[256] = fb
    for (int i = 0; i < 256; ++i) counts[i] = 0.0f;
    for (int v = a; v <= b; ++v) {
        if (ih[v] == 0) continue;
        const float fv = (float)v;
        int i = (int)(((fv - fa) / norm) * 256.0f);   // np.histogram fast path, then edge fix-ups
        if (i == 256) i = 255;
        const float e_i = (float)i * step + fa;
        if (fv < e_i) --i;
        const float e_n = (i + 1 == 256) ? fb : ((float)(i + 1) * step + fa);
        if (fv >= e_n && i != 255) ++i;
        counts[i] += (float)ih[v];
    }
    for (int i = 0; i < 256; ++i) {
        const float e0 = (float)i * step + fa;
        const float e1 = (i + 1 == 256) ? fb : ((float)(i + 1) * step + fa);
        centers[i] = (e0 + e1) / 2.0f;
    }
    // backward cumulative sums (np.cumsum of the reversed arrays): mean2 -> tmp[0..255], weight2 -> tmp[256..511]
    float* mean2 = tmp;
    float* weight2 = tmp + 256;
    float w2 = 0.0f, s2 = 0.0f;
    for (int i = 255; i >= 0; --i) {
        w2 += counts[i];
        s2 += counts[i] * centers[i];
        weight2[i] = w2;
        mean2[i] = s2 / w2;
    }
    // forward pass; variance12[i] = (w1[i]*w2[i+1]) * (m1[i]-m2[i+1])^2, i = 0..254; np.argmax = first maximum.
    // Bins 0 and 255 are always occupied (min and max pixel), so no 0/0 occurs.
    float best = -1.0f; int best_i = 0;
    float w1 = 0.0f, s1 = 0.0f;
    for (int i = 0; i < 255; ++i) {
        w1 += counts[i];
        s1 += counts[i] * centers[i];
        const float m1 = s1 / w1;
        const float dm = m1 - mean2[i + 1];
        const float var = (w1 * weight2[i + 1]) * (dm * dm);
        if (var > best) { best = var; best_i = i; }
    }
    return centers[best_i];
}

#include "gauss_tables.h"
// ---- scipy.ndimage.gaussian_filter weights (host side; passed to the kernels by value) ----------
// _gaussian_kernel1d: radius int(4*sigma+0.5), phi = exp(-0.5/sigma^2 * x^2), phi / phi.sum() in float64;
// the sum follows NumPy's pairwise routine (n < 8: plain loop; n <= 128: eight strided partial sums
// combined as a tree, then the remainder) so that the table is bit-identical to SciPy's.
static inline double fpb_numpy_pairwise_sum(const double* a, int n) {
    if (n < 8) { double s = 0.0; for (int i = 0; i < n; ++i) s += a[i]; return s; }
    double r[8];
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) for (int k = 0; k < 8; ++k) r[k] += a[i + k];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

static inline int fpb_gauss_weights_fill(double sigma, double* w, int cap) {
    for (unsigned t = 0; t < sizeof(FPB_GAUSS_TABLES) / sizeof(FPB_GAUSS_TABLES[0]); ++t)
        if (FPB_GAUSS_TABLES[t].sigma == sigma && 2 * FPB_GAUSS_TABLES[t].r + 1 <= cap) {   // frozen from SciPy
            for (int i = 0; i < 2 * FPB_GAUSS_TABLES[t].r + 1; ++i) w[i] = FPB_GAUSS_TABLES[t].w[i];
            return FPB_GAUSS_TABLES[t].r;
        }
    const int r = (int)(4.0 * sigma + 0.5);
    const int n = 2 * r + 1;
    if (n > cap) return -1;
    const double s2 = sigma * sigma;
    for (int i = 0; i < n; ++i) { const double x = (double)(i - r); w[i] = exp(-0.5 / s2 * (x * x)); }
    const double sum = fpb_numpy_pairwise_sum(w, n);
    for (int i = 0; i < n; ++i) w[i] = w[i] / sum;
    return r;
}


// ---- Zhang & Suen (CACM 1984) deletion test for the 32 pixels of a word at once ----------------------------------------------
// Inputs: the eight neighbour planes of the word (bit j of `N` = the pixel above pixel j, ...).  Returns the mask of pixels the
// sub-iteration `pass` (1 or 2) deletes:  2 <= B <= 6,  A == 1,  pass 1: N*E*S == 0 && E*S*W == 0,  pass 2: N*E*W == 0 && N*S*W == 0
// (B = number of set neighbours, A = 0 -> 1 transitions around the ring N NE E SE S SW W NW).  This is the closed form of the
// built-in 256-entry table (capi.cu zhang_suen_table); the kernels use it only when the installed table IS that table, and
// tests/hostcheck compares the two on all 256 neighbourhoods.  ~55 word operations instead of ~20 per border pixel.
FPB_HD uint32_t fpb_zs_delete_mask(uint32_t NW, uint32_t N, uint32_t NE, uint32_t E, uint32_t SE, uint32_t S, uint32_t SW,
                                   uint32_t W, int pass) {
    // B: bit-sliced sum of the eight planes
    const uint32_t s1 = N ^ NE ^ E, c1 = (N & NE) | (E & (N ^ NE));
    const uint32_t s2 = SE ^ S ^ SW, c2 = (SE & S) | (SW & (SE ^ S));
    const uint32_t s3 = W ^ NW, c3 = W & NW;
    const uint32_t b0 = s1 ^ s2 ^ s3, c4 = (s1 & s2) | (s3 & (s1 ^ s2));
    const uint32_t t1 = c1 ^ c2 ^ c3, d1 = (c1 & c2) | (c3 & (c1 ^ c2));
    const uint32_t b1 = t1 ^ c4, d2 = t1 & c4;
    const uint32_t b2 = d1 ^ d2, b3 = d1 & d2;
    const uint32_t okB = ~b3 & (b2 | b1) & ~(b2 & b1 & b0);              // B not in {0, 1, 7, 8}
    // A == 1: exactly one 0 -> 1 transition around the ring
    const uint32_t tr[8] = {~N & NE, ~NE & E, ~E & SE, ~SE & S, ~S & SW, ~SW & W, ~W & NW, ~NW & N};
    uint32_t one = 0u, two = 0u;
    for (int i = 0; i < 8; ++i) { two |= one & tr[i]; one |= tr[i]; }
    const uint32_t okA = one & ~two;
    const uint32_t okP = pass == 1 ? ~(E & S & (N | W)) : ~(N & W & (E | S));
    return okB & okA & okP;
}
