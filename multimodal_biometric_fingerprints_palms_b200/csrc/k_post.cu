// K9: postprocess_minutiae  (/root/reference/src/features/post_processing.py:69-137)
//   density = cv2.blur(skel>0, 25x25)/max ; orientation/coherence = compute_orientation_map(skel) (K5) ;
//   per-minutia gating + score (:97-128) ; nms_adaptive (:10-32, including its last-writer-wins
//   behaviour) ; remove_redundant_oriented_adaptive (:37-64) ; stable sort by quality, first 60 (:135).
//
// Scoring is one thread per raw minutia; the order-dependent list logic (NMS / redundancy / sort) is
// literally serial in the reference, so it runs as one thread per image over the <= few hundred
// survivors.  Python-float arithmetic of the reference is float64 here; NumPy-scalar arithmetic that
// NEP 50 keeps in float32 (the adaptive radii, exp(-3*std)) is float32 here.
#include "fpb_kernels.h"
#include <math.h>


#define DN_T 32
#define DN_RMAX 16          // windows up to 33 x 33; cv2.blur anchors an even window at win / 2: [x - win/2, x - win/2 + win - 1]
#define DN_IN (DN_T + 2 * DN_RMAX)

// cv2.blur of the {0,1} skeleton (float32, normalised, BORDER_REFLECT_101): exact integer window counts
// (sliding window sums as in k_box25_stats: the counts are integers, any order is exact)
__global__ void __launch_bounds__(256)
k_density(const uint8_t* __restrict__ skel, int W, int H, const int4* __restrict__ roi, int win,
          float* __restrict__ dens, unsigned* __restrict__ dmax_bits) {
    __shared__ uint8_t tin[DN_IN][68];
    __shared__ int h1[DN_IN][DN_T + 1];
    const int b = blockIdx.z;
    const FpbDims d = fpb_dims(roi, b, W, H);
    const int x0 = blockIdx.x * DN_T, y0 = blockIdx.y * DN_T;
    if (x0 >= d.w || y0 >= d.h) return;
    const int r = win / 2, in = DN_T + 2 * r;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const uint8_t* p = skel + (size_t)b * W * H;
    {   // tile load: the reflected column indices once per thread, the row index once per row
        const int tx = threadIdx.x, ty = threadIdx.y;
        const int gxa = fpb_reflect101(x0 - r + tx, d.w), gxb = fpb_reflect101(x0 - r + tx + 32, d.w);
        for (int rr = ty; rr < in; rr += 8) {
            const uint8_t* q = p + (size_t)fpb_reflect101(y0 - r + rr, d.h) * W;
            tin[rr][tx] = q[gxa] != 0;
            if (tx + 32 < in) tin[rr][tx + 32] = q[gxb] != 0;
        }
    }
    __syncthreads();
    for (int i = tid; i < in * (DN_T / 8); i += 256) {               // item = (row, segment of 8 outputs)
        const int seg = i / in, rr = i - seg * in, c0 = seg * 8;
        int s = 0;
        for (int k = 0; k < win; ++k) s += tin[rr][c0 + k];
        h1[rr][c0] = s;
#pragma unroll
        for (int j = 1; j < 8; ++j) { s += (int)tin[rr][c0 + j + win - 1] - (int)tin[rr][c0 + j - 1]; h1[rr][c0 + j] = s; }
    }
    __syncthreads();
    float lmax = 0.0f;
    const double scale = 1.0 / (double)(win * win);
    {                                                                 // item = (column, segment of 4 rows)
        const int c = tid & 31, r0 = (tid >> 5) * 4;
        const int gx = x0 + c;
        int s = 0;
        for (int k = 0; k < win; ++k) s += h1[r0 + k][c];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j) s += h1[r0 + j + win - 1][c] - h1[r0 + j - 1][c];
            const int gy = y0 + r0 + j;
            if (gx >= d.w || gy >= d.h) continue;
            const float v = (float)((double)s * scale);
            dens[(size_t)b * W * H + (size_t)gy * W + gx] = v;
            lmax = fmaxf(lmax, v);
        }
    }
    for (int off = 16; off; off >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, off));
    if ((tid & 31) == 0) atomicMax(&dmax_bits[b], __float_as_uint(lmax));
}

void fpb_density(FpbLaunch L, const uint8_t* skel, int n, int W, int H, const int4* roi, int win,
                 float* dens, unsigned* dmax_bits) {
    cudaMemsetAsync(dmax_bits, 0, (size_t)n * sizeof(unsigned), L.st);
    dim3 blk(32, 8), grid((W + DN_T - 1) / DN_T, (H + DN_T - 1) / DN_T, n);
    k_density<<<grid, blk, 0, L.st>>>(skel, W, H, roi, win, dens, dmax_bits);
    LAUNCH_COUNT(L);
}

// scratch layout per image (doubles): [0]=n_scored, then raw_cap records of 8 doubles:
//   x, y, type, orientation, quality, coherence, stability, density(float32 value)
#define REC 8
#define SCR_PER_IMG(raw_cap) (1 + (size_t)(raw_cap) * REC)

// one thread per raw minutia: gates and score (:97-128); survivors flagged in place, compacted serially later
__global__ void k_post_score(const uint8_t* __restrict__ skel, const float* __restrict__ dens,
                             const unsigned* __restrict__ dmax_bits, const float* __restrict__ orient,
                             const float* __restrict__ coher, int W, int H, const int4* __restrict__ roi,
                             const int* __restrict__ raw_count, const uint32_t* __restrict__ raw, int raw_cap, FpbPost prm,
                             double* __restrict__ scratch) {
    const int b = blockIdx.y;
    const int nraw = min(raw_count[b], raw_cap);    // a count above raw_cap is reported as FPB_E_OVERFLOW by the host
    const FpbDims d = fpb_dims(roi, b, W, H);
    const int w = d.w, h = d.h;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nraw; i += gridDim.x * blockDim.x) {
    const uint32_t pk = raw[(size_t)b * raw_cap + i];
    const int x = pk & 0x3FFF, y = (pk >> 14) & 0x3FFF, type = (pk >> 28) & 1;
    double* rec = scratch + (size_t)b * SCR_PER_IMG(raw_cap) + 1 + (size_t)i * REC;
    rec[4] = -1.0;                                  // quality < 0 marks "dropped"
    if (!(prm.margin <= x && x < w - prm.margin && prm.margin <= y && y < h - prm.margin)) continue;
    const size_t base = (size_t)b * W * H, o = base + (size_t)y * W + x;
    const float dmax = __uint_as_float(dmax_bits[b]);
    const float dn32 = dens[o] / (dmax + 1e-6f);                       // density /= (density.max() + 1e-6)
    float c32 = coher[o]; c32 = c32 < 0.0f ? 0.0f : (c32 > 1.0f ? 1.0f : c32);
    const double local_coh = (double)c32, local_den = (double)dn32;
    if (local_den < prm.quality_threshold || local_coh < prm.coherence_threshold) continue;
    // angular stability over orient[y-r:y+r, x-r:x+r] (r=15 -> 30x30), np.std in float32, exp in float32
    const int ya = max(0, y - prm.patch_radius), yb = min(h, y + prm.patch_radius);
    const int xa = max(0, x - prm.patch_radius), xb = min(w, x + prm.patch_radius);
    double stab = 0.0;
    const int cnt = (yb - ya) * (xb - xa);
    if (cnt > 0) {
        double s = 0.0;
        for (int yy = ya; yy < yb; ++yy) for (int xx = xa; xx < xb; ++xx) s += (double)orient[base + (size_t)yy * W + xx];
        const float mean32 = (float)(s / cnt);
        double v = 0.0;
        for (int yy = ya; yy < yb; ++yy) for (int xx = xa; xx < xb; ++xx) {
            const float dv = orient[base + (size_t)yy * W + xx] - mean32;
            v += (double)(dv * dv);
        }
        const float std32 = sqrtf((float)(v / cnt));
        stab = (double)expf(-3.0f * std32);
    }
    const double hw = (double)w / 2.0, hh = (double)h / 2.0;
    const double ax = fabs((double)x - hw) / hw, ay = fabs((double)y - hh) / hh;
    const double bonus = 1.0 - 0.5 * (ax * ax + ay * ay);
    const double lit = skel[o] != 0 ? 1.0 : 0.0;
    const double q = (0.5 * local_coh + 0.25 * local_den + 0.1 * stab + 0.1 * lit) * bonus;
    rec[0] = x; rec[1] = y; rec[2] = type; rec[3] = (double)orient[o];
    rec[4] = q; rec[5] = local_coh; rec[6] = stab; rec[7] = (double)dn32;
  }
}

// one thread per image: compaction in list order, NMS, redundancy removal, stable sort, top-K
__global__ void k_post_select(int n, const int* __restrict__ raw_count, int raw_cap, FpbPost prm, double* __restrict__ scratch,
                              int* __restrict__ idx_ws /* [n][3*raw_cap] */, int* __restrict__ out_count,
                              FpbMinutiaDev* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const int nraw = min(raw_count[b], raw_cap);
    double* recs = scratch + (size_t)b * SCR_PER_IMG(raw_cap) + 1;
    int* live = idx_ws + (size_t)b * 3 * raw_cap;   // indices of scored candidates, list order
    int* order = live + raw_cap;
    int* flag = order + raw_cap;
    int m = 0;
    for (int i = 0; i < nraw; ++i) if (recs[(size_t)i * REC + 4] >= 0.0) live[m++] = i;
#define R(k, f) recs[(size_t)live[k] * REC + (f)]
    // ---- nms_adaptive: visit by descending quality; every visited point re-instates itself and clears all
    //      neighbours within base_dist/(0.5+density)  (float32 radius; closed ball on squared distances)
    for (int k = 0; k < m; ++k) {                  // stable insertion sort of indices by -quality
        int j = k;
        const double qk = R(k, 4);
        while (j > 0 && R(order[j - 1], 4) < qk) { order[j] = order[j - 1]; --j; }
        order[j] = k;
    }
    for (int k = 0; k < m; ++k) flag[k] = 0;
    for (int t = 0; t < m; ++t) {
        const int i = order[t];
        if (flag[i]) continue;
        const float rad32 = (float)prm.min_distance / (0.5f + (float)R(i, 7));
        const double r2 = (double)rad32 * (double)rad32;
        flag[i] = 1;
        const int xi = (int)R(i, 0), yi = (int)R(i, 1);
        for (int j = 0; j < m; ++j) {
            if (j == i) continue;
            const int dx = (int)R(j, 0) - xi, dy = (int)R(j, 1) - yi;
            if ((double)(dx * dx + dy * dy) <= r2) flag[j] = 0;
        }
    }
    int m2 = 0;
    for (int k = 0; k < m; ++k) if (flag[k]) live[m2++] = live[k];      // keeps list order
    m = m2;
    // ---- remove_redundant_oriented_adaptive
    for (int k = 0; k < m; ++k) flag[k] = 0;                            // 1 = removed
    const double ang_thresh = 30.0 * (3.14159265358979323846 / 180.0);  // np.deg2rad(30)
    for (int i = 0; i < m; ++i) {
        if (flag[i]) continue;
        const double qi = R(i, 4);
        const float rad32 = (float)(20.0 * (1.0 + (1.0 - qi))) / (0.5f + (float)R(i, 7));
        const double r2 = (double)rad32 * (double)rad32;
        const int xi = (int)R(i, 0), yi = (int)R(i, 1);
        for (int j = i + 1; j < m; ++j) {
            if (flag[j]) continue;
            const int dx = (int)R(j, 0) - xi, dy = (int)R(j, 1) - yi;
            if ((double)(dx * dx + dy * dy) > r2) continue;
            const double dth = R(i, 3) - R(j, 3);
            const double ad = fabs(atan2(sin(dth), cos(dth)));
            if (ad < ang_thresh) flag[(qi < R(j, 4)) ? i : j] = 1;
        }
    }
    m2 = 0;
    for (int k = 0; k < m; ++k) if (!flag[k]) live[m2++] = live[k];
    m = m2;
    // ---- sorted(..., key=quality, reverse=True)[:max_m]  (stable)
    for (int k = 0; k < m; ++k) {
        int j = k;
        const double qk = R(k, 4);
        while (j > 0 && R(order[j - 1], 4) < qk) { order[j] = order[j - 1]; --j; }
        order[j] = k;
    }
    const int keep = min(min(m, prm.max_minutiae), FPB_MAX_REFINED);
    for (int t = 0; t < keep; ++t) {
        const int k = order[t];
        FpbMinutiaDev o;
        o.x = (int)R(k, 0); o.y = (int)R(k, 1); o.type = (int)R(k, 2); o.pad = 0;
        o.orientation = R(k, 3); o.quality = R(k, 4); o.coherence = R(k, 5); o.angular_stability = R(k, 6);
        out[(size_t)b * FPB_MAX_REFINED + t] = o;
    }
    out_count[b] = keep;
#undef R
}

void fpb_postprocess_core(FpbLaunch L, const uint8_t* skel, const float* dens, const unsigned* dmax_bits,
                          const float* orient, const float* coher, int n, int W, int H, const int4* roi,
                          const int* raw_count, const uint32_t* raw, int raw_cap, FpbPost prm, int* out_count,
                          FpbMinutiaDev* out, double* scratch, int* idx_ws) {
    dim3 g1(16, n);                                  // 2048 threads per image, strided over the raw list
    k_post_score<<<g1, 128, 0, L.st>>>(skel, dens, dmax_bits, orient, coher, W, H, roi, raw_count, raw, raw_cap, prm, scratch);
    LAUNCH_COUNT(L);
    k_post_select<<<(n + 31) / 32, 32, 0, L.st>>>(n, raw_count, raw_cap, prm, scratch, idx_ws, out_count, out);
    LAUNCH_COUNT(L);
}

// ------------------------------------------------------------------------------------------------
// Stand-alone nms_adaptive (post_processing.py:10-32) / remove_redundant_oriented_adaptive (:37-64) on caller-supplied
// lists: one image, one thread (the reference's loops are order dependent).  buf layout (doubles): x[n] y[n] q[n]
// ori[n] dens[n] then ints order[n] keep[n].
// ------------------------------------------------------------------------------------------------
__global__ void k_minutiae_select(int mode, int n, const double* __restrict__ buf, double p0, double p1, int* __restrict__ iws,
                                  unsigned char* __restrict__ keep_out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const double *X = buf, *Y = buf + n, *Q = buf + 2 * n, *O = buf + 3 * n, *D = buf + 4 * n;
    int* order = iws; int* flag = iws + n;
    if (mode == 1) {                       // nms_adaptive: p0 = base_dist
        for (int k = 0; k < n; ++k) {      // np.argsort(-q): descending quality (ties: list order)
            int j = k;
            while (j > 0 && Q[order[j - 1]] < Q[k]) { order[j] = order[j - 1]; --j; }
            order[j] = k;
        }
        for (int k = 0; k < n; ++k) flag[k] = 0;
        for (int t = 0; t < n; ++t) {
            const int i = order[t];
            if (flag[i]) continue;
            const float rad32 = (float)p0 / (0.5f + (float)D[i]);
            const double r2 = (double)rad32 * (double)rad32;
            flag[i] = 1;
            for (int j = 0; j < n; ++j) {
                if (j == i) continue;
                const double dx = X[j] - X[i], dy = Y[j] - Y[i];
                if (dx * dx + dy * dy <= r2) flag[j] = 0;
            }
        }
        for (int k = 0; k < n; ++k) keep_out[k] = (unsigned char)flag[k];
    } else {                               // remove_redundant_oriented_adaptive: p0 = base_radius, p1 = angle_thresh
        for (int k = 0; k < n; ++k) flag[k] = 0;
        for (int i = 0; i < n; ++i) {
            if (flag[i]) continue;
            const float rad32 = (float)(p0 * (1.0 + (1.0 - Q[i]))) / (0.5f + (float)D[i]);
            const double r2 = (double)rad32 * (double)rad32;
            for (int j = i + 1; j < n; ++j) {
                if (flag[j]) continue;
                const double dx = X[j] - X[i], dy = Y[j] - Y[i];
                if (dx * dx + dy * dy > r2) continue;
                const double dth = O[i] - O[j];
                if (fabs(atan2(sin(dth), cos(dth))) < p1) flag[(Q[i] < Q[j]) ? i : j] = 1;
            }
        }
        for (int k = 0; k < n; ++k) keep_out[k] = (unsigned char)!flag[k];
    }
}

void fpb_minutiae_select(FpbLaunch L, int mode, int n, const double* buf, double p0, double p1, int* iws, unsigned char* keep_out) {
    k_minutiae_select<<<1, 32, 0, L.st>>>(mode, n, buf, p0, p1, iws, keep_out);
    LAUNCH_COUNT(L);
}
