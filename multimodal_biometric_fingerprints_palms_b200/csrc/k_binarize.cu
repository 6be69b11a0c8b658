// K4: binarize  (/root/reference/src/preprocessing/fingerprint_preprocess.py:43-81), the float part:
//   mean / mean-of-squares over 25x25 (cv2.boxFilter, normalised, BORDER_REFLECT_101), std,
//   adaptive Sauvola threshold, 32x32 patch Otsu OR-ed in; plus the 3x3 cross erode/dilate used by
//   the opening and the marker (:76-79).  The connected-component passes are in k_ccl.cu.
//
// cv2.boxFilter on float32 accumulates in double; the inputs are integers (CLAHE output and its
// square), so the window sums are exact and the kernel carries them as int32:
//   mean = float(double(S1) * (1.0/625)),  sqmean = float(double(S2) * (1.0/625)).
// Every float32 expression is evaluated literally (nvcc -fmad=false) in numpy's order.
#include "fpb_kernels.h"
#include "hd_scalar.h"


#define BX_T 32          // output tile
#define BX_R 12          // window radius (25x25)
#define BX_IN (BX_T + 2 * BX_R)   // 56

// The window sums are integers, so any summation order is exact: both passes slide the 25-wide window (one value in,
// one out per step) instead of re-adding 25 values per output - a quarter of the shared-memory loads that bounded the
// first version (LSU pipe at 85 %).  Row pitches of 17 / 33 words keep the lanes of a warp, which walk different rows
// (pass 1) or neighbouring columns (pass 2), on different banks.
#define BX_SEG 8         // outputs per thread in the horizontal pass
#define BX_VSEG 4        // outputs per thread in the vertical pass
__global__ void __launch_bounds__(256)
k_box25_stats(const uint8_t* __restrict__ img, int W, int H, const int4* __restrict__ roi,
              float* __restrict__ mean, float* __restrict__ stdv, unsigned* __restrict__ stdmax_bits) {
    __shared__ uint8_t tin[BX_IN][68];
    __shared__ int h1[BX_IN][BX_T + 1];
    __shared__ int h2[BX_IN][BX_T + 1];
    const int b = blockIdx.z;
    const FpbDims d = fpb_dims(roi, b, W, H);
    const int x0 = blockIdx.x * BX_T, y0 = blockIdx.y * BX_T;
    if (x0 >= d.w || y0 >= d.h) return;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const uint8_t* p = img + (size_t)b * W * H;
    {   // tile load: the reflected column indices once per thread, the row index once per row
        const int tx = threadIdx.x, ty = threadIdx.y;
        const int gxa = fpb_reflect101(x0 - BX_R + tx, d.w), gxb = fpb_reflect101(x0 - BX_R + tx + 32, d.w);
        for (int r = ty; r < BX_IN; r += 8) {
            const uint8_t* q = p + (size_t)fpb_reflect101(y0 - BX_R + r, d.h) * W;
            tin[r][tx] = q[gxa];
            if (tx + 32 < BX_IN) tin[r][tx + 32] = q[gxb];
        }
    }
    __syncthreads();
    for (int i = tid; i < BX_IN * (BX_T / BX_SEG); i += 256) {       // item = (row r, segment of BX_SEG outputs)
        const int seg = i / BX_IN, r = i - seg * BX_IN, c0 = seg * BX_SEG;
        int s1 = 0, s2 = 0;
#pragma unroll
        for (int k = 0; k < 2 * BX_R + 1; ++k) { const int v = tin[r][c0 + k]; s1 += v; s2 += v * v; }
        h1[r][c0] = s1; h2[r][c0] = s2;
#pragma unroll
        for (int j = 1; j < BX_SEG; ++j) {
            const int vin = tin[r][c0 + j + 2 * BX_R], vout = tin[r][c0 + j - 1];
            s1 += vin - vout; s2 += vin * vin - vout * vout;
            h1[r][c0 + j] = s1; h2[r][c0 + j] = s2;
        }
    }
    __syncthreads();
    float lmax = 0.0f;
    {                                                                 // item = (column c, segment of BX_VSEG rows)
        const int c = tid & 31, r0 = (tid >> 5) * BX_VSEG;
        const int gx = x0 + c;
        int s1 = 0, s2 = 0;
#pragma unroll
        for (int k = 0; k < 2 * BX_R + 1; ++k) { s1 += h1[r0 + k][c]; s2 += h2[r0 + k][c]; }
#pragma unroll
        for (int j = 0; j < BX_VSEG; ++j) {
            if (j) {
                s1 += h1[r0 + j + 2 * BX_R][c] - h1[r0 + j - 1][c];
                s2 += h2[r0 + j + 2 * BX_R][c] - h2[r0 + j - 1][c];
            }
            const int gy = y0 + r0 + j;
            if (gx >= d.w || gy >= d.h) continue;
            const double scale = 1.0 / 625.0;
            const float m = (float)((double)s1 * scale);
            const float q = (float)((double)s2 * scale);
            float var = q - m * m;
            if (var < 0.0f) var = 0.0f;
            const float sd = sqrtf(var);
            const size_t o = (size_t)b * W * H + (size_t)gy * W + gx;
            mean[o] = m; stdv[o] = sd;
            lmax = fmaxf(lmax, sd);
        }
    }
    // std >= 0, so the float order equals the order of the bit patterns
    for (int off = 16; off; off >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, off));
    if ((tid & 31) == 0) atomicMax(&stdmax_bits[b], __float_as_uint(lmax));
}

__global__ void k_sauvola(const uint8_t* __restrict__ img, int W, int H, const int4* __restrict__ roi,
                          const float* __restrict__ mean, const float* __restrict__ stdv,
                          const unsigned* __restrict__ stdmax_bits, uint8_t* __restrict__ bin) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x >= d.w || y >= d.h) return;
    const size_t o = (size_t)b * W * H + (size_t)y * W + x;
    const float smax = __uint_as_float(stdmax_bits[b]);
    const float m = mean[o], sd = stdv[o];
    const float sd_n = sd / (smax + 1e-6f);
    const float kmap = 0.25f * (1.0f - 0.5f * sd_n);
    const float thr = m * (1.0f - kmap * (1.0f - sd / (m + 1e-6f)));
    bin[o] = ((float)img[o] < thr) ? 255 : 0;
}

// one block per 32x32 patch (:60-71).  Same arithmetic as fpb_patch_otsu (hd_scalar.h, checked on the host against
// skimage-compat/np.histogram), reorganised for latency: the two float32 cumulative sums whose order matters run
// sequentially on two different warps at the same time, everything else (bin assignment, divisions, variances,
// first-maximum search) is spread over the block.
__global__ void __launch_bounds__(64)
k_patch_otsu(const uint8_t* __restrict__ img, int W, int H, const int4* __restrict__ roi, uint8_t* __restrict__ bin) {
    __shared__ unsigned ih[256];
    __shared__ float counts[256], centers[256], w1a[256], s1a[256], w2a[256], s2a[256], var[256];
    __shared__ float s_t;
    __shared__ int s_go, s_a, s_b, s_arg[64];
    __shared__ long long s_red[2][4];
    const int b = blockIdx.z;
    const FpbDims d = fpb_dims(roi, b, W, H);
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    if (x0 >= d.w || y0 >= d.h) return;
    const int pw = min(32, d.w - x0), ph = min(32, d.h - y0), np = pw * ph;
    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += 64) { ih[i] = 0; counts[i] = 0.0f; }
    __syncthreads();
    const uint8_t* p = img + (size_t)b * W * H;
    {   // lanes along the patch row, the two warps on alternate rows: no division per pixel
        const int c = tid & 31;
        if (c < pw)
            for (int r = tid >> 5; r < ph; r += 2) atomicAdd(&ih[p[(size_t)(y0 + r) * W + x0 + c]], 1u);
    }
    __syncthreads();
    {   // sum, sum of squares, min and max of the patch from the histogram: four values per thread, then a reduction
        long long s1 = 0, s2 = 0;
        int a = 255, bb = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int v = tid * 4 + q;
            const long long hv = ih[v];
            if (hv) { a = min(a, v); bb = max(bb, v); }
            s1 += hv * v; s2 += hv * v * v;
        }
        for (int off = 16; off; off >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, off); s2 += __shfl_xor_sync(0xffffffffu, s2, off);
            a = min(a, __shfl_xor_sync(0xffffffffu, a, off)); bb = max(bb, __shfl_xor_sync(0xffffffffu, bb, off));
        }
        if ((tid & 31) == 0) { s_red[tid >> 5][0] = s1; s_red[tid >> 5][1] = s2; s_red[tid >> 5][2] = a; s_red[tid >> 5][3] = bb; }
        __syncthreads();
        if (tid == 0) {
            s1 = s_red[0][0] + s_red[1][0]; s2 = s_red[0][1] + s_red[1][1];
            // sub.size < 10 or sub.std() < 3  ->  skip ;  std^2 = (n*s2 - s1^2)/n^2
            s_go = (np >= 10) && ((long long)np * s2 - s1 * s1 >= 9ll * np * np);
            s_a = (int)min(s_red[0][2], s_red[1][2]); s_b = (int)max(s_red[0][3], s_red[1][3]);
        }
    }
    __syncthreads();
    if (!s_go) return;
    const int a = s_a, bmax = s_b;      // a < bmax because std >= 3
    const float fa = (float)a, fb = (float)bmax;
    const float norm = fb - fa, step = norm / 256.0f;
    for (int v = a + tid; v <= bmax; v += 64) {          // every integer value lands in its own bin (bin width < 1)
        if (ih[v] == 0) continue;
        const float fv = (float)v;
        int i = (int)(((fv - fa) / norm) * 256.0f);
        if (i == 256) i = 255;
        const float e_i = (float)i * step + fa;
        if (fv < e_i) --i;
        const float e_n = (i + 1 == 256) ? fb : ((float)(i + 1) * step + fa);
        if (fv >= e_n && i != 255) ++i;
        counts[i] = (float)ih[v];
    }
    for (int i = tid; i < 256; i += 64) {
        const float e0 = (float)i * step + fa;
        const float e1 = (i + 1 == 256) ? fb : ((float)(i + 1) * step + fa);
        centers[i] = (e0 + e1) / 2.0f;
    }
    __syncthreads();
    if (tid == 0) {             // np.cumsum(counts), np.cumsum(counts*centers)
        float w = 0.0f, sacc = 0.0f;
        for (int i = 0; i < 256; ++i) { w += counts[i]; sacc += counts[i] * centers[i]; w1a[i] = w; s1a[i] = sacc; }
    } else if (tid == 32) {     // the same on the reversed arrays
        float w = 0.0f, sacc = 0.0f;
        for (int i = 255; i >= 0; --i) { w += counts[i]; sacc += counts[i] * centers[i]; w2a[i] = w; s2a[i] = sacc; }
    }
    __syncthreads();
    float best = -1.0f; int besti = 0;
    for (int i = tid; i < 255; i += 64) {
        const float m1 = s1a[i] / w1a[i], m2 = s2a[i + 1] / w2a[i + 1];
        const float dm = m1 - m2;
        const float vv = (w1a[i] * w2a[i + 1]) * (dm * dm);
        if (vv > best) { best = vv; besti = i; }        // ascending i per thread: keeps the first maximum
    }
    for (int off = 16; off; off >>= 1) {                  // first maximum: larger value, then smaller index
        const float ov = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, off);
        if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    if ((tid & 31) == 0) { var[tid >> 5] = best; s_arg[tid >> 5] = besti; }
    __syncthreads();
    if (tid == 0) {
        float bv = var[0]; int bi = s_arg[0];
        if (var[1] > bv || (var[1] == bv && s_arg[1] < bi)) { bv = var[1]; bi = s_arg[1]; }
        s_t = centers[bi];
    }
    __syncthreads();
    const float t = s_t;
    {
        const int c = tid & 31;
        if (c < pw)
            for (int r = tid >> 5; r < ph; r += 2) {
                const size_t o = (size_t)b * W * H + (size_t)(y0 + r) * W + x0 + c;
                if ((float)p[(size_t)(y0 + r) * W + x0 + c] < t) bin[o] = 255;
            }
    }
}

void fpb_binarize_core(FpbLaunch L, const uint8_t* img_eq, int n, int W, int H, const int4* roi,
                       float* mean, float* stdv, unsigned* stdmax_bits, uint8_t* bin0) {
    cudaMemsetAsync(stdmax_bits, 0, (size_t)n * sizeof(unsigned), L.st);
    dim3 blk(32, 8);
    dim3 gt((W + BX_T - 1) / BX_T, (H + BX_T - 1) / BX_T, n);
    k_box25_stats<<<gt, blk, 0, L.st>>>(img_eq, W, H, roi, mean, stdv, stdmax_bits);       LAUNCH_COUNT(L);
    dim3 gp((W + 31) / 32, (H + 7) / 8, n);
    k_sauvola<<<gp, blk, 0, L.st>>>(img_eq, W, H, roi, mean, stdv, stdmax_bits, bin0);      LAUNCH_COUNT(L);
    dim3 go((W + 31) / 32, (H + 31) / 32, n);
    k_patch_otsu<<<go, 64, 0, L.st>>>(img_eq, W, H, roi, bin0);                              LAUNCH_COUNT(L);
}

// 3x3 cross (cv2.getStructuringElement(MORPH_ELLIPSE,(3,3))) erode / dilate on a {0,255} plane.
// cv2 border: erosion ignores out-of-image neighbours (treated as set), dilation treats them as clear.
__global__ void k_cross3(const uint8_t* __restrict__ src, int W, int H, const int4* __restrict__ roi, int erode,
                         uint8_t* __restrict__ dst) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x >= d.w || y >= d.h) return;
    const uint8_t* p = src + (size_t)b * W * H;
    const size_t o = (size_t)y * W + x;
    const int out = erode ? 1 : 0;      // value of an out-of-image neighbour
    const int c = p[o] != 0;
    const int l = x > 0 ? (p[o - 1] != 0) : out, r = x + 1 < d.w ? (p[o + 1] != 0) : out;
    const int u = y > 0 ? (p[o - W] != 0) : out, dn = y + 1 < d.h ? (p[o + W] != 0) : out;
    const int v = erode ? (c & l & r & u & dn) : (c | l | r | u | dn);
    dst[(size_t)b * W * H + o] = v ? 255 : 0;
}

void fpb_cross3(FpbLaunch L, const uint8_t* src, int n, int W, int H, const int4* roi, int erode, uint8_t* dst) {
    dim3 blk(32, 8), grid((W + 31) / 32, (H + 7) / 8, n);
    k_cross3<<<grid, blk, 0, L.st>>>(src, W, H, roi, erode, dst);
    LAUNCH_COUNT(L);
}
