// K6: smooth_fingerprint_skeleton  (/root/reference/src/preprocessing/fingerprint_preprocess.py:141-159)
//   img = bin/255; (nx,ny) = sobel(img)/(|grad|+1e-6); 3x: smoothed += 1.4*(dx*ny - dy*nx) with fresh
//   sobels of `smoothed`; gaussian_filter(0.6); *1.25, clip; > 0.35.
//
// scipy.ndimage.sobel(axis=a) = correlate1d([-1,0,1]) along a, then correlate1d([1,2,1]) along the other
// axis, float64 accumulation, FLOAT32 intermediate, mode='reflect' - reproduced literally, so this
// stage is bit-exact (float32 ops evaluated in numpy's order, nvcc -fmad=false).
#include "fpb_kernels.h"
#include "hd_scalar.h"

struct GaussW5 { double w[5]; };


struct Sobel2 { float dx, dy; };

// both scipy sobels of plane p at (x,y); T = float (plane) or uint8 {0,255} read as {0,1}
template <typename T> __device__ __forceinline__ float ldv(const T* p, size_t i);
template <> __device__ __forceinline__ float ldv<float>(const float* p, size_t i) { return p[i]; }
template <> __device__ __forceinline__ float ldv<uint8_t>(const uint8_t* p, size_t i) { return (float)p[i] / 255.0f; }

template <typename T>
__device__ __forceinline__ Sobel2 sobel_at(const T* p, int W, int w, int h, int x, int y) {
    const int xm = fpb_reflect_dup(x - 1, w), xp = fpb_reflect_dup(x + 1, w);
    const int ym = fpb_reflect_dup(y - 1, h), yp = fpb_reflect_dup(y + 1, h);
    float a[3][3];
    const int ys[3] = {ym, y, yp}, xs[3] = {xm, x, xp};
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) a[j][i] = ldv<T>(p, (size_t)ys[j] * W + xs[i]);
    Sobel2 s;
    // axis=1: derivative along x (float32 intermediate), then [1,2,1] along y
    // (float)((double)p - (double)q) is the correctly rounded float difference == the float subtraction itself
    const float d0 = a[0][2] - a[0][0], d1 = a[1][2] - a[1][0], d2 = a[2][2] - a[2][0];
    s.dx = (float)((double)d1 * 2.0 + ((double)d0 + (double)d2));
    // axis=0: derivative along y, then [1,2,1] along x
    const float e0 = a[2][0] - a[0][0], e1 = a[2][1] - a[0][1], e2 = a[2][2] - a[0][2];
    s.dy = (float)((double)e1 * 2.0 + ((double)e0 + (double)e2));
    return s;
}

__global__ void k_sm_init(const uint8_t* __restrict__ bin, int W, int H, const int4* __restrict__ roi,
                          float* __restrict__ ux, float* __restrict__ uy, float* __restrict__ acc) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x >= d.w || y >= d.h) return;
    const size_t base = (size_t)b * W * H, o = base + (size_t)y * W + x;
    const Sobel2 s = sobel_at<uint8_t>(bin + base, W, d.w, d.h, x, y);
    const float mag = sqrtf(s.dx * s.dx + s.dy * s.dy) + 1e-6f;
    ux[o] = s.dx / mag; uy[o] = s.dy / mag;
    acc[o] = (float)bin[o] / 255.0f;
}

__global__ void k_sm_step(const float* __restrict__ acc, int W, int H, const int4* __restrict__ roi,
                          const float* __restrict__ ux, const float* __restrict__ uy, float* __restrict__ out,
                          float step /* `sigma` of the reference: 1.4 */) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x >= d.w || y >= d.h) return;
    const size_t base = (size_t)b * W * H, o = base + (size_t)y * W + x;
    const Sobel2 s = sobel_at<float>(acc + base, W, d.w, d.h, x, y);
    const float proj = s.dx * uy[o] - s.dy * ux[o];
    out[o] = acc[o] + step * proj;
}

__global__ void k_sm_final(const float* __restrict__ sm, int W, int H, const int4* __restrict__ roi,
                           uint8_t* __restrict__ dst, float boost /* contrast_boost: 1.25 */) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x >= d.w || y >= d.h) return;
    const size_t o = (size_t)b * W * H + (size_t)y * W + x;
    float v = sm[o] * boost;
    v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    dst[o] = (v > 0.35f) ? 255 : 0;
}

// ------------------------------------------------------------------------------------------------
// Fused K6: one CTA = 32x32 output pixels; the binary tile (+6 halo), the unit-gradient field and the three
// diffusion steps live in shared memory, then the 5-tap Gaussian, the boost and the threshold.  Every stage applies
// scipy's 'reflect' at the IMAGE border on the array of that stage (a mirrored neighbour is looked up at its
// in-image position, which is always inside the staged region), so the arithmetic is identical to the five-kernel
// sequence above - which stays as the path for the stage-by-stage diagnostics.
// ------------------------------------------------------------------------------------------------
#define SF_T 32
#define SF_H 6                       // 1 (unit field) + 3 (steps) + 2 (gaussian radius)
#define SF_IN (SF_T + 2 * SF_H)      // 44
#define SF_P (SF_IN + 1)

struct SfTile {
    int ox, oy, w, h;                // image coordinates of tile-local (0,0); image size
    const int *rx, *ry;              // border tiles: reflected tile-local column / row*SF_P of local coordinate k-2
    __device__ __forceinline__ bool inside(int lx, int ly) const {
        return (unsigned)(ox + lx) < (unsigned)w && (unsigned)(oy + ly) < (unsigned)h;
    }
    // tile-local index of image pixel reflect(ox+lx), reflect(oy+ly): two table look-ups instead of the reflect
    // arithmetic (border tiles are 43 % of a 222x315 crop and every tap of every stage goes through here)
    __device__ __forceinline__ int at(int lx, int ly) const { return ry[ly + 2] + rx[lx + 2]; }
    __device__ __forceinline__ void fill_tables(int* rxs, int* rys, int tid) {
        for (int k = tid; k < SF_IN + 4; k += 256) {
            rxs[k] = fpb_reflect_dup(ox + k - 2, w) - ox;
            rys[k] = (fpb_reflect_dup(oy + k - 2, h) - oy) * SF_P;
        }
        rx = rxs; ry = rys;
    }
};

// EXACT: the tile holds small integers (first stage: sobel of the {0,1} image) - float arithmetic is exact there
template <bool INTERIOR, bool EXACT>
__device__ __forceinline__ Sobel2 sobel_tile(const float* a, const SfTile& t, int lx, int ly) {
    float v[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i)
            v[j][i] = INTERIOR ? a[(ly + j - 1) * SF_P + lx + i - 1] : a[t.at(lx + i - 1, ly + j - 1)];
    Sobel2 s;
    const float d0 = v[0][2] - v[0][0], d1 = v[1][2] - v[1][0], d2 = v[2][2] - v[2][0];
    const float e0 = v[2][0] - v[0][0], e1 = v[2][1] - v[0][1], e2 = v[2][2] - v[0][2];
    if (EXACT) { s.dx = d1 * 2.0f + (d0 + d2); s.dy = e1 * 2.0f + (e0 + e2); }
    else {
        s.dx = (float)((double)d1 * 2.0 + ((double)d0 + (double)d2));
        s.dy = (float)((double)e1 * 2.0 + ((double)e0 + (double)e2));
    }
    return s;
}

template <bool INTERIOR>
__device__ __forceinline__ void smooth_tile_body(const uint8_t* __restrict__ bin, int W, int H, int b, const SfTile& t, int x0, int y0,
                                                 const GaussW5& g, uint8_t* __restrict__ dst, float* base, float* ux, float* uy,
                                                 float* a0, float* a1, float2* unit_lut) {
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const uint8_t* p = bin + (size_t)b * W * H;
    {   // tile load: thread (tx, ty) takes columns tx, tx+32 of rows ty, ty+8, ...; addresses are clamped into the image
        // so that all twelve byte loads of a thread are unconditional and in flight together (the phase was bound by
        // the latency of one dependent byte load per trip)
        const int tx = threadIdx.x, ty = threadIdx.y;
        const int ca = min(max(t.ox + tx, 0), t.w - 1), cb = min(max(t.ox + tx + 32, 0), t.w - 1);
        uint8_t va[6], vb[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int ly = ty + 8 * k;
            const uint8_t* q = p + (size_t)min(max(t.oy + ly, 0), t.h - 1) * W;
            va[k] = q[ca]; vb[k] = q[cb];
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int ly = ty + 8 * k;
            if (ly < SF_IN) {
                if (t.inside(tx, ly)) base[ly * SF_P + tx] = (float)va[k] / 255.0f;
                if (tx + 32 < SF_IN && t.inside(tx + 32, ly)) base[ly * SF_P + tx + 32] = (float)vb[k] / 255.0f;
            }
        }
        // the Sobel response of a {0,1} image is a pair of integers in [-4, 4]: the 81 possible unit vectors are formed
        // once per CTA with the reference's float32 operations (sqrt, + 1e-6, two divisions) instead of once per pixel
        if (tid < 81) {
            const float dx = (float)(tid % 9 - 4), dy = (float)(tid / 9 - 4);
            const float mag = sqrtf(dx * dx + dy * dy) + 1e-6f;
            unit_lut[tid] = make_float2(dx / mag, dy / mag);
        }
    }
    __syncthreads();
    // unit gradient field and acc0 on halo-1 region (margin m = 1 from the staged border)
    for (int i = tid; i < SF_IN * SF_IN; i += 256) {
        const int ly = i / SF_IN, lx = i - ly * SF_IN;
        if (lx < 1 || ly < 1 || lx >= SF_IN - 1 || ly >= SF_IN - 1 || !t.inside(lx, ly)) continue;
        const Sobel2 s = sobel_tile<INTERIOR, true>(base, t, lx, ly);
        const int ix = (int)s.dx, iy = (int)s.dy;
        if ((float)ix == s.dx && (float)iy == s.dy && ix >= -4 && ix <= 4 && iy >= -4 && iy <= 4) {
            const float2 u = unit_lut[(iy + 4) * 9 + ix + 4];
            ux[ly * SF_P + lx] = u.x; uy[ly * SF_P + lx] = u.y;
        } else {                                            // not a {0,255} input plane: the general expressions
            const float mag = sqrtf(s.dx * s.dx + s.dy * s.dy) + 1e-6f;
            ux[ly * SF_P + lx] = s.dx / mag; uy[ly * SF_P + lx] = s.dy / mag;
        }
    }
    __syncthreads();
    // three explicit steps: src -> dst on shrinking regions (margins 2, 3, 4)
    const float* src = base; float* dstp = a0;
    for (int step = 0; step < 3; ++step) {
        const int m = 2 + step;
        for (int i = tid; i < SF_IN * SF_IN; i += 256) {
            const int ly = i / SF_IN, lx = i - ly * SF_IN;
            if (lx < m || ly < m || lx >= SF_IN - m || ly >= SF_IN - m || !t.inside(lx, ly)) continue;
            const Sobel2 s = sobel_tile<INTERIOR, false>(src, t, lx, ly);
            const int o = ly * SF_P + lx;
            const float proj = s.dx * uy[o] - s.dy * ux[o];
            dstp[o] = src[o] + 1.4f * proj;
        }
        __syncthreads();
        src = dstp; dstp = (dstp == a0) ? a1 : a0;
    }
    // src = acc after 3 steps, valid on margin 4.  gaussian_filter(0.6): axis 0 then axis 1, float64 accumulation,
    // float32 intermediate (stored in the free buffer), 'reflect'
    float* mid = dstp;
    for (int i = tid; i < SF_IN * SF_IN; i += 256) {
        const int ly = i / SF_IN, lx = i - ly * SF_IN;
        if (lx < 4 || lx >= SF_IN - 4 || ly < SF_H || ly >= SF_IN - SF_H || !t.inside(lx, ly)) continue;
        double acc = (double)src[ly * SF_P + lx] * g.w[2];
#pragma unroll
        for (int ll = -2; ll < 0; ++ll)
            acc += (INTERIOR ? ((double)src[(ly + ll) * SF_P + lx] + (double)src[(ly - ll) * SF_P + lx])
                             : ((double)src[t.at(lx, ly + ll)] + (double)src[t.at(lx, ly - ll)])) * g.w[ll + 2];
        mid[ly * SF_P + lx] = (float)acc;
    }
    __syncthreads();
    for (int i = tid; i < SF_T * SF_T; i += 256) {
        const int ry = i / SF_T, rx = i - ry * SF_T, lx = rx + SF_H, ly = ry + SF_H;
        if (!t.inside(lx, ly)) continue;
        double acc = (double)mid[ly * SF_P + lx] * g.w[2];
#pragma unroll
        for (int ll = -2; ll < 0; ++ll)
            acc += (INTERIOR ? ((double)mid[ly * SF_P + lx + ll] + (double)mid[ly * SF_P + lx - ll])
                             : ((double)mid[t.at(lx + ll, ly)] + (double)mid[t.at(lx - ll, ly)])) * g.w[ll + 2];
        float v = (float)acc * 1.25f;
        v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
        dst[(size_t)b * W * H + (size_t)(y0 + ry) * W + x0 + rx] = (v > 0.35f) ? 255 : 0;
    }
}

__global__ void __launch_bounds__(256)
k_smooth_fused(const uint8_t* __restrict__ bin, int W, int H, const int4* __restrict__ roi, GaussW5 g,
               uint8_t* __restrict__ dst) {
    __shared__ float base[SF_IN * SF_P], ux[SF_IN * SF_P], uy[SF_IN * SF_P], a0[SF_IN * SF_P], a1[SF_IN * SF_P];
    __shared__ int rxs[SF_IN + 4], rys[SF_IN + 4];
    __shared__ float2 unit_lut[81];
    const int b = blockIdx.z;
    const FpbDims d = fpb_dims(roi, b, W, H);
    const int x0 = blockIdx.x * SF_T, y0 = blockIdx.y * SF_T;
    if (x0 >= d.w || y0 >= d.h) return;
    SfTile t; t.ox = x0 - SF_H; t.oy = y0 - SF_H; t.w = d.w; t.h = d.h;
    // tiles whose staged region lies wholly inside the image never reflect: plain indexing
    const bool interior = t.ox >= 0 && t.oy >= 0 && t.ox + SF_IN <= d.w && t.oy + SF_IN <= d.h;
    t.rx = rxs; t.ry = rys;
    if (interior) smooth_tile_body<true>(bin, W, H, b, t, x0, y0, g, dst, base, ux, uy, a0, a1, unit_lut);
    else { t.fill_tables(rxs, rys, threadIdx.y * blockDim.x + threadIdx.x); smooth_tile_body<false>(bin, W, H, b, t, x0, y0, g, dst, base, ux, uy, a0, a1, unit_lut); }
}

void fpb_smooth_core(FpbLaunch L, const uint8_t* binary, int n, int W, int H, const int4* roi,
                     float* ux, float* uy, float* acc, float* acc2, float* tmp, uint8_t* dst, const FpbSmoothPrm* prm) {
    const dim3 blk(32, 8), grid((W + 31) / 32, (H + 7) / 8, n);
    if (ux == nullptr) {            // fused path (the pipeline); the planes are only needed by the unfused sequence
        GaussW5 g; double w[64];
        const int r = fpb_gauss_weights_fill(0.6, w, 64);
        if (r == 2) {
            for (int i = 0; i < 5; ++i) g.w[i] = w[i];
            const dim3 gt((W + SF_T - 1) / SF_T, (H + SF_T - 1) / SF_T, n);
            k_smooth_fused<<<gt, blk, 0, L.st>>>(binary, W, H, roi, g, dst);  LAUNCH_COUNT(L);
            return;
        }
    }
    // unfused sequence: the three keyword arguments of the reference as data (`prm` == nullptr: 1.4, 3, 1.25).  NumPy keeps the
    // Python-float arguments weak, so `sigma * grad_proj` and `smoothed * contrast_boost` are float32 products.
    const float step = prm ? (float)prm->sigma : 1.4f, boost = prm ? (float)prm->contrast_boost : 1.25f;
    const int iters = prm ? prm->diffusion_iter : 3;
    k_sm_init<<<grid, blk, 0, L.st>>>(binary, W, H, roi, ux, uy, acc);          LAUNCH_COUNT(L);
    float *cur = acc, *nxt = acc2;
    for (int it = 0; it < iters; ++it) {
        k_sm_step<<<grid, blk, 0, L.st>>>(cur, W, H, roi, ux, uy, nxt, step);   LAUNCH_COUNT(L);
        float* sw = cur; cur = nxt; nxt = sw;
    }
    fpb_gaussian_f32(L, cur, n, W, H, roi, 0.6, tmp, nxt);
    k_sm_final<<<grid, blk, 0, L.st>>>(nxt, W, H, roi, dst, boost);             LAUNCH_COUNT(L);
}
