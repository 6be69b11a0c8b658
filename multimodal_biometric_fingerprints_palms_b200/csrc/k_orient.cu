// K5: compute_orientation_map  (/root/reference/src/preprocessing/orientation.py:9-85)
//   f = img/255 (inverted when any pixel exceeds the median, :26-28) -> gaussian_filter(1.5) -> *255 ->
//   cv2.Sobel -> Gxx,Gyy,Gxy gaussian_filter(3.0) -> reliability sqrt((Gxx-Gyy)^2+4Gxy^2) normalised by
//   its own p2/p98 (np.percentile, float64) -> theta = atan2/2 + pi/2 -> per 16x16 block weighted
//   circular mean -> gaussian_filter(3.0) of sin2/cos2 on the block grid -> cv2.resize (bilinear) to
//   the image -> wrap to [-pi/2, pi/2).
// Also the generic scipy.ndimage.gaussian_filter on float32 planes used by K6/K7.
//
// Float tolerance stage (contract: 1e-4 relative).  The Gaussian passes reproduce SciPy's
// NI_Correlate1D literally (float64 accumulation in SciPy's tap order, float32 intermediate,
// mode='reflect') and are bit-exact; Sobel / atan2 / sin / cos are float32 library functions on both
// sides and agree to a few ulp.
#include "fpb_kernels.h"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#include "hd_scalar.h"
#include "ccl_bits.cuh"
#include <math.h>
#include <type_traits>


static inline dim3 px_grid(int n, int W, int H) { return dim3((W + 31) / 32, (H + 7) / 8, n); }

// ------------------------------------------------------------------------------------------------
// scipy.ndimage.gaussian_filter weights: radius int(4*sigma+0.5), exp(-0.5/sigma^2*x^2)/sum  (float64,
// sum in NumPy's pairwise order so that the table is bit-identical to SciPy's)
// ------------------------------------------------------------------------------------------------
#define GAUSS_MAX_TAPS 64
struct GaussW { double w[GAUSS_MAX_TAPS]; int r; };

GaussW fpb_gauss_weights(double sigma) {
    GaussW g;
    g.r = fpb_gauss_weights_fill(sigma, g.w, GAUSS_MAX_TAPS);
    return g;
}

// one 1-D pass along `axis` (0 = y, 1 = x): NI_Correlate1D symmetric branch
//   acc = a[i]*w[r];  for ll = -r..-1:  acc += (a[i+ll] + a[i-ll]) * w[ll+r]     (float64), store float32
__global__ void k_gauss1d(const float* __restrict__ src, int W, int H, const int4* __restrict__ roi, GaussW g,
                          int axis, float* __restrict__ dst) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x >= d.w || y >= d.h) return;
    const float* p = src + (size_t)b * W * H;
    const int r = g.r;
    double acc;
    if (axis == 0) {
        acc = (double)p[(size_t)y * W + x] * g.w[r];
        for (int ll = -r; ll < 0; ++ll) {
            const double lo = (double)p[(size_t)fpb_reflect_dup(y + ll, d.h) * W + x];
            const double hi = (double)p[(size_t)fpb_reflect_dup(y - ll, d.h) * W + x];
            acc += (lo + hi) * g.w[ll + r];
        }
    } else {
        const float* row = p + (size_t)y * W;
        acc = (double)row[x] * g.w[r];
        for (int ll = -r; ll < 0; ++ll) {
            const double lo = (double)row[fpb_reflect_dup(x + ll, d.w)];
            const double hi = (double)row[fpb_reflect_dup(x - ll, d.w)];
            acc += (lo + hi) * g.w[ll + r];
        }
    }
    dst[(size_t)b * W * H + (size_t)y * W + x] = (float)acc;
}

// Fused two-pass version: one CTA = 32x32 outputs.  The tile (+R halo, 'reflect' resolved at load time) is
// converted to float64 ONCE into shared memory (f32->f64 conversions are a slow pipe; the taps then run on
// DADD/DMUL only), pass 1 (axis 0) writes float32-rounded values back as float64, pass 2 (axis 1) writes the result.
// Same operation order as k_gauss1d => bit-identical output.
#define G2_TX 128
#define G2_TY 16
#define G2_RB 4            // outputs per thread along the filter axis (register blocking)
// 320 threads, not 256: pass 1 has 4 x (128 + 2R) items (608 at R = 12) - 2.4 rounds of 256 threads, i.e. three, with the third
// 37 % full and the other warps waiting at the barrier; with ten warps it is two rounds (pass 2's 512 items leave two warps idle)
#define G2_NY 10
#define G2_NT (32 * G2_NY)
// Shared-memory bandwidth, not FP64 issue, bounded the first version (two 8-byte LDS per tap).  Each thread now
// produces G2_RB consecutive outputs ALONG the filter axis from one register window of G2_RB+2R values, so a tap costs
// 2/G2_RB loads; the weights are read straight from the kernel-parameter constant bank.  Pass 1 stores its result
// transposed so that pass 2's window walk is conflict-free as well.  Operation order per output is unchanged
// (NI_Correlate1D: centre tap, then outermost pair inwards), so the result stays bit-identical to SciPy.
template <int R>
__global__ void __launch_bounds__(G2_NT)
k_gauss2d(const float* __restrict__ src, int W, int H, const int4* __restrict__ roi, const GaussW g, float* __restrict__ dst,
          const uint8_t* __restrict__ src8, const float* __restrict__ flut) {
    constexpr int INX = G2_TX + 2 * R, INY = G2_TY + 2 * R, P = INX + 1, PT = G2_TY + 1, NV = G2_RB + 2 * R;
    extern __shared__ double g2_sm[];
    double* tin = g2_sm;                    // [INY][P]
    double* tmidT = g2_sm + INY * P;        // [INX][PT]   (transposed: column-major)
    const int b = blockIdx.z;
    const FpbDims d = fpb_dims(roi, b, W, H);
    const int x0 = blockIdx.x * G2_TX, y0 = blockIdx.y * G2_TY;
    if (x0 >= d.w || y0 >= d.h) return;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const float* p = src + (size_t)b * W * H;
    // tile load: thread (tx, ty) takes columns tx, tx+32, ... of rows ty, ty+8, ...  The 'reflect' column indices are
    // resolved ONCE per thread (five registers), the row index once per row, so an element costs a load, a conversion
    // and a store - the first version spent half of the kernel's instructions on per-element index arithmetic.
    {
        // Every load of the thread is issued before the first conversion (clamped addresses, no branches): the phase then
        // costs one memory latency instead of one per row group - with the row-by-row loop the load phase of a CTA (eight
        // dependent L2 round trips) was longer than its two filter passes, and the FP64 pipe sat at 49 %.
        constexpr int NC = (INX + 31) / 32, NR = (INY + G2_NY - 1) / G2_NY;
        const int tx = threadIdx.x, ty = threadIdx.y;
        // columns / rows past the crop's own halo are never read by an output of this tile
        const int cmax = min(INX, d.w - x0 + 2 * R), rmax = min(INY, d.h - y0 + 2 * R);
        int gxs[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) gxs[k] = fpb_reflect_dup(x0 - R + min(tx + 32 * k, cmax - 1), d.w);
        float v[NR][NC];
        if (src8) {     // source = u8 image through the per-image 256-entry float map (K5: f = img/255, maybe inverted)
            const float* lut = flut + b * 256;
            uint8_t u[NR][NC];
#pragma unroll
            for (int kr = 0; kr < NR; ++kr) {
                const uint8_t* q = src8 + (size_t)b * W * H + (size_t)fpb_reflect_dup(y0 - R + min(ty + G2_NY * kr, rmax - 1), d.h) * W;
#pragma unroll
                for (int k = 0; k < NC; ++k) u[kr][k] = q[gxs[k]];
            }
#pragma unroll
            for (int kr = 0; kr < NR; ++kr)
#pragma unroll
                for (int k = 0; k < NC; ++k) v[kr][k] = lut[u[kr][k]];
        } else {
#pragma unroll
            for (int kr = 0; kr < NR; ++kr) {
                const float* q = p + (size_t)fpb_reflect_dup(y0 - R + min(ty + G2_NY * kr, rmax - 1), d.h) * W;
#pragma unroll
                for (int k = 0; k < NC; ++k) v[kr][k] = q[gxs[k]];
            }
        }
#pragma unroll
        for (int kr = 0; kr < NR; ++kr)
#pragma unroll
            for (int k = 0; k < NC; ++k)
                if (ty + G2_NY * kr < rmax && tx + 32 * k < cmax) tin[(ty + G2_NY * kr) * P + tx + 32 * k] = (double)v[kr][k];
    }
    __syncthreads();
    // axis 0: item = (column c, group of G2_RB output rows); lanes run along c
    for (int i = tid; i < (G2_TY / G2_RB) * INX; i += G2_NT) {
        const int grp = i / INX, c = i - grp * INX, r0 = grp * G2_RB;
        if (c >= d.w - x0 + 2 * R || r0 >= d.h - y0) continue;       // nothing downstream reads this column / these rows
        double v[NV];
#pragma unroll
        for (int t = 0; t < NV; ++t) v[t] = tin[(r0 + t) * P + c];
#pragma unroll
        for (int u = 0; u < G2_RB; ++u) {
            double acc = v[u + R] * g.w[R];
#pragma unroll
            for (int ll = -R; ll < 0; ++ll) acc += (v[u + R + ll] + v[u + R - ll]) * g.w[ll + R];
            tmidT[c * PT + r0 + u] = (double)(float)acc;
        }
    }
    __syncthreads();
    // axis 1: item = (row r, group of G2_RB output columns); lanes run along r, the window walks tmidT's rows
    for (int i = tid; i < G2_TY * (G2_TX / G2_RB); i += G2_NT) {
        const int grp = i / G2_TY, r = i - grp * G2_TY, c0 = grp * G2_RB;
        const int gy = y0 + r;
        if (gy >= d.h || x0 + c0 >= d.w) continue;
        double v[NV];
#pragma unroll
        for (int t = 0; t < NV; ++t) v[t] = tmidT[(c0 + t) * PT + r];
        float* out = dst + (size_t)b * W * H + (size_t)gy * W + x0 + c0;
#pragma unroll
        for (int u = 0; u < G2_RB; ++u) {
            double acc = v[u + R] * g.w[R];
#pragma unroll
            for (int ll = -R; ll < 0; ++ll) acc += (v[u + R + ll] + v[u + R - ll]) * g.w[ll + R];
            if (x0 + c0 + u < d.w) out[u] = (float)acc;
        }
    }
}


// ---- streaming variant ------------------------------------------------------------------------------------------------
// One CTA = a strip of up to 256 columns x a segment of rows, walked top to bottom four output rows per step.
//   axis 0: thread = column; its window of 4 + 2R input rows lives in REGISTERS (a ring rotated at compile time), every input
//           element is loaded from HBM and converted to float64 exactly once per strip (the tiled kernel loads its halo
//           2.5 times at R = 12 and walks the window through shared memory);
//   axis 1: the four intermediate rows go through shared memory (float64 of the float32-rounded value, de-interleaved by
//           four columns so that both the column-wise stores and the 4-wide window walks are conflict-free); thread =
//           (row, group of four output columns), window of 4 + 2R values in registers.
// 'reflect' needs no halo work at the image border (the mirrored column is the in-image column); strips of a wider image
// overlap by R columns on each side.  Operation order per output is NI_Correlate1D's, as in k_gauss1d: bit-identical.

// compile-time loop: f(integral_constant<int, 0>) ... f(integral_constant<int, N - 1>) - the ring indices of the streaming
// kernel below must be constants (a run-time index into a register array is compiled into select chains)
template <int S, int N> struct GsRot {
    template <class F> static __device__ __forceinline__ void run(F& f) { f(std::integral_constant<int, S>{}); GsRot<S + 1, N>::run(f); }
};
template <int N> struct GsRot<N, N> { template <class F> static __device__ __forceinline__ void run(F&) {} };
#define GS_PL 68                      // plane stride (float64 elements) of the de-interleaved intermediate rows
template <int R>
__global__ void __launch_bounds__(256)
k_gauss2d_stream(const float* __restrict__ src, int W, int H, const int4* __restrict__ roi, const GaussW g, float* __restrict__ dst,
                 const uint8_t* __restrict__ src8, const float* __restrict__ flut, int seg_rows) {
    constexpr int NV = 4 + 2 * R, STEPS = NV / 4;
    static_assert(NV % 4 == 0, "window must be a whole number of 4-row steps");
    __shared__ double mid[2][4][4 * GS_PL];
    __shared__ short col_tab[256 + 2 * R + 4];                       // de-interleaved position of window column o_lo - R + i ('reflect' resolved)
    const int b = blockIdx.z, tid = threadIdx.x;
    const FpbDims d = fpb_dims(roi, b, W, H);
    constexpr int S_OUT = 256 - 2 * R;
    int o_lo, o_hi;
    if (d.w <= 256) { if (blockIdx.x > 0) return; o_lo = 0; o_hi = d.w; }
    else { o_lo = blockIdx.x * S_OUT; if (o_lo >= d.w) return; o_hi = min(d.w, o_lo + S_OUT); }
    const int c_lo = max(0, o_lo - R), c_hi = min(d.w, o_hi + R), ncol = c_hi - c_lo;
    const int y_lo = blockIdx.y * seg_rows;
    if (y_lo >= d.h) return;
    const int y_hi = min(d.h, y_lo + seg_rows);
    const size_t plane = (size_t)b * W * H;
    const float* p = src + plane;
    const uint8_t* p8 = src8 ? src8 + plane : nullptr;
    const float* lut = src8 ? flut + b * 256 : nullptr;
    const bool col_on = tid < ncol;
    const int gx = c_lo + (col_on ? tid : 0);
    auto load_row = [&](int y) -> float {                        // input row y of this thread's column, 'reflect' in y
        const int gy = fpb_reflect_dup(y, d.h);
        return p8 ? lut[p8[(size_t)gy * W + gx]] : p[(size_t)gy * W + gx];
    };
    // axis-1 role of this thread: row u2 of the step, output columns o_lo + 4 g2 .. + 3
    const int u2 = tid >> 6, g2 = tid & 63, oc0 = o_lo + 4 * g2;
    const bool out_on = oc0 < o_hi;
    const bool interior = oc0 - R >= 0 && oc0 + 3 + R < d.w;
    const int cc0 = oc0 - R - c_lo;                              // first window column relative to the strip (interior threads)
    const int st_idx = (tid & 3) * GS_PL + (tid >> 2);           // where this thread's column goes in a de-interleaved row

    for (int i = tid; i < o_hi - o_lo + 2 * R + 3; i += 256) {   // + 3: the window of a partial last group of four runs past o_hi + R
        const int cc = min(max(fpb_reflect_dup(o_lo - R + i, d.w) - c_lo, 0), 255);
        col_tab[i] = (short)((cc & 3) * GS_PL + (cc >> 2));
    }
    double win[NV];                                              // ring: input row (y_lo - R + r) sits in win[r % NV]
#pragma unroll
    for (int r = 0; r < NV - 4; ++r) win[r] = col_on ? (double)load_row(y_lo - R + r) : 0.0;
    // the four rows a step adds are fetched one step ahead: their HBM latency passes under the previous step's arithmetic
    float nxt[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) nxt[i] = col_on ? load_row(y_lo + R + i) : 0.0f;
    int par = 0, y = y_lo;
    auto step_fn = [&](auto sc) {
        constexpr int s = decltype(sc)::value;
        const int ys = y + 4 * s;                                // first output row of this step
        if (ys >= y_hi) return;                                  // CTA-uniform
        if (col_on) {
#pragma unroll
            for (int i = 0; i < 4; ++i) win[(4 * s + NV - 4 + i) % NV] = (double)nxt[i];
            if (ys + 4 < y_hi) {
#pragma unroll
                for (int i = 0; i < 4; ++i) nxt[i] = load_row(ys + 4 + R + i);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                constexpr int c0 = 4 * s + R;
                double acc = win[(c0 + u) % NV] * g.w[R];
#pragma unroll
                for (int ll = -R; ll < 0; ++ll) acc += (win[(c0 + u + ll) % NV] + win[(c0 + u - ll) % NV]) * g.w[ll + R];
                mid[par][u][st_idx] = (double)(float)acc;
            }
        }
        __syncthreads();
        const int gy = ys + u2;
        if (out_on && gy < y_hi) {
            const double* m = mid[par][u2];
            double v[NV];
            if (interior) {
#pragma unroll
                for (int t = 0; t < NV; ++t) { const int cc = cc0 + t; v[t] = m[(cc & 3) * GS_PL + (cc >> 2)]; }
            } else {
#pragma unroll
                for (int t = 0; t < NV; ++t) v[t] = m[col_tab[4 * g2 + t]];
            }
            float o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                double acc = v[u + R] * g.w[R];
#pragma unroll
                for (int ll = -R; ll < 0; ++ll) acc += (v[u + R + ll] + v[u + R - ll]) * g.w[ll + R];
                o[u] = (float)acc;
            }
            float* out = dst + plane + (size_t)gy * W + oc0;
            if (oc0 + 3 < o_hi && ((W & 3) == 0)) *reinterpret_cast<float4*>(out) = make_float4(o[0], o[1], o[2], o[3]);
            else {
#pragma unroll
                for (int u = 0; u < 4; ++u) if (oc0 + u < o_hi) out[u] = o[u];
            }
        }
        par ^= 1;                                                // the next step writes the other buffer (one barrier per step)
    };
    for (; y < y_hi; y += 4 * STEPS) GsRot<0, STEPS>::run(step_fn);
}

template <int R>
static void launch_gauss2d(FpbLaunch L, const float* src, int n, int W, int H, const int4* roi, const GaussW& g, float* dst,
                           const uint8_t* src8 = nullptr, const float* flut = nullptr) {
    // streaming variant (FPB_GAUSS_STREAM=1; bit-identical): strips of <= 256 columns, row segments sized so that the grid keeps
    // every SM busy.  Measured on B200, nine launches of the 1480 x 320x240 step: 12.1 ms against 8.1 ms for the tiled kernel -
    // it loads and converts every element once (1.6x fewer instructions) but its two 56-register windows allow one 8-warp CTA
    // per SM, too few warps to cover the dependent float64 add chains; the tiled kernel (24 warps per SM) stays the default.
    static const bool stream = getenv("FPB_GAUSS_STREAM") != nullptr;
    if (stream) {
        constexpr int S_OUT = 256 - 2 * R;
        const int strips = W <= 256 ? 1 : (W + S_OUT - 1) / S_OUT;
        int segs = 1;
        while ((long long)n * strips * segs < 600 && H / (segs * 2) >= 4 * (4 + 2 * R)) segs *= 2;      // a segment re-reads 2R rows
        int seg_rows = ((H + segs - 1) / segs + 3) & ~3;
        const dim3 gs(strips, (H + seg_rows - 1) / seg_rows, n);
        k_gauss2d_stream<R><<<gs, 256, 0, L.st>>>(src, W, H, roi, g, dst, src8, flut, seg_rows);
        return;
    }
    constexpr int INX = G2_TX + 2 * R, INY = G2_TY + 2 * R, P = INX + 1, PT = G2_TY + 1;
    const size_t smem = (size_t)(INY * P + INX * PT) * sizeof(double);
    FPB_OPT_IN_SMEM(k_gauss2d<R>, smem);
    const dim3 blk(32, G2_NY), gt((W + G2_TX - 1) / G2_TX, (H + G2_TY - 1) / G2_TY, n);
    k_gauss2d<R><<<gt, blk, smem, L.st>>>(src, W, H, roi, g, dst, src8, flut);
}

void fpb_gaussian_f32(FpbLaunch L, const float* src, int n, int W, int H, const int4* roi, double sigma,
                      float* tmp, float* dst) {
    if (!(sigma > 1e-15)) {          // scipy.ndimage.gaussian_filter skips axes whose sigma is <= 1e-15: the output is the input
        cudaMemcpyAsync(dst, src, (size_t)n * W * H * sizeof(float), cudaMemcpyDeviceToDevice, L.st);
        return;
    }
    const GaussW g = fpb_gauss_weights(sigma);
    const dim3 blk(32, 8);
    switch (g.r) {
        case 2:  launch_gauss2d<2>(L, src, n, W, H, roi, g, dst);  LAUNCH_COUNT(L); return;
        case 6:  launch_gauss2d<6>(L, src, n, W, H, roi, g, dst);  LAUNCH_COUNT(L); return;
        case 8:  launch_gauss2d<8>(L, src, n, W, H, roi, g, dst);  LAUNCH_COUNT(L); return;
        case 12: launch_gauss2d<12>(L, src, n, W, H, roi, g, dst); LAUNCH_COUNT(L); return;
        default: break;
    }
    const dim3 grid = px_grid(n, W, H);
    k_gauss1d<<<grid, blk, 0, L.st>>>(src, W, H, roi, g, 0, tmp);  LAUNCH_COUNT(L);
    k_gauss1d<<<grid, blk, 0, L.st>>>(tmp, W, H, roi, g, 1, dst);  LAUNCH_COUNT(L);
}

// ------------------------------------------------------------------------------------------------
// f = img/255, inverted when max > median  (:19-28).  The test of the reference,
//   mean(f[f > med]) > mean(f[f <= med]),  is true exactly when some pixel exceeds the median.
// ------------------------------------------------------------------------------------------------
__global__ void k_or_flut(const unsigned* __restrict__ hist, int W, int H, const int4* __restrict__ roi,
                          float* __restrict__ flut, int allow_invert /* invert_if_needed */) {
    __shared__ int s_inv;
    const int b = blockIdx.x, t = threadIdx.x;
    if (t == 0) {
        const FpbDims d = fpb_dims(roi, b, W, H);
        const int n = d.w * d.h;
        const unsigned* h = hist + b * 256;
        // lower middle element (n even: np.median averages it with the upper one; the comparison
        // "max > median" has the same outcome with either)
        const unsigned k = (unsigned)((n - 1) / 2) + 1u;
        unsigned run = 0; int vmed = 0, vmax = 0;
        bool found = false;
        for (int v = 0; v < 256; ++v) {
            run += h[v];
            if (!found && run >= k) { vmed = v; found = true; }
            if (h[v]) vmax = v;
        }
        s_inv = allow_invert && vmax > vmed;
    }
    __syncthreads();
    const float f = (float)t / 255.0f;
    flut[b * 256 + t] = s_inv ? (1.0f - f) : f;
}

__global__ void k_or_apply_flut(const uint8_t* __restrict__ img, int W, int H, const int4* __restrict__ roi,
                                const float* __restrict__ flut, float* __restrict__ dst) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x >= d.w || y >= d.h) return;
    const size_t o = (size_t)b * W * H + (size_t)y * W + x;
    dst[o] = flut[b * 256 + img[o]];
}

// ------------------------------------------------------------------------------------------------
// Non-uint8 input (orientation.py:21-28; never taken on the hot path, reachable through the public function):
//   f = img.astype(float32);  if f.max() > 1 or f.min() < 0:  f = (f - min) / (max - min + 1e-12)   (all float32 under NumPy 2:
//   the Python float 1e-12 is weak);  invert_if_needed:  f = 1 - f  when some pixel exceeds the median.
// One CTA per image, three passes.  "max > median" needs no selection: with cnt = number of pixels equal to the maximum and n
// pixels, the median is the maximum iff cnt >= (n + 1) / 2 (n odd) or cnt >= n / 2 + 1 (n even); for n even and cnt == n / 2
// np.median is the float32 mean of the maximum and the second-largest distinct value, which can round up to the maximum.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_or_float_prep(const float* __restrict__ src, int W, int H, const int4* __restrict__ roi, int allow_invert,
                float* __restrict__ dst) {
    __shared__ float s_lo[32], s_hi[32];
    __shared__ unsigned s_n[32];
    __shared__ float s_mn, s_mx, s_sec;
    __shared__ unsigned s_cnt;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const FpbDims d = fpb_dims(roi, b, W, H);
    const size_t base = (size_t)b * W * H;
    const int n = d.w * d.h;
    float mn = INFINITY, mx = -INFINITY;
    for (int i = t; i < n; i += 1024) {
        const int y = i / d.w, x = i - y * d.w;
        const float v = src[base + (size_t)y * W + x];
        mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    for (int off = 16; off; off >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if (lane == 0) { s_lo[wid] = mn; s_hi[wid] = mx; }
    __syncthreads();
    if (t == 0) {
        float a = s_lo[0], c = s_hi[0];
        for (int k = 1; k < 32; ++k) { a = fminf(a, s_lo[k]); c = fmaxf(c, s_hi[k]); }
        s_mn = a; s_mx = c;
    }
    __syncthreads();
    mn = s_mn; mx = s_mx;
    const bool norm = mx > 1.0f || mn < 0.0f;
    const float den = (mx - mn) + 1e-12f;
    const float top = norm ? (mx - mn) / den : mx;               // the value the maximum takes after the (monotone) mapping
    unsigned cnt = 0; float sec = -INFINITY;
    for (int i = t; i < n; i += 1024) {
        const int y = i / d.w, x = i - y * d.w;
        float v = src[base + (size_t)y * W + x];
        if (norm) v = (v - mn) / den;
        if (v == top) ++cnt; else sec = fmaxf(sec, v);
    }
    for (int off = 16; off; off >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        sec = fmaxf(sec, __shfl_xor_sync(0xffffffffu, sec, off));
    }
    __syncthreads();
    if (lane == 0) { s_n[wid] = cnt; s_hi[wid] = sec; }
    __syncthreads();
    if (t == 0) {
        unsigned a = 0; float c = -INFINITY;
        for (int k = 0; k < 32; ++k) { a += s_n[k]; c = fmaxf(c, s_hi[k]); }
        s_cnt = a; s_sec = c;
    }
    __syncthreads();
    bool inv = false;
    if (allow_invert && n > 0) {
        const unsigned c = s_cnt, un = (unsigned)n;
        if (un & 1u) inv = c < (un + 1u) / 2u;
        else if (c >= un / 2u + 1u) inv = false;
        else if (c == un / 2u) inv = top > (s_sec + top) / 2.0f;  // np.median of an even count: float32 mean of the two middle values
        else inv = true;
    }
    for (int i = t; i < n; i += 1024) {
        const int y = i / d.w, x = i - y * d.w;
        const size_t o = base + (size_t)y * W + x;
        float v = src[o];
        if (norm) v = (v - mn) / den;
        dst[o] = inv ? 1.0f - v : v;
    }
}

// cv2.Sobel(pre*255, CV_32F, ksize 3, BORDER_REFLECT_101) and the three products (:33-38)
// four horizontally adjacent pixels per thread: a 3 x 6 window feeds four outputs (4.5 loads per pixel instead of 9)
__global__ void k_or_sobel(const float* __restrict__ pre, int W, int H, const int4* __restrict__ roi,
                           float* __restrict__ gxx, float* __restrict__ gyy, float* __restrict__ gxy) {
    const int b = blockIdx.z;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x0 >= d.w || y >= d.h) return;
    const float* p = pre + (size_t)b * W * H;
    float v[3][6];
    if (x0 > 0 && y > 0 && x0 + 5 <= d.w && y + 1 < d.h) {       // interior: no border arithmetic on any of the 18 taps
        const float* q = p + (size_t)(y - 1) * W + (x0 - 1);
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < 6; ++i) v[j][i] = q[j * W + i] * 255.0f;
    } else {
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < 6; ++i)
                v[j][i] = p[(size_t)fpb_reflect101(y + j - 1, d.h) * W + fpb_reflect101(x0 + i - 1, d.w)] * 255.0f;
    }
    const size_t o = (size_t)b * W * H + (size_t)y * W + x0;
    float oxx[4], oyy[4], oxy[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float d0 = v[0][k + 2] - v[0][k], d1 = v[1][k + 2] - v[1][k], d2 = v[2][k + 2] - v[2][k];
        const float gx = (d0 + d2) + 2.0f * d1;
        const float e0 = v[2][k] - v[0][k], e1 = v[2][k + 1] - v[0][k + 1], e2 = v[2][k + 2] - v[0][k + 2];
        const float gy = (e0 + e2) + 2.0f * e1;
        oxx[k] = gx * gx; oyy[k] = gy * gy; oxy[k] = gx * gy;
    }
    if (((W & 3) == 0) && x0 + 4 <= d.w) {
        *reinterpret_cast<float4*>(gxx + o) = make_float4(oxx[0], oxx[1], oxx[2], oxx[3]);
        *reinterpret_cast<float4*>(gyy + o) = make_float4(oyy[0], oyy[1], oyy[2], oyy[3]);
        *reinterpret_cast<float4*>(gxy + o) = make_float4(oxy[0], oxy[1], oxy[2], oxy[3]);
    } else {
        for (int k = 0; k < 4 && x0 + k < d.w; ++k) { gxx[o + k] = oxx[k]; gyy[o + k] = oyy[k]; gxy[o + k] = oxy[k]; }
    }
}

// rel_raw = sqrt((Jxx-Jyy)^2 + 4*Jxy^2),  theta = 0.5*atan2(2*Jxy, (Jxx-Jyy)+1e-12) + pi/2   (:40-45), float32
// four horizontally adjacent pixels per thread (16-byte loads / stores when the row pitch allows it): the one-pixel version moved
// 20 B per pixel at 54 % of the measured HBM rate with a single load in flight per thread
__global__ void k_or_rel_theta(const float* __restrict__ jxx, const float* __restrict__ jyy, const float* __restrict__ jxy,
                               int W, int H, const int4* __restrict__ roi, float* __restrict__ rel_raw,
                               float* __restrict__ theta) {
    const int b = blockIdx.z;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x0 >= d.w || y >= d.h) return;
    const size_t o = (size_t)b * W * H + (size_t)y * W + x0;
    float xx[4], yy[4], xy[4], r[4], t[4];
    const bool vec = ((W & 3) == 0) && x0 + 4 <= d.w;
    if (vec) {
        const float4 A = *reinterpret_cast<const float4*>(jxx + o), B = *reinterpret_cast<const float4*>(jyy + o),
                     C = *reinterpret_cast<const float4*>(jxy + o);
        xx[0] = A.x; xx[1] = A.y; xx[2] = A.z; xx[3] = A.w; yy[0] = B.x; yy[1] = B.y; yy[2] = B.z; yy[3] = B.w;
        xy[0] = C.x; xy[1] = C.y; xy[2] = C.z; xy[3] = C.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) { const bool in = x0 + k < d.w; xx[k] = in ? jxx[o + k] : 0.0f; yy[k] = in ? jyy[o + k] : 0.0f; xy[k] = in ? jxy[o + k] : 0.0f; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float a = xx[k] - yy[k], c = xy[k];
        r[k] = sqrtf(a * a + 4.0f * (c * c));
        t[k] = 0.5f * atan2f(2.0f * c, a + 1e-12f) + 1.5707963267948966f;
    }
    if (vec) {
        *reinterpret_cast<float4*>(rel_raw + o) = make_float4(r[0], r[1], r[2], r[3]);
        *reinterpret_cast<float4*>(theta + o) = make_float4(t[0], t[1], t[2], t[3]);
    } else {
        for (int k = 0; k < 4 && x0 + k < d.w; ++k) { rel_raw[o + k] = r[k]; theta[o + k] = t[k]; }
    }
}

// ------------------------------------------------------------------------------------------------
// np.percentile(rel_raw, [2, 98]) per image: exact order statistics by MSB-first radix select on the
// float bit patterns (all values >= 0), then NumPy's float64 _lerp.  One block per image.
// ------------------------------------------------------------------------------------------------
// keys = float bit patterns (all >= 0, so unsigned order == float order); 11 + 11 + 10 bit digits, and BOTH quantiles
// (ranks floor((n-1)*0.02), floor((n-1)*0.98)) ride the same three passes over the data, each with its own 2048-bin
// histogram.  A fourth pass fetches "the next larger element" for whichever quantile needs rank k+1 to be a new value.
#define SEL_BINS 2048
// CLUSTER = true: the image's rows are dealt to the CTAs of a thread-block cluster (small batches of large images leave most SMs
// idle with one CTA per image); every CTA histograms its rows, the histograms are summed through distributed shared memory and
// every CTA locates the bins itself (same integers everywhere, so nothing has to be broadcast).
template <bool CLUSTER>
__global__ void __launch_bounds__(1024)
k_or_percentiles(const float* __restrict__ rel_raw, int W, int H, const int4* __restrict__ roi, double* __restrict__ pct) {
    __shared__ unsigned hist[2][SEL_BINS];
    __shared__ unsigned s_rk[2], s_prefix[2], s_below[2], s_eq[2], s_next[2];
    __shared__ int s_scanw[33];
    cg::cluster_group cl = cg::this_cluster();
    const int CL = CLUSTER ? (int)cl.num_blocks() : 1, rank = CLUSTER ? (int)cl.block_rank() : 0;
    const int b = blockIdx.x / CL, tid = threadIdx.x;
    const FpbDims d = fpb_dims(roi, b, W, H);
    const int n = d.w * d.h, w = d.w;
    const int y_first = rank * 32 + (tid >> 5), y_step = 32 * CL;
    const float* p = rel_raw + (size_t)b * W * H;
    int klo[2], khi[2]; double gam[2];
    for (int t = 0; t < 2; ++t) {
        const double q = t ? (98.0 / 100.0) : (2.0 / 100.0);
        const double vi = (double)(n - 1) * q, lo_f = floor(vi);
        klo[t] = (int)lo_f; khi[t] = klo[t] + 1;
        if (vi >= (double)(n - 1)) klo[t] = khi[t] = n - 1;
        gam[t] = vi - lo_f;
    }
    if (tid < 2) { s_rk[tid] = (unsigned)klo[tid]; s_prefix[tid] = 0u; s_below[tid] = 0u; s_next[tid] = 0xFFFFFFFFu; }
    unsigned mask = 0;
    const int shifts[3] = {21, 10, 0};
    const unsigned widths[3] = {2047u, 2047u, 1023u};
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = shifts[pass];
        const unsigned wd = widths[pass];
        for (int i = tid; i < 2 * SEL_BINS; i += 1024) (&hist[0][0])[i] = 0u;
        __syncthreads();
        const unsigned pf0 = s_prefix[0], pf1 = s_prefix[1];
        // warp = row, lane = column (no division per element); four column groups in flight per thread
        // (the trip counts are the same for all lanes of a warp: the loop body holds full-warp ballots)
        for (int y = y_first; y < d.h; y += y_step)
        for (int xb = 0; xb < w; xb += 4 * 32) {
            unsigned key[4]; bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int x = xb + (tid & 31) + 32 * u;
                ok[u] = x < w;
                key[u] = ok[u] ? __float_as_uint(p[(size_t)y * W + x]) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned binv = (key[u] >> shift) & wd;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (pass == 0 && t == 1) break;         // first digit: both quantiles see the same histogram
                    const bool take = ok[u] && ((key[u] & mask) == (t ? pf1 : pf0));
                    // warp-aggregated update: flat regions put whole warps into one bin
                    const unsigned act = __ballot_sync(0xffffffffu, take);
                    if (take) {
                        const unsigned peers = __match_any_sync(act, binv);
                        if ((int)(__ffs(peers) - 1) == (tid & 31)) atomicAdd(&hist[t][binv], __popc(peers));
                    }
                }
            }
        }
        if (CLUSTER) cl.sync(); else __syncthreads();
        // locate, for each quantile, the bin holding the wanted rank: block-wide exclusive scan, two bins per thread
        for (int t = 0; t < 2; ++t) {
            unsigned* hh = hist[pass == 0 ? 0 : t];
            unsigned h0 = hh[2 * tid], h1 = hh[2 * tid + 1];
            if (CLUSTER)
                for (int r = 1; r < CL; ++r) {
                    const unsigned* rh = cl.map_shared_rank(hh, (rank + r) % CL);
                    h0 += rh[2 * tid]; h1 += rh[2 * tid + 1];
                }
            int total;
            const unsigned ex = (unsigned)cb_block_scan_excl((int)(h0 + h1), s_scanw, &total);
            const unsigned rk = s_rk[t];
            __syncthreads();
            if (rk >= ex && rk < ex + h0) { s_rk[t] = rk - ex; s_prefix[t] |= (unsigned)(2 * tid) << shift; s_below[t] += ex; s_eq[t] = h0; }
            else if (rk >= ex + h0 && rk < ex + h0 + h1) { s_rk[t] = rk - ex - h0; s_prefix[t] |= (unsigned)(2 * tid + 1) << shift; s_below[t] += ex + h0; s_eq[t] = h1; }
        }
        mask |= wd << shift;
        if (CLUSTER) cl.sync(); else __syncthreads();      // the neighbours have read this CTA's histogram before it is cleared
    }
    // rank k+1 is the same value while it still falls among the elements <= value(k); else the next larger element
    bool need[2];
    for (int t = 0; t < 2; ++t) need[t] = (khi[t] != klo[t]) && ((unsigned)khi[t] >= s_below[t] + s_eq[t]);
    if (need[0] || need[1]) {
        const unsigned v0 = s_prefix[0], v1 = s_prefix[1];
        unsigned b0 = 0xFFFFFFFFu, b1 = 0xFFFFFFFFu;
        for (int y = y_first; y < d.h; y += y_step)
            for (int x = tid & 31; x < w; x += 32) {
                const unsigned key = __float_as_uint(p[(size_t)y * W + x]);
                if (key > v0 && key < b0) b0 = key;
                if (key > v1 && key < b1) b1 = key;
            }
        for (int off = 16; off; off >>= 1) {
            b0 = min(b0, __shfl_xor_sync(0xffffffffu, b0, off));
            b1 = min(b1, __shfl_xor_sync(0xffffffffu, b1, off));
        }
        if ((tid & 31) == 0) { atomicMin(&s_next[0], b0); atomicMin(&s_next[1], b1); }
        __syncthreads();
    }
    if (CLUSTER) cl.sync();
    if (tid < 2 && rank == 0) {
        const float a = __uint_as_float(s_prefix[tid]);
        unsigned nx = s_next[tid];
        if (CLUSTER) for (int r = 1; r < CL; ++r) nx = min(nx, *cl.map_shared_rank(&s_next[tid], r));
        const float c = need[tid] ? __uint_as_float(nx) : a;
        const float diff = c - a;                       // NumPy _lerp: float32 difference, float64 blend
        const double g = gam[tid];
        pct[b * 2 + tid] = (g >= 0.5) ? ((double)c - (double)diff * (1.0 - g)) : ((double)a + (double)diff * g);
    }
    if (CLUSTER) cl.sync();      // no CTA may exit while rank 0 can still read its shared memory
}

// ------------------------------------------------------------------------------------------------
// per 16x16 block: weighted circular mean of theta, mean reliability (:52-72).  One warp per block.
// reliability (float64) = clip((rel_raw - p2)/(p98 - p2 + 1e-12), 0, 1)
// ------------------------------------------------------------------------------------------------
__global__ void k_or_blocks(const float* __restrict__ rel_raw, const float* __restrict__ theta,
                            const uint8_t* __restrict__ mask, int W, int H, const int4* __restrict__ roi,
                            const double* __restrict__ pct, int NBX, int NBY, float* __restrict__ blk_theta,
                            float* __restrict__ blk_rel) {
    // one warp per 16x16 block, four blocks of a grid row per CTA (a 32-thread CTA caps the SM at 32 resident warps)
    const int b = blockIdx.z, bx = blockIdx.x * 4 + (threadIdx.x >> 5), by = blockIdx.y, lane = threadIdx.x & 31;
    const FpbDims d = fpb_dims(roi, b, W, H);
    const int nbx = d.w / 16, nby = d.h / 16;
    if (bx >= nbx || by >= nby) return;
    const size_t base = (size_t)b * W * H;
    const double r_lo = pct[b * 2], r_hi = pct[b * 2 + 1];
    const double den = r_hi - r_lo + 1e-12;
    double s = 0.0, c = 0.0, rs = 0.0;
    int on = 0;
    float rv[8], tv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {                             // all sixteen loads of the lane issued up front
        const int i = lane + 32 * u, yy = by * 16 + i / 16, xx = bx * 16 + (i & 15);
        const size_t o = base + (size_t)yy * W + xx;
        rv[u] = rel_raw[o]; tv[u] = theta[o];
        if (mask) on += mask[o] > 0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        double r = ((double)rv[u] - r_lo) / den;
        r = r < 0.0 ? 0.0 : (r > 1.0 ? 1.0 : r);
        const float t2 = 2.0f * tv[u];
        const double wt = r + 1e-6;
        s += wt * (double)sinf(t2);
        c += wt * (double)cosf(t2);
        rs += r;
    }
    for (int off = 16; off; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        c += __shfl_xor_sync(0xffffffffu, c, off);
        rs += __shfl_xor_sync(0xffffffffu, rs, off);
        on += __shfl_xor_sync(0xffffffffu, on, off);
    }
    if (lane == 0) {
        const size_t g = (size_t)b * NBX * NBY + (size_t)by * NBX + bx;
        // np.mean(submask > 0) < 0.3  <=>  count < 76.8
        if (mask && on < 77) { blk_theta[g] = 0.0f; blk_rel[g] = 0.0f; }
        else { blk_theta[g] = (float)(0.5 * atan2(s, c)); blk_rel[g] = (float)(rs / 256.0); }
    }
}

// The same for any block_size (orientation.py:9 `block_size`; the hot path passes 16 and takes k_or_blocks): one warp per
// bs x bs block, lanes stride over its pixels.  Skipped when np.mean(submask > 0) < 0.3.
__global__ void k_or_blocks_any(const float* __restrict__ rel_raw, const float* __restrict__ theta,
                                const uint8_t* __restrict__ mask, int W, int H, const int4* __restrict__ roi,
                                const double* __restrict__ pct, int bs, int NBX, int NBY, float* __restrict__ blk_theta,
                                float* __restrict__ blk_rel) {
    const int b = blockIdx.z, bx = blockIdx.x * 4 + (threadIdx.x >> 5), by = blockIdx.y, lane = threadIdx.x & 31;
    const FpbDims d = fpb_dims(roi, b, W, H);
    const int nbx = d.w / bs, nby = d.h / bs;
    if (bx >= nbx || by >= nby) return;
    const size_t base = (size_t)b * W * H;
    const double r_lo = pct[b * 2], r_hi = pct[b * 2 + 1];
    const double den = r_hi - r_lo + 1e-12;
    double s = 0.0, c = 0.0, rs = 0.0;
    int on = 0;
    const int area = bs * bs;
    for (int i = lane; i < area; i += 32) {
        const int yy = by * bs + i / bs, xx = bx * bs + i % bs;
        const size_t o = base + (size_t)yy * W + xx;
        double r = ((double)rel_raw[o] - r_lo) / den;
        r = r < 0.0 ? 0.0 : (r > 1.0 ? 1.0 : r);
        const float t2 = 2.0f * theta[o];
        const double wt = r + 1e-6;
        s += wt * (double)sinf(t2);
        c += wt * (double)cosf(t2);
        rs += r;
        if (mask) on += mask[o] > 0;
    }
    for (int off = 16; off; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        c += __shfl_xor_sync(0xffffffffu, c, off);
        rs += __shfl_xor_sync(0xffffffffu, rs, off);
        on += __shfl_xor_sync(0xffffffffu, on, off);
    }
    if (lane == 0) {
        const size_t g = (size_t)b * NBX * NBY + (size_t)by * NBX + bx;
        if (mask && (double)on / (double)area < 0.3) { blk_theta[g] = 0.0f; blk_rel[g] = 0.0f; }
        else { blk_theta[g] = (float)(0.5 * atan2(s, c)); blk_rel[g] = (float)(rs / (double)area); }
    }
}

// block grid: gaussian_filter(sin 2theta), gaussian_filter(cos 2theta), sigma 3, then 0.5*atan2 (:75-79).
// One CTA per image; the grid is tiny (19x13 for 320x240) and the 25-tap kernel wraps around it
// several times ('reflect' of any distance).
// BS = 16: the hot path's block size as a compile-time constant; BS = 0: `bs_rt` (fpb_orientation_ex)
template <int BS>
__global__ void __launch_bounds__(256)
k_or_grid_smooth(float* __restrict__ blk_theta, int W, int H, const int4* __restrict__ roi, int bs_rt, int NBX, int NBY,
                 GaussW g, float* __restrict__ scratch /* [n][4][NBX*NBY] */) {
    const int b = blockIdx.x;
    const int bs = BS ? BS : bs_rt;
    const FpbDims d = fpb_dims(roi, b, W, H);
    const int nbx = d.w / bs, nby = d.h / bs, N = NBX * NBY;
    float* th = blk_theta + (size_t)b * N;
    float* s0 = scratch + (size_t)b * 4 * N; float* c0 = s0 + N; float* s1 = c0 + N; float* c1 = s1 + N;
    for (int i = threadIdx.x; i < nbx * nby; i += blockDim.x) {
        const int y = i / nbx, x = i - y * nbx;
        const float t2 = 2.0f * th[y * NBX + x];
        s0[y * NBX + x] = sinf(t2); c0[y * NBX + x] = cosf(t2);
    }
    __syncthreads();
    const int r = g.r;
    for (int i = threadIdx.x; i < nbx * nby; i += blockDim.x) {       // axis 0
        const int y = i / nbx, x = i - y * nbx;
        double as = (double)s0[y * NBX + x] * g.w[r], ac = (double)c0[y * NBX + x] * g.w[r];
        for (int ll = -r; ll < 0; ++ll) {
            const int ya = fpb_reflect_dup(y + ll, nby), yb = fpb_reflect_dup(y - ll, nby);
            as += ((double)s0[ya * NBX + x] + (double)s0[yb * NBX + x]) * g.w[ll + r];
            ac += ((double)c0[ya * NBX + x] + (double)c0[yb * NBX + x]) * g.w[ll + r];
        }
        s1[y * NBX + x] = (float)as; c1[y * NBX + x] = (float)ac;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nbx * nby; i += blockDim.x) {       // axis 1
        const int y = i / nbx, x = i - y * nbx;
        double as = (double)s1[y * NBX + x] * g.w[r], ac = (double)c1[y * NBX + x] * g.w[r];
        for (int ll = -r; ll < 0; ++ll) {
            const int xa = fpb_reflect_dup(x + ll, nbx), xb = fpb_reflect_dup(x - ll, nbx);
            as += ((double)s1[y * NBX + xa] + (double)s1[y * NBX + xb]) * g.w[ll + r];
            ac += ((double)c1[y * NBX + xa] + (double)c1[y * NBX + xb]) * g.w[ll + r];
        }
        th[y * NBX + x] = 0.5f * atan2f((float)as, (float)ac);
    }
}

// cv2.resize(grid, (w,h), INTER_LINEAR) for orientation and reliability, then the wrap (:81-83)
__device__ __forceinline__ void resize_coef(int dpos, int dn, int sn, double scale, int* s0, int* s1, float* f) {
    float fx = (float)(((double)dpos + 0.5) * scale - 0.5);
    int sx = (int)floorf(fx);
    fx -= (float)sx;
    if (sx < 0) { fx = 0.0f; sx = 0; }
    if (sx >= sn - 1) { fx = 0.0f; sx = sn - 1; }
    *s0 = sx; *s1 = min(sx + 1, sn - 1); *f = fx;
}

template <int BS>
__global__ void k_or_resize(const float* __restrict__ blk_theta, const float* __restrict__ blk_rel, int W, int H,
                            const int4* __restrict__ roi, int bs_rt, int NBX, int NBY, float* __restrict__ orient_img,
                            float* __restrict__ rel_img) {
    const int b = blockIdx.z;
    const int bs = BS ? BS : bs_rt;
    const int xb = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y * blockDim.y + threadIdx.y;   // four pixels per thread
    const FpbDims d = fpb_dims(roi, b, W, H);
    const int nbx = d.w / bs, nby = d.h / bs;
    // the two float64 scale factors are per-image constants: one division each per CTA instead of per pixel
    __shared__ double s_scale[2];
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        s_scale[0] = (double)nbx / (double)max(d.w, 1); s_scale[1] = (double)nby / (double)max(d.h, 1);
    }
    __syncthreads();
    if (xb >= d.w || y >= d.h) return;
    const size_t o = (size_t)b * W * H + (size_t)y * W + xb;
    float ot[4], orl[4];
    if (nbx < 1 || nby < 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { ot[k] = 0.0f; orl[k] = 0.0f; }
    } else {
        int y0, y1; float fy;
        resize_coef(y, d.h, nby, s_scale[1], &y0, &y1, &fy);
        const float* T0 = blk_theta + (size_t)b * NBX * NBY + y0 * NBX;
        const float* T1 = blk_theta + (size_t)b * NBX * NBY + y1 * NBX;
        const float* R0 = blk_rel + (size_t)b * NBX * NBY + y0 * NBX;
        const float* R1 = blk_rel + (size_t)b * NBX * NBY + y1 * NBX;
        const float ay0 = 1.0f - fy;
        const float pi = 3.14159265358979323846f, hpi = 1.5707963267948966f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int x0, x1; float fx;
            resize_coef(min(xb + k, d.w - 1), d.w, nbx, s_scale[0], &x0, &x1, &fx);
            const float ax0 = 1.0f - fx;
            const float t_top = T0[x0] * ax0 + T0[x1] * fx;
            const float t_bot = T1[x0] * ax0 + T1[x1] * fx;
            const float r_top = R0[x0] * ax0 + R0[x1] * fx;
            const float r_bot = R1[x0] * ax0 + R1[x1] * fx;
            const float t = t_top * ay0 + t_bot * fy;
            // (t + pi/2) % pi - pi/2 with Python's sign convention.  fmod is exact and |t + pi/2| < 2 pi here, so the
            // three common cases need no library call: x in [0, pi) is its own remainder, x in [pi, 2 pi) leaves x - pi
            // (exact by Sterbenz), x in (-pi, 0) leaves x and gets the one rounded + pi
            const float xw = t + hpi;
            float m;
            if (xw >= 0.0f && xw < pi) m = xw;
            else if (xw >= pi && xw < 2.0f * pi) m = xw - pi;
            else if (xw < 0.0f && xw > -pi) m = xw + pi;
            else { m = fmodf(xw, pi); if (m != 0.0f && m < 0.0f) m += pi; }
            ot[k] = m - hpi;
            orl[k] = r_top * ay0 + r_bot * fy;
        }
    }
    if (((W & 3) == 0) && xb + 4 <= d.w) {
        *reinterpret_cast<float4*>(orient_img + o) = make_float4(ot[0], ot[1], ot[2], ot[3]);
        *reinterpret_cast<float4*>(rel_img + o) = make_float4(orl[0], orl[1], orl[2], orl[3]);
    } else {
        for (int k = 0; k < 4 && xb + k < d.w; ++k) { orient_img[o + k] = ot[k]; rel_img[o + k] = orl[k]; }
    }
}

// `prm` == nullptr: the values on the reference's hot path (block 16, sigmas 3.0 / 3.0, invert_if_needed;
// fingerprint_preprocess.py:192-195, post_processing.py:93).  Other values (orientation.py:9-14 as a public function) take the
// same kernels with the generic Gaussian for radii outside {2, 6, 8, 12} and k_or_blocks_any for block sizes other than 16;
// the caller sizes orient_blocks / ws.blk_rel / ws.blk_scratch for (W / block_size) * (H / block_size) entries per image.
void fpb_orientation_core(FpbLaunch L, const uint8_t* img, const uint8_t* mask, int n, int W, int H,
                          const int4* roi, FpbOrientWs ws, float* orient_blocks, float* orient_img, float* rel_img,
                          const FpbOrientPrm* prm, const float* img_f32) {
    const dim3 blk(32, 8), grid = px_grid(n, W, H);
    const int bs = prm ? prm->block_size : 16;
    const double sig_t = prm ? prm->smooth_sigma : 3.0, sig_b = prm ? prm->smooth_orientation_sigma : 3.0;
    const double sig_pre = sig_t / 2.0 > 0.5 ? sig_t / 2.0 : 0.5;                                     // max(0.5, smooth_sigma / 2)  (:30)
    const int NBX = W / bs, NBY = H / bs;
    if (img_f32) {              // non-uint8 input (orientation.py:21-24): f formed from the float plane; must not alias ws.t0 .. t4
        k_or_float_prep<<<n, 1024, 0, L.st>>>(img_f32, W, H, roi, prm ? prm->invert_if_needed : 1, ws.t0);         LAUNCH_COUNT(L);
        fpb_gaussian_f32(L, ws.t0, n, W, H, roi, sig_pre, ws.t1, ws.t2);
    } else {
        fpb_hist256(L, img, n, W, H, roi, ws.hist);
        k_or_flut<<<n, 256, 0, L.st>>>(ws.hist, W, H, roi, ws.flut, prm ? prm->invert_if_needed : 1);              LAUNCH_COUNT(L);
        // pre = gaussian_filter(f, 1.5) with f = flut[img] formed while the tile is loaded (no float plane for f)
        const GaussW g15 = fpb_gauss_weights(sig_pre);
        if (g15.r == 6) { launch_gauss2d<6>(L, ws.t0, n, W, H, roi, g15, ws.t2, img, ws.flut); LAUNCH_COUNT(L); }
        else {
            k_or_apply_flut<<<grid, blk, 0, L.st>>>(img, W, H, roi, ws.flut, ws.t0);                                LAUNCH_COUNT(L);
            fpb_gaussian_f32(L, ws.t0, n, W, H, roi, sig_pre, ws.t1, ws.t2);
        }
    }
    const dim3 grid4((W + 127) / 128, (H + 7) / 8, n);       // kernels that take four pixels per thread
    k_or_sobel<<<grid4, blk, 0, L.st>>>(ws.t2, W, H, roi, ws.t0, ws.t1, ws.t3);                       LAUNCH_COUNT(L);   // gxx,gyy,gxy
    fpb_gaussian_f32(L, ws.t0, n, W, H, roi, sig_t, ws.t4, ws.t2);           // jxx = t2
    fpb_gaussian_f32(L, ws.t1, n, W, H, roi, sig_t, ws.t4, ws.t0);           // jyy = t0
    fpb_gaussian_f32(L, ws.t3, n, W, H, roi, sig_t, ws.t4, ws.t1);           // jxy = t1
    k_or_rel_theta<<<grid4, blk, 0, L.st>>>(ws.t2, ws.t0, ws.t1, W, H, roi, ws.t3, ws.t4);           LAUNCH_COUNT(L);   // rel_raw=t3, theta=t4
    {   // one CTA per image, or - small batches of large images - a cluster of 2 / 4 / 8 CTAs per image (about two CTAs per SM)
        int clsz = 1;
        static const bool no_cl = getenv("FPB_NO_CLUSTER") != nullptr;
        if (!no_cl && (size_t)W * H >= 256 * 256) while (clsz < 8 && (size_t)n * clsz * 2 <= 296) clsz *= 2;
        bool done = false;
        if (clsz > 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(n * clsz)); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 0; cfg.stream = L.st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)clsz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            const float* rel = ws.t3; double* pct = ws.pct;
            done = cudaLaunchKernelEx(&cfg, k_or_percentiles<true>, rel, W, H, roi, pct) == cudaSuccess;
            if (!done) (void)cudaGetLastError();
        }
        if (!done) k_or_percentiles<false><<<n, 1024, 0, L.st>>>(ws.t3, W, H, roi, ws.pct);
        LAUNCH_COUNT(L);
    }
    if (NBX > 0 && NBY > 0) {
        float* blk_rel = ws.blk_rel;                                // [n][NBX*NBY]
        float* scratch = ws.blk_scratch;                            // [n][4][NBX*NBY]
        cudaMemsetAsync(orient_blocks, 0, (size_t)n * NBX * NBY * sizeof(float), L.st);
        cudaMemsetAsync(blk_rel, 0, (size_t)n * NBX * NBY * sizeof(float), L.st);
        dim3 gb((NBX + 3) / 4, NBY, n);
        if (bs == 16) { k_or_blocks<<<gb, 128, 0, L.st>>>(ws.t3, ws.t4, mask, W, H, roi, ws.pct, NBX, NBY, orient_blocks, blk_rel); LAUNCH_COUNT(L); }
        else { k_or_blocks_any<<<gb, 128, 0, L.st>>>(ws.t3, ws.t4, mask, W, H, roi, ws.pct, bs, NBX, NBY, orient_blocks, blk_rel); LAUNCH_COUNT(L); }
        GaussW gb_w;
        if (sig_b > 1e-15) gb_w = fpb_gauss_weights(sig_b);
        else { gb_w.r = 0; gb_w.w[0] = 1.0; }                      // SciPy leaves an axis with sigma <= 1e-15 unfiltered
        if (bs == 16) {
            k_or_grid_smooth<16><<<n, 256, 0, L.st>>>(orient_blocks, W, H, roi, bs, NBX, NBY, gb_w, scratch);               LAUNCH_COUNT(L);
            k_or_resize<16><<<grid4, blk, 0, L.st>>>(orient_blocks, blk_rel, W, H, roi, bs, NBX, NBY, orient_img, rel_img); LAUNCH_COUNT(L);
        } else {
            k_or_grid_smooth<0><<<n, 256, 0, L.st>>>(orient_blocks, W, H, roi, bs, NBX, NBY, gb_w, scratch);                LAUNCH_COUNT(L);
            k_or_resize<0><<<grid4, blk, 0, L.st>>>(orient_blocks, blk_rel, W, H, roi, bs, NBX, NBY, orient_img, rel_img);  LAUNCH_COUNT(L);
        }
    }
}
