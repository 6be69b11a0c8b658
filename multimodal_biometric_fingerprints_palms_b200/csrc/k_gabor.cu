// EXTENSION (rows G1/G2 of SURVEY.md 8(a)): per-block ridge frequency + oriented Gabor enhancement.
//
// NOT IN THE REFERENCE: `grep -ri "gabor|frequen"` over the reference finds nothing, so there is no oracle in the
// reference and parity is UNPINNED.  BASELINE.json's north_star names these two kernels, and the reference's driver
// reads a result key "enhanced" that its pipeline never produces (run_preprocessing.py:133), so they are provided as an
// opt-in extension under that key; `oracle/gabor_ext.py` is the NumPy statement of exactly this arithmetic and the
// tests compare against it.  The default pipeline (and every parity claim) is untouched when the extension is off.
//
// Algorithm (Hong, Wan & Jain 1998, on the K5 block grid):
//   k_ridge_freq  one warp per 16x16 block: x-signature over an oriented 32 (across ridges) x 16 (along ridges) window
//                 centred on the block, lane k = position k across the ridges; [1 2 1]/4 smoothing and peak detection
//                 with warp shuffles / ballot; period = mean peak spacing, valid in [min_period, max_period].
//   k_freq_fill   one CTA per image on the block grid in shared memory: invalid blocks filled from valid 8-neighbours
//                 (up to 8 sweeps), the rest from the global mean, then a 3x3 mean.
//   k_gabor       32x32-pixel tile + halo staged in shared memory; for each of its four 16x16 blocks the block's filter
//                 (orientation bin of the block's theta, integer period bin of its frequency) is staged from the
//                 L2-resident bank into shared memory and applied as a dense (2R+1)^2 stencil with explicit FMAs.
//                 No tensor cores: the filter changes from block to block, it is not a dense contraction.
#include <math.h>
#include <vector>

#include "fpb_kernels.h"

#define GB_PI 3.14159265358979323846

__device__ __forceinline__ int gb_reflect101(int i, int n) { return fpb_reflect101(i, n); }

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_ridge_freq(const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask, int W,
                                                   int H, const int4* __restrict__ roi, const float* __restrict__ blk_theta,
                                                   int NBX, int NBY, float pmin, float pmax, float min_amp,
                                                   float* __restrict__ blk_freq) {
    const int b = blockIdx.z, bx = blockIdx.x, by = blockIdx.y, k = threadIdx.x;
    const FpbDims d = fpb_dims(roi, b, W, H);
    const int nbx = d.w / 16, nby = d.h / 16;
    if (bx >= nbx || by >= nby) return;
    const size_t base = (size_t)b * W * H;
    const size_t g = (size_t)b * NBX * NBY + (size_t)by * NBX + bx;
    const double th = (double)blk_theta[g];
    const double c = cos(th), s = sin(th);                     // ridge direction (c, s), normal (-s, c)
    const double cx = bx * 16 + 7.5, cy = by * 16 + 7.5;
    const double off = (double)k - 15.5;
    float acc = 0.0f;
    for (int t = 0; t < 16; ++t) {
        const double along = (double)t - 7.5;
        const double x = cx + along * c + off * (-s), y = cy + along * s + off * c;
        int xi = (int)floor(x + 0.5), yi = (int)floor(y + 0.5);
        xi = xi < 0 ? 0 : (xi >= d.w ? d.w - 1 : xi);
        yi = yi < 0 ? 0 : (yi >= d.h ? d.h - 1 : yi);
        acc += (float)img[base + (size_t)yi * W + xi];
    }
    const float X = acc * (1.0f / 16.0f);
    float lft = __shfl_up_sync(0xffffffffu, X, 1), rgt = __shfl_down_sync(0xffffffffu, X, 1);
    if (k == 0) lft = X;
    if (k == 31) rgt = X;
    const float Y = (lft + 2.0f * X + rgt) * 0.25f;
    float yl = __shfl_up_sync(0xffffffffu, Y, 1), yr = __shfl_down_sync(0xffffffffu, Y, 1);
    const bool peak = k >= 1 && k <= 30 && Y > yl && Y >= yr;
    const unsigned pm = __ballot_sync(0xffffffffu, peak);
    float mx = Y, mn = Y;
    int on = 0;
    if (mask)
        for (int i = k; i < 256; i += 32) on += mask[base + (size_t)(by * 16 + i / 16) * W + bx * 16 + (i & 15)] > 0;
    for (int o = 16; o; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        on += __shfl_xor_sync(0xffffffffu, on, o);
    }
    if (k == 0) {
        float f = 0.0f;
        const int np = __popc(pm);
        if (np >= 2 && (mx - mn) >= min_amp && (!mask || on >= 77)) {
            const int first = __ffs(pm) - 1, last = 31 - __clz(pm);
            const float period = (float)(last - first) / (float)(np - 1);
            if (period >= pmin && period <= pmax) f = 1.0f / period;
        }
        blk_freq[g] = f;
    }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_freq_fill(float* __restrict__ blk_freq, int W, int H, const int4* __restrict__ roi,
                                                   int NBX, int NBY, float default_freq) {
    extern __shared__ float sm[];
    const int b = blockIdx.x;
    const FpbDims d = fpb_dims(roi, b, W, H);
    const int nbx = d.w / 16, nby = d.h / 16, N = nbx * nby;
    if (N == 0) return;
    float* A = sm;
    float* B = sm + NBX * NBY;
    float* F = blk_freq + (size_t)b * NBX * NBY;
    __shared__ float s_sum;
    __shared__ int s_cnt, s_changed;
    for (int i = threadIdx.x; i < N; i += blockDim.x) A[i] = F[(i / nbx) * NBX + (i % nbx)];
    __syncthreads();
    for (int sweep = 0; sweep < 8; ++sweep) {
        if (threadIdx.x == 0) s_changed = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            float v = A[i];
            if (v == 0.0f) {
                const int y = i / nbx, x = i % nbx;
                float sum = 0.0f; int cnt = 0;
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int yy = y + dy, xx = x + dx;
                        if (yy < 0 || yy >= nby || xx < 0 || xx >= nbx) continue;
                        const float q = A[yy * nbx + xx];
                        if (q > 0.0f) { sum += q; ++cnt; }
                    }
                if (cnt) { v = sum / (float)cnt; s_changed = 1; }
            }
            B[i] = v;
        }
        __syncthreads();
        float* t = A; A = B; B = t;
        if (!s_changed) break;
        __syncthreads();
    }
    // remaining holes: mean of the valid blocks (sequential, fixed order), or the default
    if (threadIdx.x == 0) {
        float sum = 0.0f; int cnt = 0;
        for (int i = 0; i < N; ++i) if (A[i] > 0.0f) { sum += A[i]; ++cnt; }
        s_sum = cnt ? sum / (float)cnt : default_freq; s_cnt = cnt;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) if (A[i] == 0.0f) A[i] = s_sum;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {              // 3x3 mean over the in-range neighbours
        const int y = i / nbx, x = i % nbx;
        float sum = 0.0f; int cnt = 0;
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                const int yy = y + dy, xx = x + dx;
                if (yy < 0 || yy >= nby || xx < 0 || xx >= nbx) continue;
                sum += A[yy * nbx + xx]; ++cnt;
            }
        F[y * NBX + x] = sum / (float)cnt;
    }
}

// ---------------------------------------------------------------------------------------------------------------
struct GaborBankDev {
    const float* taps;      // all filters back to back
    const int* offset;      // [n_period][n_orient] start of each filter in `taps`
    const int* radius;      // [n_period]
    int n_orient, pmin, pmax, rmax;
    int r_lo, r_hi;          // this launch handles the tiles whose largest filter radius is in (r_lo, r_hi]
};

__device__ __forceinline__ void gb_pick(const GaborBankDev& bank, float theta, float freq, int& fi, int& R) {
    const double step = GB_PI / bank.n_orient;
    double t = fmod((double)theta, GB_PI);
    if (t < 0.0) t += GB_PI;
    int oi = (int)floor(t / step + 0.5);
    if (oi >= bank.n_orient) oi -= bank.n_orient;
    int p = (int)floorf(1.0f / freq + 0.5f);
    p = p < bank.pmin ? bank.pmin : (p > bank.pmax ? bank.pmax : p);
    fi = (p - bank.pmin) * bank.n_orient + oi;
    R = bank.radius[p - bank.pmin];
}

__global__ void __launch_bounds__(256) k_gabor(const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask, int W, int H,
                                               const int4* __restrict__ roi, const float* __restrict__ blk_theta,
                                               const float* __restrict__ blk_freq, int NBX, int NBY, GaborBankDev bank,
                                               float* __restrict__ response, uint8_t* __restrict__ enhanced) {
    extern __shared__ float sm[];
    const int b = blockIdx.z;
    const FpbDims d = fpb_dims(roi, b, W, H);
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    if (x0 >= d.w || y0 >= d.h) return;
    const int nbx = d.w / 16, nby = d.h / 16;
    const size_t base = (size_t)b * W * H;
    const int tid = threadIdx.x;
    if (nbx == 0 || nby == 0) {                                   // smaller than one block: nothing to estimate from
        for (int i = tid; i < 1024; i += 256) {
            const int x = x0 + (i & 31), y = y0 + (i >> 5);
            if (x < d.w && y < d.h) { enhanced[base + (size_t)y * W + x] = 255; if (response) response[base + (size_t)y * W + x] = 0.0f; }
        }
        return;
    }
    int fi[4], Rq[4], Rt = 0;
    const float* th = blk_theta + (size_t)b * NBX * NBY;
    const float* fr = blk_freq + (size_t)b * NBX * NBY;
    for (int q = 0; q < 4; ++q) {
        const int qx = x0 + (q & 1) * 16, qy = y0 + (q >> 1) * 16;
        fi[q] = -1; Rq[q] = 0;
        if (qx >= d.w || qy >= d.h) continue;
        const int bx = min(qx / 16, nbx - 1), by = min(qy / 16, nby - 1);
        gb_pick(bank, th[by * NBX + bx], fr[by * NBX + bx], fi[q], Rq[q]);
        Rt = max(Rt, Rq[q]);
    }
    if (Rt <= bank.r_lo || Rt > bank.r_hi) return;                // the other launch (other shared-memory size) has it
    const int TW = ((32 + 2 * Rt + 7) & ~7) + 4;                  // row pitch = 4 (mod 8): the two half-warps hit disjoint banks
    const int TH = 32 + 2 * Rt;
    float* tile = sm;
    const int THm = 32 + 2 * bank.r_hi;
    float* filt = sm + (THm + 4) * (((THm + 7) & ~7) + 4);        // four filters, (Kmax+6) x Kmax floats apart
    const int fstride = (2 * bank.r_hi + 7) * (2 * bank.r_hi + 1);
    for (int i = tid; i < (TH + 4) * TW; i += 256) {              // four spare rows: read against zero taps only
        const int ty = i / TW, tx = i - ty * TW;
        float v = 0.0f;
        if (tx < TH && ty < TH) {
            const int x = gb_reflect101(x0 - Rt + tx, d.w), y = gb_reflect101(y0 - Rt + ty, d.h);
            v = (float)img[base + (size_t)y * W + x];
        }
        tile[i] = v;
    }
    for (int q = 0; q < 4; ++q) {                                 // filter rows padded with zero rows to a multiple of 4
        if (fi[q] < 0) continue;                                  // steps that also covers the 3 extra rows of a strip
        const int K = 2 * Rq[q] + 1, Kp = (K + 6) & ~3;
        const float* src = bank.taps + bank.offset[fi[q]];
        float* dst = filt + q * fstride;
        for (int i = tid; i < Kp * K; i += 256) dst[i] = i < K * K ? src[i] : 0.0f;
    }
    __syncthreads();
    // 64 threads per 16x16 block (two warps, so a warp never mixes filters); each thread owns a 1 x 4 column strip and
    // slides the filter column over it: per step one tile value and one tap from shared memory feed four FMAs
    const int q = tid >> 6, t = tid & 63;
    if (fi[q] < 0) return;
    const int R = Rq[q], K = 2 * R + 1, Kp = (K + 6) & ~3;
    const int lx = (q & 1) * 16 + (t & 15), ly = (q >> 1) * 16 + (t >> 4) * 4;
    const float* f = filt + q * fstride;
    if (mask) {                                                   // a strip entirely outside the hull: nothing to filter
        bool any = false;
        for (int j = 0; j < 4; ++j) {
            const int y = y0 + ly + j;
            if (x0 + lx < d.w && y < d.h) any |= mask[base + (size_t)y * W + x0 + lx] != 0;
        }
        if (!__any_sync(__activemask(), any)) {
            for (int j = 0; j < 4; ++j) {
                const int y = y0 + ly + j;
                if (x0 + lx < d.w && y < d.h) { enhanced[base + (size_t)y * W + x0 + lx] = 255; if (response) response[base + (size_t)y * W + x0 + lx] = 0.0f; }
            }
            return;
        }
    }
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    for (int u = 0; u < K; ++u) {
        const float* tp = tile + (ly + Rt - R) * TW + (lx + Rt - R + u);
        const float* fp = f + u;
        float f1 = 0.0f, f2 = 0.0f, f3 = 0.0f;                    // taps of rows v-1, v-2, v-3
        int v = 0;
        for (; v < Kp; v += 4, tp += 4 * TW, fp += 4 * K) {       // four steps at a time: no register shuffling
            const float t0 = tp[0], t1 = tp[TW], t2 = tp[2 * TW], t3 = tp[3 * TW];
            const float g0 = fp[0], g1 = fp[K], g2 = fp[2 * K], g3 = fp[3 * K];
            a0 = __fmaf_rn(g0, t0, a0); a1 = __fmaf_rn(f1, t0, a1); a2 = __fmaf_rn(f2, t0, a2); a3 = __fmaf_rn(f3, t0, a3);
            a0 = __fmaf_rn(g1, t1, a0); a1 = __fmaf_rn(g0, t1, a1); a2 = __fmaf_rn(f1, t1, a2); a3 = __fmaf_rn(f2, t1, a3);
            a0 = __fmaf_rn(g2, t2, a0); a1 = __fmaf_rn(g1, t2, a1); a2 = __fmaf_rn(g0, t2, a2); a3 = __fmaf_rn(f1, t2, a3);
            a0 = __fmaf_rn(g3, t3, a0); a1 = __fmaf_rn(g2, t3, a1); a2 = __fmaf_rn(g1, t3, a2); a3 = __fmaf_rn(g0, t3, a3);
            f1 = g3; f2 = g2; f3 = g1;
        }
    }
    const float accs[4] = {a0, a1, a2, a3};
    const int x = x0 + lx;
    if (x >= d.w) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int y = y0 + ly + j;
        if (y >= d.h) break;
        const size_t o = base + (size_t)y * W + x;
        if (mask && mask[o] == 0) { enhanced[o] = 255; if (response) response[o] = 0.0f; continue; }
        if (response) response[o] = accs[j];
        float e = rintf(128.0f + 2.0f * accs[j]);
        e = e < 0.0f ? 0.0f : (e > 255.0f ? 255.0f : e);
        enhanced[o] = (uint8_t)e;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host: bank construction (float64 math, float32 taps) and launchers
// ---------------------------------------------------------------------------------------------------------------
int fpb_gabor_build_bank(const FpbGaborParams& p, std::vector<float>& taps, std::vector<int>& offset, std::vector<int>& radius) {
    taps.clear(); offset.clear(); radius.clear();
    int rmax = 0;
    for (int per = p.min_period; per <= p.max_period; ++per) {
        const double sigma = p.sigma_factor * per;
        int R = (int)ceil(p.radius_factor * sigma);
        if (R > FPB_GABOR_RMAX) R = FPB_GABOR_RMAX;
        if (R < 1) R = 1;
        radius.push_back(R);
        rmax = R > rmax ? R : rmax;
        const int K = 2 * R + 1;
        for (int oi = 0; oi < p.n_orient; ++oi) {
            const double phi = oi * GB_PI / p.n_orient;
            const double cs = cos(phi), sn = sin(phi);
            std::vector<double> g((size_t)K * K);
            double mean = 0.0;
            for (int v = -R; v <= R; ++v)
                for (int u = -R; u <= R; ++u) {
                    const double xn = -u * sn + v * cs;                     // coordinate across the ridges
                    const double val = exp(-(double)(u * u + v * v) / (2.0 * sigma * sigma)) * cos(2.0 * GB_PI * xn / per);
                    g[(size_t)(v + R) * K + (u + R)] = val;
                    mean += val;
                }
            mean /= (double)(K * K);
            double l1 = 0.0;
            for (double& x : g) { x -= mean; l1 += fabs(x); }
            offset.push_back((int)taps.size());
            for (double x : g) taps.push_back((float)(x / l1));
        }
    }
    return rmax;
}

void fpb_ridge_frequency(FpbLaunch L, const uint8_t* img, const uint8_t* mask, int n, int W, int H, const int4* roi,
                         const float* blk_theta, const FpbGaborParams& p, float* blk_freq) {
    const int NBX = W / 16, NBY = H / 16;
    if (NBX == 0 || NBY == 0) return;
    k_ridge_freq<<<dim3(NBX, NBY, n), 32, 0, L.st>>>(img, mask, W, H, roi, blk_theta, NBX, NBY, (float)p.min_period,
                                                     (float)p.max_period, (float)p.min_amplitude, blk_freq);
    LAUNCH_COUNT(L);
    k_freq_fill<<<n, 256, 2 * NBX * NBY * sizeof(float), L.st>>>(blk_freq, W, H, roi, NBX, NBY, (float)(1.0 / p.default_period));
    LAUNCH_COUNT(L);
}

void fpb_gabor_apply(FpbLaunch L, const uint8_t* img, const uint8_t* mask, int n, int W, int H, const int4* roi,
                     const float* blk_theta, const float* blk_freq, const FpbGaborBank& bank, float* response, uint8_t* enhanced) {
    const int NBX = W / 16, NBY = H / 16;
    FPB_OPT_IN_SMEM(k_gabor, 200 * 1024);
    // Shared memory is sized by the largest radius a launch accepts, and occupancy follows it; the block periods are
    // only known on the device, so there are two launches: one provisioned for radius <= 14 (periods up to 12 with the
    // default factors; 30 KB, 7 CTAs per SM) and one for the rest of the bank.  A tile is taken by exactly one of them.
    const int r_fast = 14;
    const int bounds[3] = {0, bank.rmax < r_fast ? bank.rmax : r_fast, bank.rmax};
    for (int pass = 0; pass < 2; ++pass) {
        if (bounds[pass + 1] <= bounds[pass]) continue;
        const int r_hi = bounds[pass + 1];
        const int THm = 32 + 2 * r_hi, TWm = ((THm + 7) & ~7) + 4, Km = 2 * r_hi + 1;
        GaborBankDev dv = {bank.d_taps, bank.d_offset, bank.d_radius, bank.n_orient, bank.pmin, bank.pmax, bank.rmax, bounds[pass], r_hi};
        const size_t smem = sizeof(float) * ((size_t)(THm + 4) * TWm + 4 * (size_t)(Km + 6) * Km);
        k_gabor<<<dim3((W + 31) / 32, (H + 31) / 32, n), 256, smem, L.st>>>(img, mask, W, H, roi, blk_theta, blk_freq, NBX, NBY, dv,
                                                                            response, enhanced);
        LAUNCH_COUNT(L);
    }
}
