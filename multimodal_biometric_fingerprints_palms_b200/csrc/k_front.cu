// K1 (percentile stretch + CLAHE), generic CLAHE, K2 (non-local means + fixed-point Gaussian).
//
// Reference behaviour replaced (paths under /root/reference/src/preprocessing):
//   normalize_image   fingerprint_preprocess.py:13-29   np.percentile stretch, cv2 CLAHE(2.5, 8x8)
//   denoise_image     fingerprint_preprocess.py:34-38   cv2.fastNlMeansDenoising(h=10,7,21), GaussianBlur 3x3 0.6
//   CLAHE(2.0)/(2.5)  fingerprint_preprocess.py:46-47, 97-98
//   GaussianBlur 5x5  fingerprint_preprocess.py:99
// All of it is integer / fixed-point / literal-float32 arithmetic and is bit-exact against OpenCV
// (semantics pinned in oracle/stages.py).
#include "fpb_kernels.h"
#include "hd_scalar.h"
#include <cuda.h>
#include <string.h>
#include <stdlib.h>          // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)


// ------------------------------------------------------------------------------------------------
// per-image 256-bin histogram
// ------------------------------------------------------------------------------------------------
__global__ void k_hist256(const uint8_t* __restrict__ src, int W, int H, const int4* __restrict__ roi,
                          unsigned* __restrict__ hist) {
    __shared__ unsigned sh[256];
    const int b = blockIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    sh[threadIdx.x] = 0;
    __syncthreads();
    const uint8_t* p = src + (size_t)b * W * H;
    const int rows_per = (d.h + gridDim.x - 1) / gridDim.x;
    const int y0 = blockIdx.x * rows_per, y1 = min(d.h, y0 + rows_per);
    for (int x = threadIdx.x; x < d.w; x += blockDim.x)
        for (int y = y0; y < y1; y += 4) {                   // four row loads in flight per trip
            int v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = (y + u < y1) ? (int)p[(size_t)(y + u) * W + x] : -1;
#pragma unroll
            for (int u = 0; u < 4; ++u) if (v[u] >= 0) atomicAdd(&sh[v[u]], 1u);
        }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[b * 256 + threadIdx.x], sh[threadIdx.x]);
}

void fpb_hist256(FpbLaunch L, const uint8_t* src, int n, int W, int H, const int4* roi, unsigned* hist) {
    cudaMemsetAsync(hist, 0, (size_t)n * 256 * sizeof(unsigned), L.st);
    dim3 grid(8, n);
    k_hist256<<<grid, 256, 0, L.st>>>(src, W, H, roi, hist);
    LAUNCH_COUNT(L);
}

// ------------------------------------------------------------------------------------------------
// K1a: 256-entry stretch map from the histogram (one block per image)
// ------------------------------------------------------------------------------------------------
__global__ void k_stretch_lut(const unsigned* __restrict__ hist, int npix, uint8_t* __restrict__ lut) {
    __shared__ unsigned cum[256];
    __shared__ float plo, phi;
    const int b = blockIdx.x;
    cum[threadIdx.x] = hist[b * 256 + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned run = 0;
        for (int i = 0; i < 256; ++i) { run += cum[i]; cum[i] = run; }
        plo = fpb_percentile_u8_unit(cum, npix, 0.5f);
        phi = fpb_percentile_u8_unit(cum, npix, 99.5f);
    }
    __syncthreads();
    lut[b * 256 + threadIdx.x] = fpb_stretch_value(threadIdx.x, plo, phi);
}

void fpb_stretch_lut(FpbLaunch L, const unsigned* hist, int n, int W, int H, uint8_t* lut) {
    k_stretch_lut<<<n, 256, 0, L.st>>>(hist, W * H, lut);
    LAUNCH_COUNT(L);
}

// ------------------------------------------------------------------------------------------------
// CLAHE (OpenCV imgproc/clahe.cpp, 8x8 tiles): per-tile clipped-histogram LUT, then bilinear blend
// ------------------------------------------------------------------------------------------------
struct ClaheGeom { int w, h, tw, th; };

__device__ __forceinline__ ClaheGeom clahe_geom(FpbDims d) {
    ClaheGeom g; g.w = d.w; g.h = d.h;
    int ew = d.w, eh = d.h;
    if ((d.w % 8) || (d.h % 8)) { ew = d.w + 8 - (d.w % 8); eh = d.h + 8 - (d.h % 8); }   // both axes padded
    g.tw = ew / 8; g.th = eh / 8;
    return g;
}

// One warp per tile, the eight tiles of a tile row per CTA (a CTA per tile made 64 tiny CTAs per image and the launch
// was bound by CTA turnover).  Lane l owns bins 8l .. 8l+7: clip, redistribution and the inclusive scan are lane-local
// work plus one warp reduction and one warp scan.
__global__ void __launch_bounds__(256)
k_clahe_tiles(const uint8_t* __restrict__ src, const uint8_t* __restrict__ premap, int W, int H,
              const int4* __restrict__ roi, double clip, uint8_t* __restrict__ tilelut) {
    __shared__ unsigned hist[8][256];
    const int b = blockIdx.y, ty = blockIdx.x, tx = threadIdx.x >> 5, lane = threadIdx.x & 31, tile = ty * 8 + tx;
    const ClaheGeom g = clahe_geom(fpb_dims(roi, b, W, H));
    uint8_t* out = tilelut + ((size_t)b * 64 + tile) * 256 + lane * 8;
    if (g.w < 2 || g.h < 2) {
#pragma unroll
        for (int k = 0; k < 8; ++k) out[k] = (uint8_t)(lane * 8 + k);
        return;
    }
    const uint8_t* p = src + (size_t)b * W * H;
    const uint8_t* pm = premap ? premap + b * 256 : nullptr;
    unsigned* hh = hist[tx];
#pragma unroll
    for (int k = 0; k < 8; ++k) hh[lane + 32 * k] = 0;
    __syncwarp();
    const int area = g.tw * g.th;
    // lanes along the tile row: no division per pixel; only the padded right / bottom tiles reflect
    for (int c = lane; c < g.tw; c += 32) {
        const int ex = tx * g.tw + c;
        const int sx = ex < g.w ? ex : fpb_reflect101(ex, g.w);
        for (int r0 = 0; r0 < g.th; r0 += 4) {                        // four row loads in flight per trip
            int v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int ey = ty * g.th + r0 + u;
                v[u] = -1;
                if (r0 + u < g.th) v[u] = p[(size_t)(ey < g.h ? ey : fpb_reflect101(ey, g.h)) * W + sx];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (v[u] >= 0) atomicAdd(&hh[pm ? pm[v[u]] : v[u]], 1u);
        }
    }
    __syncwarp();
    int clip_limit = (int)(clip * (double)area / 256.0);
    if (clip_limit < 1) clip_limit = 1;
    int hv[8], clipped = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        hv[k] = (int)hh[lane * 8 + k];
        if (hv[k] > clip_limit) { clipped += hv[k] - clip_limit; hv[k] = clip_limit; }
    }
    for (int off = 16; off; off >>= 1) clipped += __shfl_xor_sync(0xffffffffu, clipped, off);
    const int batch = clipped / 256, resid = clipped - batch * 256;
    const int step = resid ? max(256 / resid, 1) : 1;
    unsigned run = 0, pre[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int t = lane * 8 + k;
        hv[k] += batch;
        if (resid != 0 && t % step == 0 && t / step < resid) hv[k] += 1;
        run += (unsigned)hv[k]; pre[k] = run;
    }
    unsigned inc = run;                                   // inclusive scan of the lane totals
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += o; }
    const unsigned basev = inc - run;
    const float lut_scale = 255.0f / (float)area;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float v = rintf((float)(basev + pre[k]) * lut_scale);
        v = fminf(fmaxf(v, 0.0f), 255.0f);
        out[k] = (uint8_t)v;
    }
}

// four horizontally adjacent pixels per thread: the tile geometry, the row's vertical coefficients and the two LUT row
// bases are computed once for the four; pixels go in and out as one 32-bit word when the row pitch allows it
__global__ void k_clahe_interp(const uint8_t* __restrict__ src, const uint8_t* __restrict__ premap, int W, int H,
                               const int4* __restrict__ roi, const uint8_t* __restrict__ tilelut,
                               uint8_t* __restrict__ dst) {
    const int b = blockIdx.z;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y * blockDim.y + threadIdx.y;
    const ClaheGeom g = clahe_geom(fpb_dims(roi, b, W, H));
    if (x0 >= g.w || y >= g.h) return;
    const size_t o = (size_t)b * W * H + (size_t)y * W + x0;
    const float inv_tw = 1.0f / (float)g.tw, inv_th = 1.0f / (float)g.th;
    const float tyf = (float)y * inv_th - 0.5f;
    int ty1 = (int)floorf(tyf);
    const float ya = tyf - (float)ty1, ya1 = 1.0f - ya;
    const int ty2 = min(ty1 + 1, 7);
    ty1 = max(ty1, 0);
    const uint8_t* L = tilelut + (size_t)b * 64 * 256;
    const uint8_t* L1 = L + ty1 * 8 * 256;
    const uint8_t* L2 = L + ty2 * 8 * 256;
    const uint8_t* pm = premap ? premap + b * 256 : nullptr;
    const bool vec = ((W & 3) == 0) && x0 + 4 <= g.w;
    uint32_t in4 = 0;
    if (vec) in4 = *reinterpret_cast<const uint32_t*>(src + o);
    else for (int k = 0; k < 4 && x0 + k < g.w; ++k) in4 |= (uint32_t)src[o + k] << (8 * k);
    uint32_t out4 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = x0 + k;
        int v = (in4 >> (8 * k)) & 255;
        if (pm) v = pm[v];
        const float txf = (float)x * inv_tw - 0.5f;
        int tx1 = (int)floorf(txf);
        const float xa = txf - (float)tx1, xa1 = 1.0f - xa;
        const int tx2 = min(tx1 + 1, 7);
        tx1 = max(tx1, 0);
        const float l11 = L1[tx1 * 256 + v], l12 = L1[tx2 * 256 + v];
        const float l21 = L2[tx1 * 256 + v], l22 = L2[tx2 * 256 + v];
        const float res = (l11 * xa1 + l12 * xa) * ya1 + (l21 * xa1 + l22 * xa) * ya;
        float r = rintf(res);
        r = fminf(fmaxf(r, 0.0f), 255.0f);
        out4 |= (uint32_t)r << (8 * k);
    }
    if (vec) *reinterpret_cast<uint32_t*>(dst + o) = out4;
    else for (int k = 0; k < 4 && x0 + k < g.w; ++k) dst[o + k] = (uint8_t)(out4 >> (8 * k));
}

void fpb_clahe(FpbLaunch L, const uint8_t* src, const uint8_t* premap, int n, int W, int H, const int4* roi,
               double clip, uint8_t* tilelut, uint8_t* dst) {
    dim3 g1(8, n);
    k_clahe_tiles<<<g1, 256, 0, L.st>>>(src, premap, W, H, roi, clip, tilelut);
    LAUNCH_COUNT(L);
    dim3 blk(32, 8), g2((W + 127) / 128, (H + 7) / 8, n);
    k_clahe_interp<<<g2, blk, 0, L.st>>>(src, premap, W, H, roi, tilelut, dst);
    LAUNCH_COUNT(L);
}

// ------------------------------------------------------------------------------------------------
// K2: non-local means, OpenCV FastNlMeansDenoisingInvoker<uchar,int,unsigned,DistSquared,int>
//     h=10, template 7x7, search 21x21, BORDER_REFLECT_101 by 13.
//
//   weight(p,o) = T[ SSD7x7(p, p+o) >> 6 ],  T[i] = round(19096*exp(-i*(64/49)/100)), 0 below 19.096
//   out(p) = (sum_o w*I(p+o) + sum_o w / 2) / sum_o w            (unsigned division)
//
// One CTA = 128 x 32 output pixels, image tile (+13 halo) staged once in shared memory by TMA, then as eight
// byte-shifted, pre-masked copies so that any 7-byte window is one aligned 8-byte load.
// One thread = 1 column x 16 rows: a 7-wide row SSD is LDS.64 + 2 x __vabsdiffu4 + 2 x __dp4a, the 7-row sums slide
// down the column in registers (22 row SSDs for 16 outputs), so a (pixel, offset) pair costs ~10 instructions
// instead of 49 multiply-adds; offsets whose weights are zero for the whole warp skip the table look-ups.
// ------------------------------------------------------------------------------------------------
#ifndef NLM_SLIDE_IADD3
#define NLM_SLIDE_IADD3 0
#endif
#define NLM_TW 128
#define NLM_TH 32
#define NLM_R 16                       // rows per thread
#define NLM_B 13                       // halo = 21/2 + 7/2
#define NLM_X0 16                      // tile column of output pixel 0: the tile starts 16 px left of the outputs so that
                                       // its first byte is 16-byte aligned in the image (TMA box / uint4 loads)
#define NLM_SW 176                     // smem row stride in bytes = TMA box width (multiple of 16)
#define NLM_ROWS (NLM_TH + 2 * NLM_B)  // 58
#define NLM_NW 529                     // non-zero weights: indices 0..527, [528] = 0
#define NLM_TILE_BYTES (NLM_ROWS * NLM_SW)
#define NLM_COPY_WORDS 2564            // words per byte-shifted copy of the tile: >= NLM_TILE_BYTES/4 (2552) and == 4 (mod 32),
                                       // so the 8-byte loads of a half-warp (8 copies x 2 neighbouring words) hit 32 distinct banks
#define NLM_RAW_WORDS 2560             // raw tile (TMA destination), 10 240 B >= NLM_TILE_BYTES
#define NLM_SMEM_BYTES ((NLM_RAW_WORDS + 8 * NLM_COPY_WORDS) * 4)

__constant__ int c_nlm_w[NLM_NW];

void fpb_upload_nlm_table(cudaStream_t st) {
    // almost_dist2weight of the OpenCV invoker for h = 10, template 7, search 21 (oracle/stages.py::nlm_weight_table)
    static int tab[NLM_NW];
    const int fixed_mult = 2147483647 / (21 * 21 * 255);           // 19096
    const double mult = 64.0 / 49.0;
    for (int i = 0; i < NLM_NW; ++i) {
        const double w = exp(-((double)i * mult) / (double)(10.0f * 10.0f));
        int v = (int)nearbyint(fixed_mult * w);
        if ((double)v < 0.001 * fixed_mult) v = 0;
        tab[i] = v;
    }
    tab[NLM_NW - 1] = 0;
    cudaMemcpyToSymbolAsync(c_nlm_w, tab, sizeof(tab), 0, cudaMemcpyHostToDevice, st);
    fpb_upload_nlm_table_mma(tab, st);
}

// ---- TMA (cp.async.bulk.tensor) helpers: the image batch is a 3-D u8 tensor (W, H, n); one box = one tile + halo.
// Out-of-image elements of the box arrive as zeros; the reflect-101 border of OpenCV is patched in afterwards.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_tile_3d(void* smem_dst, const CUtensorMap* tmap, uint64_t* mbar, int x, int y, int z, uint32_t bytes) {
    const uint32_t mb = smem_u32(mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // make the init visible to the async proxy
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smem_u32(smem_dst)), "l"((uint64_t)tmap), "r"(mb), "r"(x), "r"(y), "r"(z) : "memory");
}

__device__ __forceinline__ void mbar_wait_parity0(uint64_t* mbar) {
    const uint32_t mb = smem_u32(mbar);
    uint32_t done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(mb) : "memory");
        if (spin > (1u << 26)) __trap();         // a lost transaction must fail loudly, never hang the GPU
    }
}

template <bool USE_TMA>
__global__ void __launch_bounds__(256, 2)
k_nlm(const uint8_t* __restrict__ src, int W, int H, uint8_t* __restrict__ dst, const __grid_constant__ CUtensorMap tmap,
      unsigned one, unsigned mone) {
    // nlm_sm = [raw tile][copy 0 .. copy 7].  copy s = the tile shifted left by s bytes with byte 7 of every aligned
    // 8-byte group cleared: the 7-byte window that starts at ANY column cs is then exactly one aligned 8-byte load from
    // copy (cs & 7) - one LDS.64, no funnel shift and no mask in the offset loop.
    extern __shared__ __align__(128) uint32_t nlm_sm[];
    uint8_t* tile = reinterpret_cast<uint8_t*>(nlm_sm);
    uint32_t* copies = nlm_sm + NLM_RAW_WORDS;
    __shared__ int wtab[NLM_NW];
    __shared__ __align__(8) uint64_t mbar;
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * NLM_TW, y0 = blockIdx.y * NLM_TH;
    const int tx0 = x0 - NLM_X0, ty0 = y0 - NLM_B;           // image coordinates of tile byte (0,0)
    const uint8_t* p = src + (size_t)b * W * H;
    if (USE_TMA) {
        if (threadIdx.x == 0) tma_load_tile_3d(tile, &tmap, &mbar, tx0, ty0, b, NLM_TILE_BYTES);
        for (int i = threadIdx.x; i < NLM_NW; i += 256) wtab[i] = c_nlm_w[i];
        __syncthreads();                                     // mbarrier init visible to all before they wait on it
        mbar_wait_parity0(&mbar);
        // patch the reflected border (only tiles that stick out of the image have any): rows first, then columns
        const bool xin = tx0 >= 0 && tx0 + NLM_SW <= W, yin = ty0 >= 0 && ty0 + NLM_ROWS <= H;
        if (!(xin && yin)) {
            for (int i = threadIdx.x; i < NLM_TILE_BYTES; i += 256) {
                const int r = i / NLM_SW, c = i - r * NLM_SW;
                const int gx = tx0 + c, gy = ty0 + r;
                if ((unsigned)gx >= (unsigned)W || (unsigned)gy >= (unsigned)H)
                    tile[i] = p[(size_t)fpb_reflect101(gy, H) * W + fpb_reflect101(gx, W)];
            }
        }
    } else {
        for (int i = threadIdx.x; i < NLM_NW; i += 256) wtab[i] = c_nlm_w[i];
        for (int i = threadIdx.x; i < NLM_TILE_BYTES; i += 256) {
            const int r = i / NLM_SW, c = i - r * NLM_SW;
            tile[i] = p[(size_t)fpb_reflect101(ty0 + r, H) * W + fpb_reflect101(tx0 + c, W)];
        }
    }
    __syncthreads();
    // masked byte-shifted copies 0..7 (word w of copy s = bytes 4w+s .. 4w+s+3 of the tile; odd words lose their top byte)
    {
        const uint32_t* T = nlm_sm;
        for (int i = threadIdx.x; i < 8 * (NLM_TILE_BYTES / 4); i += 256) {
            const int sft = i / (NLM_TILE_BYTES / 4), w = i - sft * (NLM_TILE_BYTES / 4);
            const int w0 = w + (sft >> 2);
            const uint32_t lo = T[w0], hi = (w0 + 1 < NLM_TILE_BYTES / 4) ? T[w0 + 1] : 0u;
            uint32_t v = __funnelshift_r(lo, hi, (sft & 3) * 8);
            if (w & 1) v &= 0x00FFFFFFu;
            copies[sft * NLM_COPY_WORDS + w] = v;
        }
    }
    __syncthreads();
    // lane -> column: the 16 lanes of a half-warp take columns 8 apart (same copy, consecutive 8-byte words: conflict-free
    // LDS.64), the two half-warps neighbouring sub-columns
    const int lx = ((threadIdx.x & 15) << 3) | ((threadIdx.x >> 4) & 7), ty = threadIdx.x >> 7;
    const int row0 = ty * NLM_R + NLM_B - 3;      // first tile row of the unshifted 22-row strip
    const int col0 = lx + NLM_X0 - 3;             // first tile column of the unshifted 7-byte window
    // unshifted 7-byte windows, cached for the 22 rows of the strip
    uint32_t A0[NLM_R + 6], A1[NLM_R + 6];
    {
        const uint2* cp = reinterpret_cast<const uint2*>(copies + (col0 & 7) * NLM_COPY_WORDS) + (row0 * NLM_SW + (col0 & ~7)) / 8;
#pragma unroll
        for (int i = 0; i < NLM_R + 6; ++i) {
            const uint2 v = cp[i * (NLM_SW / 8)];
            A0[i] = v.x; A1[i] = v.y;
        }
    }
    unsigned est[NLM_R], wsum[NLM_R];
#pragma unroll
    for (int j = 0; j < NLM_R; ++j) { est[j] = 0; wsum[j] = 0; }

    for (int oy = -10; oy <= 10; ++oy) {
        for (int ox = -10; ox <= 10; ++ox) {
            const int cs = col0 + ox;
            const uint2* base = reinterpret_cast<const uint2*>(copies + (cs & 7) * NLM_COPY_WORDS) + ((row0 + oy) * NLM_SW + (cs & ~7)) / 8;
            unsigned rs[NLM_R + 6];
#pragma unroll
            for (int i = 0; i < NLM_R + 6; ++i) {
                const uint2 v = base[i * (NLM_SW / 8)];
                const uint32_t d0 = __vabsdiffu4(A0[i], v.x), d1 = __vabsdiffu4(A1[i], v.y);
                rs[i] = __dp4a(d0, d0, __dp4a(d1, d1, 0u));
            }
            // 7-row sliding sums.  On contrast-stretched ridge images ~98 % of the (pixel, offset) pairs have SSD >= 33792,
            // i.e. weight 0, and for ~80 % of the offsets that holds for the warp's whole 32 x 16 pixel tile: find the
            // smallest SSD of the strip first (2 instructions per output) and skip the table lookups / accumulations
            // warp-uniformly when no lane can contribute.
            // (the ALU pipe - VABSDIFF4 / IDP.4A / min - is the busy one: the window updates are written as multiply-adds
            //  by the run-time constants +1 / -1 so that they issue on the otherwise idle FMA pipe)
            const unsigned S0 = rs[0] + rs[1] + rs[2] + rs[3] + rs[4] + rs[5] + rs[6];
            unsigned S = S0, smin = S0;
#if NLM_SLIDE_IADD3
            // one three-input add per window step (IADD3 issues on either integer pipe) and one three-input minimum per two steps
            unsigned Sv[NLM_R];
            Sv[0] = S0;
#pragma unroll
            for (int j = 1; j < NLM_R; ++j) { S = S + rs[j + 6] - rs[j - 1]; Sv[j] = S; }
#pragma unroll
            for (int j = 1; j + 1 < NLM_R; j += 2) smin = min(min(smin, Sv[j]), Sv[j + 1]);
            smin = min(smin, Sv[NLM_R - 1]);
#else
#pragma unroll
            for (int j = 1; j < NLM_R; ++j) { S = rs[j + 6] * one + S; S = rs[j - 1] * mone + S; smin = min(smin, S); }
#endif
            if (!__any_sync(0xffffffffu, smin < (unsigned)((NLM_NW - 1) << 6))) continue;
            S = S0;
            const uint8_t* pc = tile + (row0 + 3 + oy) * NLM_SW + lx + NLM_X0 + ox;
#pragma unroll
            for (int j = 0; j < NLM_R; ++j) {
#if NLM_SLIDE_IADD3
                S = Sv[j];
#endif
                const unsigned idx = min(S >> 6, (unsigned)(NLM_NW - 1));
                const unsigned w = (unsigned)wtab[idx];
                est[j] += w * (unsigned)pc[j * NLM_SW];
                wsum[j] += w;
#if !NLM_SLIDE_IADD3
                if (j + 1 < NLM_R) { S = rs[j + 7] * one + S; S = rs[j] * mone + S; }
#endif
            }
        }
    }
    // results: through shared memory (the raw tile is free now) so that the global stores are row-contiguous
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NLM_R; ++j)
        tile[(ty * NLM_R + j) * NLM_TW + lx] = (uint8_t)min((est[j] + wsum[j] / 2u) / wsum[j], 255u);
    __syncthreads();
    for (int i = threadIdx.x; i < NLM_TW * NLM_TH; i += 256) {
        const int r = i / NLM_TW, c = i - r * NLM_TW;
        const int gx = x0 + c, gy = y0 + r;
        if (gx < W && gy < H) dst[(size_t)b * W * H + (size_t)gy * W + gx] = tile[i];
    }
}


// ---- k_nlm3: the same arithmetic with the candidate rows REUSED across three vertical offsets ---------------------------
// k_nlm is bound by shared-memory wavefronts (83 %): every (thread, offset) re-loads its 22 candidate row windows.  For a
// fixed horizontal offset the windows of oy, oy+1, oy+2 are the same rows shifted by one, so this kernel loads 24 rows once
// per THREE offsets (8 LDS.64 per offset instead of 22) and keeps them in registers next to the 22 own rows.  That costs
// ~40 more registers, so the CTA is one 16-row strip of 128 consecutive columns (128 threads, 42 x 176 B tile, 67 KB of
// shared memory: three CTAs per SM, their tile set-up phases staggered).  Window steps are one IADD3, minima pair into VIMNMX3.
#define NLM3_TH 16
#define NLM3_ROWS (NLM3_TH + 2 * NLM_B)            // 42
#define NLM3_TILE_BYTES (NLM3_ROWS * NLM_SW)       // 7392
#define NLM3_RAW_WORDS 1856                        // >= 7392 / 4 = 1848
#define NLM3_COPY_WORDS 1848                       // words per copy (7392 / 4)
#define NLM3_SMEM_BYTES ((NLM3_RAW_WORDS + 8 * NLM3_COPY_WORDS) * 4)

template <bool USE_TMA>
__global__ void __launch_bounds__(128, 3)
k_nlm3(const uint8_t* __restrict__ src, int W, int H, uint8_t* __restrict__ dst, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) uint32_t nlm_sm[];
    uint8_t* tile = reinterpret_cast<uint8_t*>(nlm_sm);
    uint32_t* copies = nlm_sm + NLM3_RAW_WORDS;
    __shared__ int wtab[NLM_NW];
    __shared__ __align__(8) uint64_t mbar;
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * NLM_TW, y0 = blockIdx.y * NLM3_TH;
    const int tx0 = x0 - NLM_X0, ty0 = y0 - NLM_B;           // image coordinates of tile byte (0,0)
    const uint8_t* p = src + (size_t)b * W * H;
    if (USE_TMA) {
        if (threadIdx.x == 0) tma_load_tile_3d(tile, &tmap, &mbar, tx0, ty0, b, NLM3_TILE_BYTES);
        for (int i = threadIdx.x; i < NLM_NW; i += 128) wtab[i] = c_nlm_w[i];
        __syncthreads();                                     // mbarrier init visible to all before they wait on it
        mbar_wait_parity0(&mbar);
        const bool xin = tx0 >= 0 && tx0 + NLM_SW <= W, yin = ty0 >= 0 && ty0 + NLM3_ROWS <= H;
        if (!(xin && yin)) {
            for (int i = threadIdx.x; i < NLM3_TILE_BYTES; i += 128) {
                const int r = i / NLM_SW, c = i - r * NLM_SW;
                const int gx = tx0 + c, gy = ty0 + r;
                if ((unsigned)gx >= (unsigned)W || (unsigned)gy >= (unsigned)H)
                    tile[i] = p[(size_t)fpb_reflect101(gy, H) * W + fpb_reflect101(gx, W)];
            }
        }
    } else {
        for (int i = threadIdx.x; i < NLM_NW; i += 128) wtab[i] = c_nlm_w[i];
        for (int i = threadIdx.x; i < NLM3_TILE_BYTES; i += 128) {
            const int r = i / NLM_SW, c = i - r * NLM_SW;
            tile[i] = p[(size_t)fpb_reflect101(ty0 + r, H) * W + fpb_reflect101(tx0 + c, W)];
        }
    }
    __syncthreads();
    {   // masked byte-shifted copies 0..7 (as k_nlm), laid out [row][8-byte group][copy]: the eight copies' words of one group are
        // contiguous (64 B), so 16 consecutive columns - any alignment - read 32 distinct banks with one LDS.64 each
        const uint32_t* T = nlm_sm;
        for (int i = threadIdx.x; i < 8 * (NLM3_TILE_BYTES / 4); i += 128) {
            const int sft = i / (NLM3_TILE_BYTES / 4), w = i - sft * (NLM3_TILE_BYTES / 4);
            const int w0 = w + (sft >> 2);
            const uint32_t lo = T[w0], hi = (w0 + 1 < NLM3_TILE_BYTES / 4) ? T[w0 + 1] : 0u;
            uint32_t v = __funnelshift_r(lo, hi, (sft & 3) * 8);
            if (w & 1) v &= 0x00FFFFFFu;
            const int r = w / (NLM_SW / 4), wr = w - r * (NLM_SW / 4);
            copies[(((r * (NLM_SW / 8) + (wr >> 1)) * 8 + sft) << 1) + (wr & 1)] = v;
        }
    }
    __syncthreads();
    // consecutive lanes = consecutive columns: a warp covers a compact 32 x 16 pixel patch, which makes the warp-uniform skip
    // of weight-less offsets fire much more often than a warp spread over 128 columns (20.5 vs 22.2 ms)
    const int lx = threadIdx.x;
    const int row0 = NLM_B - 3;                   // first tile row of the unshifted 22-row strip
    const int col0 = lx + NLM_X0 - 3;             // first tile column of the unshifted 7-byte window
    uint32_t A0[NLM_R + 6], A1[NLM_R + 6];
    {
        const uint2* cp = reinterpret_cast<const uint2*>(copies) + (row0 * (NLM_SW / 8) + (col0 >> 3)) * 8 + (col0 & 7);
#pragma unroll
        for (int i = 0; i < NLM_R + 6; ++i) { const uint2 v = cp[i * NLM_SW]; A0[i] = v.x; A1[i] = v.y; }
    }
    unsigned est[NLM_R], wsum[NLM_R];
#pragma unroll
    for (int j = 0; j < NLM_R; ++j) { est[j] = 0; wsum[j] = 0; }

    for (int ox = -10; ox <= 10; ++ox) {
        const int cs = col0 + ox;
        const uint2* colbase = reinterpret_cast<const uint2*>(copies) + (cs >> 3) * 8 + (cs & 7);
#pragma unroll 1
        for (int oyb = 0; oyb < 7; ++oyb) {       // vertical offsets oy = -10 + 3 oyb + {0, 1, 2}: tile rows 3 oyb .. 3 oyb + 23
            uint32_t B0[NLM_R + 8], B1[NLM_R + 8];
            const uint2* base = colbase + (3 * oyb) * NLM_SW;
#pragma unroll
            for (int i = 0; i < NLM_R + 8; ++i) { const uint2 v = base[i * NLM_SW]; B0[i] = v.x; B1[i] = v.y; }
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                unsigned rs[NLM_R + 6];
#pragma unroll
                for (int i = 0; i < NLM_R + 6; ++i) {
                    const uint32_t d0 = __vabsdiffu4(A0[i], B0[i + d]), d1 = __vabsdiffu4(A1[i], B1[i + d]);
                    rs[i] = __dp4a(d0, d0, __dp4a(d1, d1, 0u));
                }
                unsigned Sv[NLM_R];
                Sv[0] = rs[0] + rs[1] + rs[2] + rs[3] + rs[4] + rs[5] + rs[6];
#pragma unroll
                for (int j = 1; j < NLM_R; ++j) Sv[j] = Sv[j - 1] + rs[j + 6] - rs[j - 1];
                // ~98 % of the pairs have weight 0; the table look-ups are skipped warp-uniformly, separately for the upper and the
                // lower half of the strip (a 32 x 8 pixel patch is weight-less more often than a 32 x 16 one)
                const int oy = -10 + 3 * oyb + d;
                const uint8_t* pc = tile + (row0 + 3 + oy) * NLM_SW + lx + NLM_X0 + ox;
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                    const int j0 = hlf * (NLM_R / 2);
                    unsigned smin = min(min(Sv[j0], Sv[j0 + 1]), Sv[j0 + 2]);
                    smin = min(min(smin, Sv[j0 + 3]), Sv[j0 + 4]);
                    smin = min(min(smin, Sv[j0 + 5]), min(Sv[j0 + 6], Sv[j0 + 7]));
                    if (!__any_sync(0xffffffffu, smin < (unsigned)((NLM_NW - 1) << 6))) continue;
#pragma unroll
                    for (int j = j0; j < j0 + NLM_R / 2; ++j) {
                        const unsigned idx = min(Sv[j] >> 6, (unsigned)(NLM_NW - 1));
                        const unsigned w = (unsigned)wtab[idx];
                        est[j] += w * (unsigned)pc[j * NLM_SW];
                        wsum[j] += w;
                    }
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NLM_R; ++j) tile[j * NLM_TW + lx] = (uint8_t)min((est[j] + wsum[j] / 2u) / wsum[j], 255u);
    __syncthreads();
    for (int i = threadIdx.x; i < NLM_TW * NLM3_TH / 4; i += 128) {                 // four pixels per store where the row allows it
        const int r = i / (NLM_TW / 4), c = (i - r * (NLM_TW / 4)) * 4;
        const int gx = x0 + c, gy = y0 + r;
        if (gy >= H || gx >= W) continue;
        uint8_t* o = dst + (size_t)b * W * H + (size_t)gy * W + gx;
        if (gx + 3 < W && (W & 3) == 0) *reinterpret_cast<uint32_t*>(o) = *reinterpret_cast<const uint32_t*>(tile + r * NLM_TW + c);
        else for (int u = 0; u < 4 && gx + u < W; ++u) o[u] = tile[r * NLM_TW + c + u];
    }
}


// ---- k_nlm_sym: every patch distance computed ONCE and used for both of its pixels -----------------------------------------
// SSD(p, q) is symmetric in the reflect-padded image and the weight is a function of the SSD alone, so the pair {p, q = p + o}
// needs one distance for the two contributions  est(p) += w I(q)  and  est(q) += w I(p).  The kernel walks only the half plane
// H+ = {ox > 0} u {ox = 0, oy > 0} (220 of the 440 non-zero offsets), keeps the p side in registers as k_nlm3 does, and adds the
// q side - for the ~2 % of the pairs whose weight is not zero - with shared-memory atomics into a ring of accumulator rows.  The
// sums are unsigned integers, so the order of the additions does not matter: the result stays bit-identical to OpenCV's.
// The centre offset (SSD 0, weight T[0]) is added when a row is written.
//   * one CTA = one band of <= 240 image columns (lane = p column x0 - 10 + lane: the 10 columns left of the band are walked for
//     their q side only) x one vertical segment of the image, processed top to bottom in chunks of 16 p rows starting 10 rows
//     above the segment and ending 10 rows below it;
//   * per chunk: 42 x 288 B tile by two TMA boxes (the next chunk's tile is in flight during the offset loop), eight masked
//     byte-shifted copies in the [row][8-byte group][copy] layout of k_nlm3, then the k_nlm3 offset loop over 11 x 21 offsets;
//   * accumulator ring: 48 rows x 264 columns x (est, wsum); after chunk k the 16 rows that no later chunk touches are
//     normalised, written and cleared.
// Work per image: 8 warps x 22 chunks x 220 offsets against 160 warps x 441 offsets for k_nlm3 (0.55 x).
#define NLMS_SW 288                                   // tile / copy row pitch in bytes = two TMA boxes of 144
#define NLMS_BOXW 144
#define NLMS_R 20                                     // p rows per chunk (the row-streaming body keeps only a half strip in registers)
#define NLMS_G 10                                     // rows per warp-uniform skip group
#define NLMS_ROWS (NLMS_R + 2 * NLM_B)                // 46 tile rows: the p rows + 13 above / below
#define NLMS_BOX_WORDS 1664                           // 144 x 46 = 6624 B per box, padded to 6656 B (128-byte aligned TMA destinations)
#define NLMS_GROUPS (NLMS_SW / 8)                     // 36 eight-byte groups per row
#define NLMS_COPY_WORDS (NLMS_ROWS * NLMS_SW * 2)     // all eight copies: 26 496 words
#define NLMS_RING (NLMS_R + 20)                       // accumulator rows alive at a time: the chunk's p rows and ten above / below
#define NLMS_AW 264                                   // accumulator columns: q columns x0 - 10 .. x0 + TW + 9 (<= 260)
#define NLMS_SMEM_BYTES ((2 * NLMS_BOX_WORDS + NLMS_COPY_WORDS + 2 * NLMS_RING * NLMS_AW) * 4)     // 203 776

__device__ __forceinline__ void mbar_wait_parity(uint64_t* mbar, uint32_t parity) {
    const uint32_t mb = smem_u32(mbar);
    uint32_t done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(mb), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();         // a lost transaction must fail loudly, never hang the GPU
    }
}

__device__ __forceinline__ void nlms_issue_tile(uint32_t* raw, const CUtensorMap* tmap, uint64_t* mbar, int tx0, int ty0, int b) {
    const uint32_t mb = smem_u32(mbar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the previous tile was read / patched through the generic proxy
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(2 * NLMS_BOXW * NLMS_ROWS) : "memory");
#pragma unroll
    for (int bx = 0; bx < 2; ++bx)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     :: "r"(smem_u32(raw + bx * NLMS_BOX_WORDS)), "l"((uint64_t)tmap), "r"(mb), "r"(tx0 + bx * NLMS_BOXW), "r"(ty0), "r"(b)
                     : "memory");
}

// one lane's fetch-and-increment of the unit queue (plain atom.shared: the compiler's warp-aggregation preamble around
// atomicAdd costs a dozen instructions per unit and there is a single caller per warp anyway)
__device__ __forceinline__ int nlms_next_unit(int* ctr) {
    int v;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(v) : "r"(smem_u32(ctr)) : "memory");
    return v;
}

__device__ __forceinline__ uint32_t nlms_raw_word(const uint32_t* raw, int r, int wr) {     // word wr (0..71) of tile row r
    if (wr >= NLMS_SW / 4) return 0u;
    return raw[(wr >= NLMS_BOXW / 4 ? NLMS_BOX_WORDS - NLMS_BOXW / 4 : 0) + r * (NLMS_BOXW / 4) + wr];
}

static_assert(NLMS_R % NLMS_G == 0 && NLMS_RING % NLMS_G == 0 && 10 % NLMS_G == 0, "skip groups tile the strip and wrap around the ring whole");
#define NLMS_THREADS 768                              // 12 warps pull (strip, ox, oy-group) units from a per-chunk queue
#define NLMS_GROUPS_PER_STRIP 74                      // ox = 0: oy groups 3..6; ox = 1..10: oy groups 0..6

template <bool USE_TMA>
__global__ void __launch_bounds__(NLMS_THREADS, 1)
k_nlm_sym(const uint8_t* __restrict__ src, int W, int H, uint8_t* __restrict__ dst, const __grid_constant__ CUtensorMap tmap,
          int TW, int seg_rows) {
    extern __shared__ __align__(128) uint32_t nlm_sm[];
    uint32_t* raw = nlm_sm;
    uint32_t* copies = nlm_sm + 2 * NLMS_BOX_WORDS;
    uint32_t* accE = copies + NLMS_COPY_WORDS;
    uint32_t* accW = accE + NLMS_RING * NLMS_AW;
    __shared__ int wtab[NLM_NW];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ int unit_ctr;
    __shared__ uint32_t utab[8 * NLMS_GROUPS_PER_STRIP];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
    const int b = blockIdx.z, x0 = blockIdx.x * TW;
    const int ya = blockIdx.y * seg_rows, yb = min(ya + seg_rows, H);
    if (ya >= H) return;
    const int nchunks = (yb - ya + 20 + NLMS_R - 1) / NLMS_R;
    const uint8_t* p = src + (size_t)b * W * H;
    const int nl = TW + 10;                       // p columns x0 - 10 .. x0 + TW - 1 (the first ten for their q side only)
    const int nstrips = (nl + 31) >> 5, nunits = nstrips * NLMS_GROUPS_PER_STRIP;
    const int tx0 = x0 - 16;                      // image column of tile column 0
    const int cmax = TW + 29;                     // tile columns 3 .. cmax - 1 are read
    for (int i = tid; i < NLM_NW; i += nthr) wtab[i] = c_nlm_w[i];
    for (int i = tid; i < 2 * NLMS_RING * NLMS_AW; i += nthr) accE[i] = 0u;
    for (int u = tid; u < nunits; u += nthr) {    // unit -> strip | ox << 8 | oyb << 16 (oy = -10 + 3 oyb + {0, 1, 2}; ox = 0: oy = 1 .. 10)
        const int g = u / nstrips, strip = u - g * nstrips;
        int ox, oyb;
        if (g < 4) { ox = 0; oyb = 3 + g; } else { const int t = g - 4; ox = 1 + t / 7; oyb = t - (ox - 1) * 7; }
        utab[u] = (uint32_t)(strip | (ox << 8) | (oyb << 16));
    }
    if (USE_TMA && tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        nlms_issue_tile(raw, &tmap, &mbar, tx0, ya - 10 - NLM_B, b);
    }
    __syncthreads();
    const uint2* C2 = reinterpret_cast<const uint2*>(copies);

    // tile of chunk k: wait for its TMA boxes (out-of-image elements arrive as zeros and are patched with OpenCV's reflect-101) or
    // load it with plain loads.  Called for chunk 0 before the loop and for chunk k + 1 next to the output phase of chunk k.
    auto fetch_tile = [&](int k) {
        const int ty0 = ya - 10 + NLMS_R * k - NLM_B;           // image row of tile row 0
        if (USE_TMA) mbar_wait_parity(&mbar, (uint32_t)(k & 1));
        const bool inside = USE_TMA && tx0 + 3 >= 0 && tx0 + cmax <= W && ty0 >= 0 && ty0 + NLMS_ROWS <= H;
        if (inside) return;
        uint8_t* rawb = reinterpret_cast<uint8_t*>(raw);
        for (int r = warp; r < NLMS_ROWS; r += nwarps) {
            const int gy = ty0 + r;
            const bool yin = (unsigned)gy < (unsigned)H;
            const uint8_t* q = p + (size_t)fpb_reflect101(gy, H) * W;
            // rows inside the image that TMA delivered: only the columns left / right of the image need the reflection
            const int c_skip0 = (USE_TMA && yin) ? max(3, -tx0) : cmax, c_skip1 = (USE_TMA && yin) ? min(cmax, W - tx0) : cmax;
            for (int c = 3 + lane; c < cmax; c += 32) {
                if (c >= c_skip0 && c < c_skip1) { c += (c_skip1 - 1 - c) / 32 * 32; continue; }   // jump over the in-image span
                const int gx = tx0 + c;
                rawb[(c >= NLMS_BOXW ? NLMS_BOX_WORDS * 4 - NLMS_BOXW : 0) + r * NLMS_BOXW + c] = q[fpb_reflect101(gx, W)];
            }
        }
    };
    fetch_tile(0);
    __syncthreads();

    for (int k = 0; k < nchunks; ++k) {
        const int pr0 = ya - 10 + NLMS_R * k;      // first p row of the chunk
        const int ty0 = pr0 - NLM_B;              // image row of tile row 0
        // ---- eight masked byte-shifted copies: copy s of a group = bytes 8g + s .. 8g + s + 6 (byte 7 cleared); one thread
        //      forms two copies of one group and stores them as one 16-byte word (consecutive threads, consecutive words)
        for (int i = tid; i < NLMS_ROWS * NLMS_GROUPS * 4; i += nthr) {
            const int pr = i & 3, gi = i >> 2;
            const int r = gi / NLMS_GROUPS, g = gi - r * NLMS_GROUPS;
            const int wb = 2 * g + (pr >> 1);
            const uint32_t u0 = nlms_raw_word(raw, r, wb), u1 = nlms_raw_word(raw, r, wb + 1), u2 = nlms_raw_word(raw, r, wb + 2);
            const int sh = (pr & 1) * 16;
            uint4 v;
            v.x = __funnelshift_r(u0, u1, sh);     v.y = __funnelshift_r(u1, u2, sh) & 0x00FFFFFFu;
            v.z = __funnelshift_r(u0, u1, sh + 8); v.w = __funnelshift_r(u1, u2, sh + 8) & 0x00FFFFFFu;
            reinterpret_cast<uint4*>(copies)[i] = v;
        }
        if (tid == 0) unit_ctr = 0;
        __syncthreads();
        if (USE_TMA && tid == 0 && k + 1 < nchunks) nlms_issue_tile(raw, &tmap, &mbar, tx0, ty0 + NLMS_R, b);

        // ---- offset loop.  A unit = one 32-column strip x one horizontal offset x three vertical offsets (k_nlm3's inner body);
        //      the warps take units from a queue, so no warp waits at the end of the chunk for one that met more live weights.
        const int rb = (NLMS_R * k) % NLMS_RING;   // ring row of image row pr0 - 10
        int u = 0;
        if (lane == 0) u = nlms_next_unit(&unit_ctr);
        u = __shfl_sync(0xffffffffu, u, 0);
        while (u < nunits) {
            int u_next = 0;
            if (lane == 0) u_next = nlms_next_unit(&unit_ctr);      // asked for now, needed at the end of this unit
            const uint32_t ut = utab[u];
            const int strip = ut & 255, ox = (ut >> 8) & 255, oyb = ut >> 16;
            const int lcol = strip * 32 + lane;
            const bool lane_ok = lcol < nl;       // surplus lanes repeat the last column and add nothing
            const int lx = min(lcol, nl - 1);
            const int col0 = lx + 3;              // first tile column of the unshifted 7-byte window
            const int cs = col0 + ox;
            // accumulator columns of p and of q = p + (ox, .); ring rows rb + 10 + j and rq0 + d + j wrap at most once
            int rq0 = rb + 3 * oyb; rq0 -= (rq0 >= NLMS_RING) ? NLMS_RING : 0;
            uint32_t* const pE = accE + lx + (rb + 10) * NLMS_AW;
            uint32_t* const qE = accE + lx + ox + rq0 * NLMS_AW;
            const int pwrap = NLMS_RING - (rb + 10), qwrap = NLMS_RING - rq0;
            // SSD below thr <=> weight != 0; surplus lanes never; for ox = 0 the group oy = -1, 0, 1 keeps oy = 1 only
            const unsigned thr2 = lane_ok ? (unsigned)((NLM_NW - 1) << 6) : 0u;
            const unsigned thr01 = (ox == 0 && oyb == 3) ? 0u : thr2;
            // Row-streaming form of k_nlm3's body: the own row i and the candidate rows i, i + 1, i + 2 are loaded when row i is
            // due, the three offsets' row SSDs slide into the window sums at once, and only the sums of the current half strip
            // stay in registers - ~80 live registers instead of ~160, so twice the warps hide the latencies.
            const uint2* const ap = C2 + ((NLM_B - 3) * NLMS_GROUPS + (col0 >> 3)) * 8 + (col0 & 7);
            const uint2* const bp = C2 + (3 * oyb * NLMS_GROUPS + (cs >> 3)) * 8 + (cs & 7);
            unsigned rs[3][NLMS_R + 6], Sv[3][NLMS_R];
            uint2 bw[NLMS_R + 8];
            bw[0] = bp[0]; bw[1] = bp[NLMS_SW];
#pragma unroll
            for (int i = 0; i < NLMS_R + 6; ++i) {
                const uint2 av = ap[i * NLMS_SW];
                bw[i + 2] = bp[(i + 2) * NLMS_SW];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const uint32_t d0 = __vabsdiffu4(av.x, bw[i + d].x), d1 = __vabsdiffu4(av.y, bw[i + d].y);
                    rs[d][i] = __dp4a(d0, d0, __dp4a(d1, d1, 0u));
                    if (i == 6) Sv[d][0] = rs[d][0] + rs[d][1] + rs[d][2] + rs[d][3] + rs[d][4] + rs[d][5] + rs[d][6];
                    if (i > 6) Sv[d][i - 6] = Sv[d][i - 7] + rs[d][i] - rs[d][i - 7];
                }
                if (i >= 6 && (i - 5) % NLMS_G == 0) {    // a group of rows j0 .. j0 + G - 1 of the three offsets is complete
                    const int j0 = i - 5 - NLMS_G;
                    bool any[3];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        unsigned smin = Sv[d][j0];
#pragma unroll
                        for (int j = 1; j < NLMS_G; ++j) smin = min(smin, Sv[d][j0 + j]);
                        any[d] = __any_sync(0xffffffffu, smin < (d < 2 ? thr01 : thr2));
                    }
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        if (!any[d]) continue;    // ~98 % of the pairs have weight 0: warp-uniform skip per half strip
                        const unsigned thr = d < 2 ? thr01 : thr2;
                        // the p rows of a group wrap around the ring together (ring = 2 R, group = R / 2)
                        uint32_t* const pEg = (j0 >= pwrap ? pE - NLMS_RING * NLMS_AW : pE);
#pragma unroll
                        for (int j = j0; j < j0 + NLMS_G; ++j) {
                            if (Sv[d][j] < thr) {
                                const unsigned w = (unsigned)wtab[Sv[d][j] >> 6];
                                // centres of the candidate and of the own patch: byte 3 of their middle rows
                                const unsigned iq = reinterpret_cast<const uint8_t*>(bp + (j + 3 + d) * NLMS_SW)[3];
                                const unsigned ip = reinterpret_cast<const uint8_t*>(ap + (j + 3) * NLMS_SW)[3];
                                uint32_t* const pa = pEg + j * NLMS_AW;
                                uint32_t* const qa = (d + j >= qwrap ? qE - NLMS_RING * NLMS_AW : qE) + (d + j) * NLMS_AW;
                                atomicAdd(pa, w * iq);
                                atomicAdd(pa + NLMS_RING * NLMS_AW, w);
                                atomicAdd(qa, w * ip);
                                atomicAdd(qa + NLMS_RING * NLMS_AW, w);
                            }
                        }
                    }
                }
            }
            u = __shfl_sync(0xffffffffu, u_next, 0);
        }
        __syncthreads();
        // ---- write (and clear) the ring rows that no later chunk touches: est + T[0] I(p) over wsum + T[0].  I(p) is byte 0 of
        //      copy (column & 7) of the chunk's tile, still in shared memory.  The next chunk's tile is waited for / patched in the
        //      same phase (it touches the raw buffer only).
        {
            const int nrows = (k == nchunks - 1) ? NLMS_RING : NLMS_R;
            const unsigned w0 = (unsigned)wtab[0];
            const uint8_t* cbytes = reinterpret_cast<const uint8_t*>(copies);
            for (int rr = warp; rr < nrows; rr += nwarps) {
                const int y = pr0 - 10 + rr;
                int r = rb + rr; r -= (r >= NLMS_RING) ? NLMS_RING : 0;
                const bool yok = y >= ya && y < yb;
                for (int c = lane; c < NLMS_AW; c += 32) {
                    const int x = x0 + c - 10;
                    if (yok && c >= 10 && c < 10 + TW && x < W) {
                        const int tc = c + 6;                   // tile column of image column x: x - tx0 = c - 10 + 16
                        const unsigned ctr = cbytes[(((rr + 3) * NLMS_GROUPS + (tc >> 3)) * 8 + (tc & 7)) * 8];
                        const unsigned e = accE[r * NLMS_AW + c] + w0 * ctr, ws = accW[r * NLMS_AW + c] + w0;
                        dst[(size_t)b * W * H + (size_t)y * W + x] = (uint8_t)min((e + ws / 2u) / ws, 255u);
                    }
                    accE[r * NLMS_AW + c] = 0u; accW[r * NLMS_AW + c] = 0u;
                }
            }
        }
        if (k + 1 < nchunks) fetch_tile(k + 1);
        __syncthreads();
    }
}

// cuTensorMapEncodeTiled through the runtime's driver-entry-point query (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr; static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)ptr;
        (void)cudaGetLastError();
    }
    return fn;
}

// Launch geometry of k_nlm_sym: bands of <= 240 columns (a multiple of 16), (band width + 10) lanes rounded up to whole warps,
// and as many vertical segments as make the number of CTA waves x chunks per CTA smallest (every segment re-walks 20 rows).
static void fpb_nlm_sym(FpbLaunch L, const uint8_t* src, int n, int W, int H, uint8_t* dst) {
    static int n_sm = 0;
    if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (n_sm < 1) n_sm = 148; }
    const int nb = (W + 239) / 240;
    const int TW = (((W + nb - 1) / nb) + 15) / 16 * 16;
    static const int env_thr = getenv("FPB_NLM_THREADS") ? atoi(getenv("FPB_NLM_THREADS")) : 0;     // experiment: fewer warps per CTA
    const int nthr = (env_thr >= 32 && env_thr <= NLMS_THREADS) ? env_thr / 32 * 32 : NLMS_THREADS;
    static const int force_segs = getenv("FPB_NLM_SEGS") ? atoi(getenv("FPB_NLM_SEGS")) : 0;
    int segs = 1; long long best = -1;
    for (int s = 1; s <= 16; ++s) {
        const int rows = (H + s - 1) / s;
        if (s > 1 && rows < 16) break;
        const long long waves = ((long long)n * nb * s + n_sm - 1) / n_sm, chunks = (rows + 20 + NLMS_R - 1) / NLMS_R;
        if (best < 0 || waves * chunks < best) { best = waves * chunks; segs = s; }
    }
    if (force_segs > 0) segs = force_segs;
    const int seg_rows = (H + segs - 1) / segs;
    segs = (H + seg_rows - 1) / seg_rows;
    CUtensorMap tmap; memset(&tmap, 0, sizeof(tmap));
    bool use_tma = false;
    static const bool no_tma = getenv("FPB_NO_TMA") != nullptr;
    EncodeTiledFn enc = no_tma ? nullptr : get_encode_tiled();
    if (enc && (W % 16) == 0 && (((uintptr_t)src) % 16) == 0) {
        const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
        const cuuint64_t gstr[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
        const cuuint32_t box[3] = {NLMS_BOXW, NLMS_ROWS, 1}, estr[3] = {1, 1, 1};
        use_tma = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    FPB_OPT_IN_SMEM(k_nlm_sym<true>, NLMS_SMEM_BYTES);
    FPB_OPT_IN_SMEM(k_nlm_sym<false>, NLMS_SMEM_BYTES);
    dim3 grid(nb, segs, n);
    if (use_tma) k_nlm_sym<true><<<grid, nthr, NLMS_SMEM_BYTES, L.st>>>(src, W, H, dst, tmap, TW, seg_rows);
    else k_nlm_sym<false><<<grid, nthr, NLMS_SMEM_BYTES, L.st>>>(src, W, H, dst, tmap, TW, seg_rows);
    LAUNCH_COUNT(L);
}

void fpb_nlm(FpbLaunch L, const uint8_t* src, int n, int W, int H, uint8_t* dst) {
    // two bit-exact formulations: the tensor-core one (k_nlm_mma.cu, FPB_NLM_MMA=1) and the integer-ALU kernel below;
    // the default is whichever measures faster on the 1480-image batch (DESIGN.md section 4 keeps the A/B numbers)
    static const bool mma = getenv("FPB_NLM_MMA") != nullptr && getenv("FPB_NLM_MMA")[0] == '1';
    if (mma && fpb_nlm_mma(L, src, n, W, H, dst)) return;
    // FPB_NLM_V=1: k_nlm (22 row loads per offset, 128 x 32 tiles); FPB_NLM_V=3: k_nlm3 (candidate rows shared by three
    // offsets); default: k_nlm_sym (k_nlm3's offset loop over half of the offsets, each distance used for both of its pixels)
    static const char* ver = getenv("FPB_NLM_V");
    static const bool v1 = ver != nullptr && ver[0] == '1';
    if (!(ver != nullptr && (ver[0] == '1' || ver[0] == '3'))) { fpb_nlm_sym(L, src, n, W, H, dst); return; }
    const int th = v1 ? NLM_TH : NLM3_TH;
    dim3 grid((W + NLM_TW - 1) / NLM_TW, (H + th - 1) / th, n);
    CUtensorMap tmap; memset(&tmap, 0, sizeof(tmap));
    bool use_tma = false;
    static const bool no_tma = getenv("FPB_NO_TMA") != nullptr;
    EncodeTiledFn enc = no_tma ? nullptr : get_encode_tiled();
    // TMA needs 16-byte aligned base and strides (W % 16 == 0); other widths take the plain load path
    if (enc && (W % 16) == 0 && (((uintptr_t)src) % 16) == 0) {
        const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
        const cuuint64_t gstr[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
        const cuuint32_t box[3] = {NLM_SW, (cuuint32_t)(v1 ? NLM_ROWS : NLM3_ROWS), 1}, estr[3] = {1, 1, 1};
        use_tma = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    if (!v1) {
        FPB_OPT_IN_SMEM(k_nlm3<true>, NLM3_SMEM_BYTES);
        FPB_OPT_IN_SMEM(k_nlm3<false>, NLM3_SMEM_BYTES);
        static const int pad = getenv("FPB_NLM3_SMEM_PAD") ? atoi(getenv("FPB_NLM3_SMEM_PAD")) : 0;   // experiment: fewer CTAs per SM
        if (pad) { cudaFuncSetAttribute(k_nlm3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NLM3_SMEM_BYTES + pad);
                   cudaFuncSetAttribute(k_nlm3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NLM3_SMEM_BYTES + pad); }
        if (use_tma) k_nlm3<true><<<grid, 128, NLM3_SMEM_BYTES + pad, L.st>>>(src, W, H, dst, tmap);
        else k_nlm3<false><<<grid, 128, NLM3_SMEM_BYTES + pad, L.st>>>(src, W, H, dst, tmap);
        LAUNCH_COUNT(L);
        return;
    }
    FPB_OPT_IN_SMEM(k_nlm<true>, NLM_SMEM_BYTES);
    FPB_OPT_IN_SMEM(k_nlm<false>, NLM_SMEM_BYTES);
    if (use_tma) k_nlm<true><<<grid, 256, NLM_SMEM_BYTES, L.st>>>(src, W, H, dst, tmap, 1u, 0xFFFFFFFFu);
    else k_nlm<false><<<grid, 256, NLM_SMEM_BYTES, L.st>>>(src, W, H, dst, tmap, 1u, 0xFFFFFFFFu);
    LAUNCH_COUNT(L);
}

// ------------------------------------------------------------------------------------------------
// cv2.GaussianBlur on uint8: 8.8 fixed-point taps, horizontal then vertical, (acc + 2^15) >> 16
//   3 taps: [43,170,43]  (ksize 3, sigma 0.6)      5 taps: [16,64,96,64,16]  (ksize 5, sigma 0)
// ------------------------------------------------------------------------------------------------
// Separable in shared memory: one CTA = 64 x 16 outputs; the horizontal pass writes 8.8 fixed-point rows (u16 range
// fits: 255*256), the vertical pass forms the 16.16 sum.  Same integer arithmetic as OpenCV's fixedSmoothInvoker.
#define GU_TX 64
#define GU_TY 16
template <int NT>
__global__ void __launch_bounds__(256)
k_gauss_u8(const uint8_t* __restrict__ src, int W, int H, uint8_t* __restrict__ dst) {
    constexpr int R = NT / 2, INX = GU_TX + 2 * R, INY = GU_TY + 2 * R;
    __shared__ uint8_t tin[INY][INX + 4];
    __shared__ unsigned hrow[INY][GU_TX];
    const int taps3[3] = {43, 170, 43}, taps5[5] = {16, 64, 96, 64, 16};
    const int b = blockIdx.z, x0 = blockIdx.x * GU_TX, y0 = blockIdx.y * GU_TY;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const uint8_t* p = src + (size_t)b * W * H;
    {   // tile load: reflected column indices once per thread, row index once per row
        const int tx = threadIdx.x, ty = threadIdx.y;
        int gxs[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) gxs[k] = fpb_reflect101(x0 - R + tx + 32 * k, W);
        for (int r = ty; r < INY; r += 8) {
            const uint8_t* q = p + (size_t)fpb_reflect101(y0 - R + r, H) * W;
#pragma unroll
            for (int k = 0; k < 3; ++k) if (tx + 32 * k < INX) tin[r][tx + 32 * k] = q[gxs[k]];
        }
    }
    __syncthreads();
    for (int i = tid; i < INY * GU_TX; i += 256) {
        const int r = i / GU_TX, c = i - r * GU_TX;
        unsigned acc = 0;
#pragma unroll
        for (int k = 0; k < NT; ++k) acc += (unsigned)(NT == 3 ? taps3[k] : taps5[k]) * tin[r][c + k];
        hrow[r][c] = acc;
    }
    __syncthreads();
    for (int i = tid; i < GU_TY * GU_TX; i += 256) {
        const int r = i / GU_TX, c = i - r * GU_TX;
        const int gx = x0 + c, gy = y0 + r;
        if (gx >= W || gy >= H) continue;
        unsigned acc = 0;
#pragma unroll
        for (int k = 0; k < NT; ++k) acc += (unsigned)(NT == 3 ? taps3[k] : taps5[k]) * hrow[r + k][c];
        dst[(size_t)b * W * H + (size_t)gy * W + gx] = (uint8_t)((acc + 32768u) >> 16);
    }
}

void fpb_gauss_u8(FpbLaunch L, const uint8_t* src, int n, int W, int H, int ntaps, uint8_t* dst) {
    dim3 blk(32, 8), grid((W + GU_TX - 1) / GU_TX, (H + GU_TY - 1) / GU_TY, n);
    if (ntaps == 3) k_gauss_u8<3><<<grid, blk, 0, L.st>>>(src, W, H, dst);
    else k_gauss_u8<5><<<grid, blk, 0, L.st>>>(src, W, H, dst);
    LAUNCH_COUNT(L);
}
