// K2 on the tensor cores: cv2.fastNlMeansDenoising(h=10, template 7, search 21) of
// /root/reference/src/preprocessing/fingerprint_preprocess.py:36, bit-exact (same integers as k_nlm of k_front.cu).
//
//   SSD(p,q) = sum_7x7 (I[p+t] - I[q+t])^2 = N(p) + N(q) - 2 G(p,q),   G(p,q) = <patch(p), patch(q)>,  N(x) = G(x,x)
//
// G over a block of pixels and its candidate window is a GEMM of 49-byte patch vectors: exact in u8 x u8 -> s32.
// One CTA works on blocks of 16 x 8 output pixels (M = 128 rows of the MMA).  The candidates of a block are the
// (16+20) x (8+20) = 36 x 28 positions its pixels can reach; they form N = 36 rows x 32 (28 + 4 pad) columns = 1152,
// taken in nine chunks of 128 (four candidate rows).  K = 7 patch rows x 8 bytes (the 8th byte of every patch row is
// zero in the A operand, so the B operand can keep whatever image byte follows) + 8 zero bytes = 64.
//
//   tile (42 x 48 bytes, reflect-101) -> im2col by the threads into the canonical no-swizzle K-major UMMA layout
//   (8-row x 16-byte core matrices: byte (row, k) at (row%8)*16 + (row/8)*512 + (k/16)*128 + k%16; checked against a CPU
//   GEMM by tools/ubench) -> tcgen05.mma.cta_group::1.kind::i8 M128 N128 K32 x 2 per chunk, accumulators in TMEM
//   (2 x 128 columns per CTA: the MMA of chunk i+1 runs under the epilogue of chunk i; two CTAs per SM share the 512
//   columns) -> eight epilogue warps: warp (quadrant, half) owns the 32 TMEM lanes of an 8 x 4 pixel tile and two of the
//   chunk's four candidate rows: tcgen05.ld of the 24 columns its tile can reach,
//       e = G - (N(q) >> 1) - b(p),   b(p) = floor((N(p) - 33791) / 2),       e >= 0  <=  weight != 0
//   one add + one funnel shift per (pixel, candidate) collect the sign bits; the rare survivors (2 % of the pairs on
//   contrast-stretched prints) are finished per lane: SSD = N(p) - 2 b(p) + (N(q) & 1) - 2 e exactly, table look-up,
//   accumulate.  Per (pixel, candidate) that is ~2.7 issue slots instead of the ~14 of the scalar formulation.
//
// A ninth warp allocates the tensor memory and issues the MMAs (one elected lane); mbarriers `full` (tcgen05.commit)
// and `empty` (one arrival per epilogue warp) hand the two accumulator buffers back and forth.
#include "fpb_kernels.h"

#define MM_BH 16                       // pixel block
#define MM_BW 8
#define MM_CR (MM_BH + 20)             // 36 candidate rows
#define MM_CC (MM_BW + 20)             // 28 candidate columns
#define MM_NPR 32                      // candidate columns per row in N (padded)
#define MM_TR (MM_CR + 6)              // 42 tile rows
#define MM_TS 48                       // tile row stride (bytes): 35 used + room for aligned 12-byte windows
#define MM_NW 529                      // weight table: indices 0..527 live, [528] = 0
#define MM_CHUNKS 9
#define MM_WORKERS 256
#define MM_THREADS (MM_WORKERS + 32)
#define MM_SSD_MAX 33791               // largest SSD with a non-zero weight: (528 << 6) - 1

#define MM_B_BYTES (MM_CR * MM_NPR * 64)                 // 73 728
#define MM_A_BYTES (128 * 64)                            //  8 192
#define MM_TILE_BYTES (MM_TR * MM_TS)                    //  2 016
#define MM_NQ_WORDS (MM_CR * MM_NPR)                     //  1 152
#define MM_SCR_BYTES (8 * 32 * 48)                       // 12 288: per warp, per lane 24 x u16 (aliases the row-sum plane)
#define MM_OFF_A MM_B_BYTES
#define MM_OFF_TILE (MM_OFF_A + MM_A_BYTES)
#define MM_OFF_NA (MM_OFF_TILE + 2048)
#define MM_OFF_NQI (MM_OFF_NA + MM_NQ_WORDS * 4)
#define MM_OFF_LUT (MM_OFF_NQI + MM_NQ_WORDS * 4)
#define MM_OFF_SCR (MM_OFF_LUT + 2176)
#define MM_OFF_COMB (MM_OFF_SCR + MM_SCR_BYTES)
#define MM_OFF_BAR (MM_OFF_COMB + 128 * 8)
#define MM_SMEM_BYTES (MM_OFF_BAR + 64)

__constant__ int c_nlm_w_mma[MM_NW];

void fpb_upload_nlm_table_mma(const int* tab, cudaStream_t st) {
    cudaMemcpyToSymbolAsync(c_nlm_w_mma, tab, sizeof(int) * MM_NW, 0, cudaMemcpyHostToDevice, st);
}

__device__ __forceinline__ uint32_t mm_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle, version 1 (sm_100): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | 1 << 46
__device__ __forceinline__ uint64_t mm_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46);
}
// kind::i8: D = s32 (2 << 4), A = B = u8 (0), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
#define MM_IDESC ((2u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24))

__device__ __forceinline__ void mm_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(MM_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mm_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mm_bar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mm_bar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mm_bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();         // a lost arrival must fail loudly, never hang the GPU
    }
}
#define MM_FENCE_BEFORE() asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory")
#define MM_FENCE_AFTER() asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory")
#define MM_BAR_SYNC(id, n) asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory")

#define MM_LD16(taddr, v, o) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
                 : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7]), \
                   "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]), "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15]) \
                 : "r"(taddr) : "memory")
#define MM_LD8(taddr, v, o) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" \
                 : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7]) \
                 : "r"(taddr) : "memory")

// byte offset of (row n, patch row pr) inside an operand: 8 bytes go to k = pr*8 .. pr*8+7
__device__ __forceinline__ uint32_t mm_op_off(int n, int pr) {
    return (uint32_t)((n >> 3) * 512 + (pr >> 1) * 128 + (n & 7) * 16 + (pr & 1) * 8);
}

// chunk order: 0 and 7 (then 1 and 8) are each used by one half of the epilogue warps only - taking them in pairs keeps
// all eight warps busy while both accumulator buffers are in flight
__constant__ int c_mm_order[MM_CHUNKS] = {0, 7, 1, 8, 2, 3, 4, 5, 6};

__global__ void __launch_bounds__(MM_THREADS, 2)
k_nlm_mma(const uint8_t* __restrict__ src, int W, int H, int n_img, uint8_t* __restrict__ dst, int fma_one) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* sB = sm;
    uint8_t* sA = sm + MM_OFF_A;
    uint8_t* tile = sm + MM_OFF_TILE;
    int* sNA = reinterpret_cast<int*>(sm + MM_OFF_NA);          // -(N(q) >> 1) per candidate [36][32]
    uint32_t* sNQI = reinterpret_cast<uint32_t*>(sm + MM_OFF_NQI);   // N(q) << 8 | I(q)
    int* sLut = reinterpret_cast<int*>(sm + MM_OFF_LUT);
    uint32_t* sHs = reinterpret_cast<uint32_t*>(sm + MM_OFF_SCR);    // build phase: 7-tap row sums of squares [42][28]
    uint16_t* sScr = reinterpret_cast<uint16_t*>(sm + MM_OFF_SCR);   // epilogue: per warp, per lane 24 x u16
    uint32_t* sComb = reinterpret_cast<uint32_t*>(sm + MM_OFF_COMB);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + MM_OFF_BAR);   // full[2], empty[2]
    __shared__ uint32_t tmem_base_sh;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full = mm_smem(&bars[0]), bar_empty = mm_smem(&bars[2]);
    const int bx_n = (W + MM_BW - 1) / MM_BW, by_n = (H + MM_BH - 1) / MM_BH;
    const long long total = (long long)n_img * bx_n * by_n;

    // ---- one-time setup
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(mm_smem(&tmem_base_sh)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        if (lane == 0) {
            mm_bar_init(bar_full, 1); mm_bar_init(bar_full + 8, 1);
            mm_bar_init(bar_empty, 8); mm_bar_init(bar_empty + 8, 8);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    } else {
        for (int i = tid; i < MM_NW; i += MM_WORKERS) sLut[i] = c_nlm_w_mma[i];
        for (int i = tid; i < MM_A_BYTES / 4; i += MM_WORKERS) reinterpret_cast<uint32_t*>(sA)[i] = 0u;   // k 56..63 stay zero
    }
    MM_FENCE_BEFORE();
    __syncthreads();
    MM_FENCE_AFTER();
    const uint32_t tmem = tmem_base_sh;

    if (warp == 8) {
        // =========================================================== MMA issuer
        uint32_t it = 0;
        const uint64_t adesc = mm_desc(mm_smem(sA));
        const uint32_t b_addr = mm_smem(sB);
        for (long long blk = blockIdx.x; blk < total; blk += gridDim.x) {
            MM_BAR_SYNC(1, MM_THREADS);                     // operands of this block are in shared memory
            MM_FENCE_AFTER();
            for (int i = 0; i < MM_CHUNKS; ++i, ++it) {
                const uint32_t buf = it & 1u, use = it >> 1;
                mm_bar_wait(bar_empty + 8 * buf, (use & 1u) ^ 1u);      // the epilogue has drained this buffer
                MM_FENCE_AFTER();
                if (lane == 0) {
                    const int chunk = c_mm_order[i];
                    const uint64_t bdesc = mm_desc(b_addr + (uint32_t)chunk * (128 / 8) * 512);
                    mm_mma(tmem + buf * 128, adesc, bdesc, 0u);
                    mm_mma(tmem + buf * 128, adesc + 16, bdesc + 16, 1u);     // + 256 bytes: k 32..63
                    mm_commit(bar_full + 8 * buf);
                }
                __syncwarp();
            }
        }
    } else {
        // =========================================================== tile / im2col / epilogue warps
        const int quad = warp & 3, half = warp >> 2;
        const int rbase = (quad >> 1) * 8, cbase = (quad & 1) * 4;
        const int r_abs = rbase + (lane >> 2), c_rel = lane & 3;         // this lane's pixel inside the block / its tile
        const uint32_t colmask = (((1u << 21) - 1u) << c_rel);           // bit j: candidate column cbase + j is in reach
        const uint32_t colmask_rev = __brev(colmask) >> 8;               // the same with bit (23 - j): the order the signs arrive in
        uint16_t* scr = sScr + (warp * 32 + lane) * 24;
        uint32_t itw = 0;
        for (long long blk = blockIdx.x; blk < total; blk += gridDim.x) {
            const int b = (int)(blk / ((long long)bx_n * by_n));
            const int rem = (int)(blk - (long long)b * bx_n * by_n);
            const int by = rem / bx_n, bx = rem - by * bx_n;
            const int x0 = bx * MM_BW, y0 = by * MM_BH;
            const uint8_t* img = src + (size_t)b * W * H;
            // ---- tile: rows y0-13 .. y0+28, columns x0-13 .. x0+34 (BORDER_REFLECT_101 like OpenCV's padded copy)
            for (int i = tid; i < MM_TR * MM_TS; i += MM_WORKERS) {
                const int r = i / MM_TS, c = i - r * MM_TS;
                tile[i] = img[(size_t)fpb_reflect101(y0 - 13 + r, H) * W + fpb_reflect101(x0 - 13 + c, W)];
            }
            MM_BAR_SYNC(2, MM_WORKERS);
            // ---- im2col: tile row R, candidate column cxi -> the 8-byte patch row, stored for the up to 7 candidate rows
            //      (and pixel rows) it belongs to; 7-tap sum of squares for N(q)
            for (int R = warp; R < MM_TR; R += 8) {
                if (lane < MM_CC) {
                    const int cxi = lane;
                    const uint32_t* wp = reinterpret_cast<const uint32_t*>(tile + R * MM_TS + (cxi & ~3));
                    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
                    const int sh = (cxi & 3) * 8;
                    const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
                    const uint32_t hi7 = hi & 0x00FFFFFFu;
                    sHs[R * MM_CC + cxi] = __dp4a(lo, lo, __dp4a(hi7, hi7, 0u));
                    const bool pixcol = cxi >= 10 && cxi < 10 + MM_BW;
#pragma unroll
                    for (int pr = 0; pr < 7; ++pr) {
                        const int cyi = R - pr;
                        if (cyi < 0 || cyi >= MM_CR) continue;
                        *reinterpret_cast<uint2*>(sB + mm_op_off(cyi * MM_NPR + cxi, pr)) = make_uint2(lo, hi);
                        if (pixcol && cyi >= 10 && cyi < 10 + MM_BH) {
                            const int r = cyi - 10, c = cxi - 10;
                            const int m = ((r >> 3) * 2 + (c >> 2)) * 32 + (r & 7) * 4 + (c & 3);
                            *reinterpret_cast<uint2*>(sA + mm_op_off(m, pr)) = make_uint2(lo, hi7);
                        }
                    }
                }
            }
            MM_BAR_SYNC(2, MM_WORKERS);
            // ---- N(q) = sum of 7 row sums; -(N >> 1) for the sign test, N << 8 | I(q) for the survivors
            for (int i = tid; i < MM_CR * MM_CC; i += MM_WORKERS) {
                const int cyi = i / MM_CC, cxi = i - cyi * MM_CC;
                uint32_t nq = 0;
#pragma unroll
                for (int pr = 0; pr < 7; ++pr) nq += sHs[(cyi + pr) * MM_CC + cxi];
                sNA[cyi * MM_NPR + cxi] = -(int)(nq >> 1);
                sNQI[cyi * MM_NPR + cxi] = (nq << 8) | tile[(cyi + 3) * MM_TS + cxi + 3];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // operand stores -> visible to the MMA (async proxy)
            MM_BAR_SYNC(1, MM_THREADS);
            // ---- per-lane constants of this block
            const int np = (int)(sNQI[(r_abs + 10) * MM_NPR + cbase + c_rel + 10] >> 8);
            const int bp = (np - MM_SSD_MAX) >> 1;                          // floor
            const int cp = np - 2 * bp;                                     // 33791 or 33792
            unsigned sw = 0, swp = 0;
            for (int i = 0; i < MM_CHUNKS; ++i, ++itw) {
                const uint32_t buf = itw & 1u, use = itw >> 1;
                const int chunk = c_mm_order[i];
                mm_bar_wait(bar_full + 8 * buf, use & 1u);
                MM_FENCE_AFTER();
                const int crow0 = 4 * chunk;
                if (crow0 >= rbase && crow0 < rbase + 28) {
#pragma unroll 1
                    for (int rr = 2 * half; rr < 2 * half + 2; ++rr) {
                        const int cyi = crow0 + rr;
                        const int nbr = ((unsigned)(cyi - r_abs) <= 20u) ? -bp : -(1 << 30);
                        uint32_t g[24];
                        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + buf * 128 + rr * 32 + cbase;
                        MM_LD16(taddr, g, 0);
                        MM_LD8(taddr + 16, g, 16);
                        const int4* nap = reinterpret_cast<const int4*>(sNA + cyi * MM_NPR + cbase);
                        int4 na[6];
#pragma unroll
                        for (int k = 0; k < 6; ++k) na[k] = nap[k];
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        uint32_t sgn = 0;
                        int e[24];
#pragma unroll
                        for (int k = 0; k < 6; ++k) {
                            // half of the adds are written as multiply-adds by a run-time 1: they issue on the FMA pipe, the
                            // funnel shifts and the other adds on the ALU pipe (both pipes take one warp instruction per 2 cycles)
                            e[4 * k + 0] = (int)g[4 * k + 0] + na[k].x + nbr;
                            e[4 * k + 1] = ((int)g[4 * k + 1] * fma_one + na[k].y) * fma_one + nbr;
                            e[4 * k + 2] = ((int)g[4 * k + 2] * fma_one + na[k].z) * fma_one + nbr;
                            e[4 * k + 3] = (int)g[4 * k + 3] + na[k].w + nbr;
#pragma unroll
                            for (int u = 0; u < 4; ++u) sgn = __funnelshift_l((uint32_t)e[4 * k + u], sgn, 1);
                        }
                        uint32_t pm = ~sgn & colmask_rev;                   // bit 23-j: column j survives for this lane
                        if (__any_sync(0xffffffffu, pm != 0u)) {
                            uint4* sp = reinterpret_cast<uint4*>(scr);
#pragma unroll
                            for (int k = 0; k < 3; ++k)
                                sp[k] = make_uint4(__byte_perm(e[8 * k + 0], e[8 * k + 1], 0x5410), __byte_perm(e[8 * k + 2], e[8 * k + 3], 0x5410),
                                                   __byte_perm(e[8 * k + 4], e[8 * k + 5], 0x5410), __byte_perm(e[8 * k + 6], e[8 * k + 7], 0x5410));
                            __syncwarp();
                            const uint32_t* nq = sNQI + cyi * MM_NPR + cbase;
                            while (pm) {
                                const int bit = 31 - __clz(pm);
                                pm ^= 1u << bit;
                                const int j = 23 - bit;
                                const int ev = (int)scr[j];
                                const uint32_t v = nq[j];
                                const int ssd = cp + (int)((v >> 8) & 1u) - 2 * ev;          // = N(p) + N(q) - 2 G exactly
                                const int idx = min(ssd >> 6, MM_NW - 1);
                                const unsigned w = (unsigned)sLut[idx];
                                sw += w; swp += w * (v & 255u);
                            }
                            __syncwarp();
                        }
                    }
                }
                MM_FENCE_BEFORE();
                __syncwarp();
                if (lane == 0) mm_bar_arrive(bar_empty + 8 * buf);
            }
            // ---- the two halves of a quadrant hold partial sums of the same 32 pixels
            if (half == 1) { sComb[(quad * 32 + lane) * 2] = sw; sComb[(quad * 32 + lane) * 2 + 1] = swp; }
            MM_BAR_SYNC(2, MM_WORKERS);
            if (half == 0) {
                sw += sComb[(quad * 32 + lane) * 2]; swp += sComb[(quad * 32 + lane) * 2 + 1];
                const int gy = y0 + r_abs, gx = x0 + cbase + c_rel;
                if (gy < H && gx < W) dst[(size_t)b * W * H + (size_t)gy * W + gx] = (uint8_t)min((swp + sw / 2u) / sw, 255u);
            }
            MM_BAR_SYNC(2, MM_WORKERS);                     // tile / N planes / scratch are rebuilt for the next block
        }
    }
    MM_FENCE_BEFORE();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256));
}

bool fpb_nlm_mma(FpbLaunch L, const uint8_t* src, int n, int W, int H, uint8_t* dst) {
    static int sms[64]; int dev = 0; cudaGetDevice(&dev);
    if (!sms[dev & 63]) cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    FPB_OPT_IN_SMEM(k_nlm_mma, MM_SMEM_BYTES);
    const long long total = (long long)n * ((W + MM_BW - 1) / MM_BW) * ((H + MM_BH - 1) / MM_BH);
    long long grid = 2LL * sms[dev & 63];                // persistent: two CTAs per SM (107 KB of shared memory, 256 TMEM columns each)
    if (grid > total) grid = total;
    k_nlm_mma<<<(unsigned)grid, MM_THREADS, MM_SMEM_BYTES, L.st>>>(src, W, H, n, dst, 1);
    LAUNCH_COUNT(L);
    return true;
}
