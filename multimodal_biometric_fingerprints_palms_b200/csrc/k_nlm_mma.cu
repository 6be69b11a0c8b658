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
#include <cuda.h>
#include <string.h>
#include <stdlib.h>

#define MM_BH 16                       // pixel block
#define MM_BW 8
#define MM_CR (MM_BH + 20)             // 36 candidate rows
#define MM_CC (MM_BW + 20)             // 28 candidate columns
#define MM_NPR 32                      // candidate columns per row in N (padded)
#define MM_TR (MM_CR + 6)              // 42 tile rows
#define MM_TS 64                       // tile row stride (bytes) = TMA box width; tile column 0 = image column (x0 & ~15) - 16, so
                                       // that the box starts on a 16-byte boundary of the image row; candidate column 0's patch
                                       // (image column x0 - 13) then starts at tile column 3 + (x0 & 8)
#define MM_NW 529                      // weight table: indices 0..527 live, [528] = 0
#define MM_CHUNKS 9
#define MM_WORKERS 256
#define MM_THREADS (MM_WORKERS + 32)
#define MM_SSD_MAX 33791               // largest SSD with a non-zero weight: (528 << 6) - 1

#define MM_B_BYTES (MM_CR * MM_NPR * 64)                 // 73 728
#define MM_A_BYTES (128 * 64)                            //  8 192
#define MM_TILE_BYTES (MM_TR * MM_TS)                    //  2 688
#define MM_NQ_WORDS (MM_CR * MM_NPR)                     //  1 152
#define MM_SCR_BYTES (8 * 32 * 48)                       // 12 288: per warp, per lane 24 x u16 (aliases the row-sum plane)
#define MM_OFF_A MM_B_BYTES
#define MM_OFF_TILE (MM_OFF_A + MM_A_BYTES)
#define MM_OFF_NA (MM_OFF_TILE + 2 * MM_TILE_BYTES)              // two tile buffers: the next block's tile lands under this block's epilogue
#define MM_OFF_NQI (MM_OFF_NA + MM_NQ_WORDS * 4)
#define MM_OFF_LUT (MM_OFF_NQI + MM_NQ_WORDS * 4)
#define MM_OFF_SCR (MM_OFF_LUT + 2176)
#define MM_OFF_COMB (MM_OFF_SCR + MM_SCR_BYTES)
#define MM_OFF_BAR (MM_OFF_COMB + 128 * 8)
#define MM_SMEM_BYTES (MM_OFF_BAR + 64)                 // bars: full[2], empty[2], tile[2]

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
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");       // suspend-time hint (ns)
        if (spin > (1u << 20)) __trap();         // a lost arrival must fail loudly, never hang the GPU
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
#define MM_ORDER 0x654328170ull         // chunk of step i = nibble i

__device__ __forceinline__ void mm_tile_tma(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int x, int y, int z) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(MM_TILE_BYTES) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(dst), "l"((uint64_t)tmap), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
}

struct MmBlock { int b, x0, y0; };
__device__ __forceinline__ MmBlock mm_block(unsigned blk, unsigned bx_n, unsigned by_n) {      // n * blocks < 2^31 (checked by the launcher)
    MmBlock k;
    const unsigned per = bx_n * by_n, b = blk / per, rem = blk - b * per, by = rem / bx_n;
    k.b = (int)b; k.x0 = (int)(rem - by * bx_n) * MM_BW; k.y0 = (int)by * MM_BH;
    return k;
}

// 8-byte window of a tile row that starts at byte `col` (any alignment): three aligned words, two funnel shifts
__device__ __forceinline__ uint2 mm_window(const uint8_t* row, int col) {
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(row + (col & ~3));
    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
    const int sh = (col & 3) * 8;
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}

template <bool USE_TMA>
__global__ void __launch_bounds__(MM_THREADS, 2)
k_nlm_mma(const uint8_t* __restrict__ src, int W, int H, int n_img, uint8_t* __restrict__ dst, int fma_one,
          const __grid_constant__ CUtensorMap tmap, unsigned long long* __restrict__ prof) {
    // optional phase timing (FPB_NLM_PROF=1): clock64 deltas of worker warp 0, lane 0, summed over the CTA's blocks
    unsigned long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pc = 0;
#define MM_STAMP(k) do { if (prof && tid == 0) { const unsigned long long c_ = clock64(); pt[k] += c_ - pc; pc = c_; } } while (0)
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* sB = sm;
    uint8_t* sA = sm + MM_OFF_A;
    int* sNA = reinterpret_cast<int*>(sm + MM_OFF_NA);               // -(N(q) >> 1) per candidate [36][32]
    uint32_t* sNQI = reinterpret_cast<uint32_t*>(sm + MM_OFF_NQI);   // (N(q) & 1) << 8 | I(q)
    int* sLut = reinterpret_cast<int*>(sm + MM_OFF_LUT);
    uint32_t* sHs = reinterpret_cast<uint32_t*>(sm + MM_OFF_SCR);    // build phase: 7-tap row sums of squares [42][28]
    uint16_t* sScr = reinterpret_cast<uint16_t*>(sm + MM_OFF_SCR);   // epilogue: per warp, per lane 24 x u16
    uint32_t* sComb = reinterpret_cast<uint32_t*>(sm + MM_OFF_COMB);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + MM_OFF_BAR);   // full[2], empty[2], tile[2]
    __shared__ uint32_t tmem_base_sh;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full = mm_smem(&bars[0]), bar_empty = mm_smem(&bars[2]), bar_tile = mm_smem(&bars[4]);
    const int bx_n = (W + MM_BW - 1) / MM_BW, by_n = (H + MM_BH - 1) / MM_BH;
    const unsigned total = (unsigned)n_img * (unsigned)bx_n * (unsigned)by_n;

    // ---- one-time setup
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(mm_smem(&tmem_base_sh)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        if (lane == 0) {
            mm_bar_init(bar_full, 1); mm_bar_init(bar_full + 8, 1);
            mm_bar_init(bar_empty, 8); mm_bar_init(bar_empty + 8, 8);
            mm_bar_init(bar_tile, 1); mm_bar_init(bar_tile + 8, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
        // =========================================================== TMA + MMA issuer
        uint32_t it = 0, nb = 0;
        const uint64_t adesc = mm_desc(mm_smem(sA));
        const uint32_t b_addr = mm_smem(sB), tile_addr = mm_smem(sm + MM_OFF_TILE);
        if (USE_TMA && lane == 0 && blockIdx.x < total) {
            const MmBlock k = mm_block(blockIdx.x, bx_n, by_n);
            mm_tile_tma(tile_addr, &tmap, bar_tile, (k.x0 & ~15) - 16, k.y0 - 13, k.b);
        }
        for (unsigned blk = blockIdx.x; blk < total; blk += gridDim.x, ++nb) {
            MM_BAR_SYNC(1, MM_THREADS);                     // operands of this block are in shared memory
            MM_FENCE_AFTER();
            if (USE_TMA && lane == 0 && blk + gridDim.x < total) {      // the next block's tile lands under this block's epilogue
                const MmBlock k = mm_block(blk + gridDim.x, bx_n, by_n);
                mm_tile_tma(tile_addr + ((nb + 1) & 1u) * MM_TILE_BYTES, &tmap, bar_tile + 8 * ((nb + 1) & 1u), (k.x0 & ~15) - 16, k.y0 - 13, k.b);
            }
            for (int i = 0; i < MM_CHUNKS; ++i, ++it) {
                const uint32_t buf = it & 1u, use = it >> 1;
                mm_bar_wait(bar_empty + 8 * buf, (use & 1u) ^ 1u);      // the epilogue has drained this buffer
                MM_FENCE_AFTER();
                if (lane == 0) {
                    const int chunk = (int)((MM_ORDER >> (4 * i)) & 15ull);
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
        const uint32_t colmask_rev = __brev(((1u << 21) - 1u) << c_rel) >> 8;   // bit 23-j: candidate column cbase + j is in reach
        uint16_t* scr = sScr + (warp * 32 + lane) * 24;
        const uint32_t scr_top = mm_smem(scr) + 46, nqi_base = mm_smem(sNQI) + (cbase + 23) * 4, lut_base = mm_smem(sLut);
        const uint32_t tbase = tmem + ((uint32_t)(quad * 32) << 16) + cbase;
        // im2col: this warp's band of tile rows, this lane's candidate column
        const int Rs = warp * 5 + min(warp, 2), Re = Rs + (warp < 2 ? 6 : 5);
        const uint32_t b_lane = (uint32_t)((lane >> 3) * 512 + (lane & 7) * 16);
        uint32_t itw = 0, nb = 0;
        for (unsigned blk = blockIdx.x; blk < total; blk += gridDim.x, ++nb) {
            const MmBlock k = mm_block(blk, bx_n, by_n);
            const int x0 = k.x0, y0 = k.y0, b = k.b;
            const uint8_t* img = src + (size_t)b * W * H;
            if (prof && tid == 0) pc = clock64();
            uint8_t* tile = sm + MM_OFF_TILE + (USE_TMA ? (nb & 1u) * MM_TILE_BYTES : 0);
            const int ox = (x0 & ~15) - 16, tx0 = 3 + (x0 & 8);            // image column of tile column 0; tile column of candidate column 0
            // ---- tile: rows y0-13 .. y0+28, columns ox .. ox+63 (BORDER_REFLECT_101 like OpenCV's padded copy)
            if (USE_TMA) {
                mm_bar_wait(bar_tile + 8 * (nb & 1u), (nb >> 1) & 1u);
                if (ox < 0 || ox + MM_TS > W || y0 < 13 || y0 + 29 > H) {    // the box left the image: zeros arrived there
                    for (int i = tid; i < MM_TR * MM_TS; i += MM_WORKERS) {
                        const int r = i / MM_TS, c = i - r * MM_TS;
                        const int gy = y0 - 13 + r, gx = ox + c;
                        if ((unsigned)gy >= (unsigned)H || (unsigned)gx >= (unsigned)W)
                            tile[i] = img[(size_t)fpb_reflect101(gy, H) * W + fpb_reflect101(gx, W)];
                    }
                }
            } else {
                for (int i = tid; i < MM_TR * MM_TS; i += MM_WORKERS) {
                    const int r = i / MM_TS, c = i - r * MM_TS;
                    tile[i] = img[(size_t)fpb_reflect101(y0 - 13 + r, H) * W + fpb_reflect101(ox + c, W)];
                }
            }
            MM_BAR_SYNC(2, MM_WORKERS);
            MM_STAMP(0);
            // ---- im2col of the B operand.  Patch rows (2j, 2j+1) of a candidate are the 16 bytes of one K chunk: a lane walks
            //      down its band of tile rows with the previous row in registers and stores {row R-1, row R} for the three
            //      candidates that have them as patch rows (0,1), (2,3), (4,5), and row R alone as patch row 6.
            if (lane < MM_CC) {
                uint2 prev = make_uint2(0u, 0u);
                if (Rs > 0) prev = mm_window(tile + (Rs - 1) * MM_TS, tx0 + lane);
                uint8_t* bp0 = sB + b_lane;
                for (int R = Rs; R < Re; ++R) {
                    const uint2 cur = mm_window(tile + R * MM_TS, tx0 + lane);
                    const uint32_t hi7 = cur.y & 0x00FFFFFFu;
                    sHs[R * MM_CC + lane] = __dp4a(cur.x, cur.x, __dp4a(hi7, hi7, 0u));
                    uint8_t* q = bp0 + R * 2048;
#pragma unroll
                    for (int pr = 0; pr < 6; pr += 2) {
                        const int cyi = R - 1 - pr;                           // warp-uniform
                        if (cyi >= 0 && cyi < MM_CR)
                            *reinterpret_cast<uint4*>(q - (1 + pr) * 2048 + (pr >> 1) * 128) = make_uint4(prev.x, prev.y, cur.x, cur.y);
                    }
                    if (R >= 6) *reinterpret_cast<uint2*>(q - 6 * 2048 + 3 * 128) = cur;
                    prev = cur;
                }
            }
            // ---- A operand: the 128 pixels' own patches with byte 7 of every patch row cleared
            for (int i = tid; i < 128 * 7; i += MM_WORKERS) {
                const int m = i & 127, pr = i >> 7;
                const int r = ((m >> 6) << 3) + ((m & 31) >> 2), c = (((m >> 5) & 1) << 2) + (m & 3);
                uint2 v = mm_window(tile + (r + 10 + pr) * MM_TS, tx0 + c + 10);
                v.y &= 0x00FFFFFFu;
                *reinterpret_cast<uint2*>(sA + mm_op_off(m, pr)) = v;
            }
            MM_BAR_SYNC(2, MM_WORKERS);
            MM_STAMP(1);
            // ---- N(q) = sum of 7 row sums, sliding down six candidate rows per thread; -(N >> 1) for the sign test,
            //      (N & 1) << 8 | I(q) for the survivors
            if (tid < 6 * MM_CC) {
                const int cxi = tid % MM_CC, c0 = (tid / MM_CC) * 6;
                uint32_t nq = 0;
#pragma unroll
                for (int pr = 0; pr < 7; ++pr) nq += sHs[(c0 + pr) * MM_CC + cxi];
#pragma unroll
                for (int t = 0; t < 6; ++t) {
                    const int cyi = c0 + t;
                    sNA[cyi * MM_NPR + cxi] = -(int)(nq >> 1);
                    sNQI[cyi * MM_NPR + cxi] = ((nq & 1u) << 8) | tile[(cyi + 3) * MM_TS + tx0 + cxi + 3];
                    if (t < 5) nq += sHs[(cyi + 7) * MM_CC + cxi] - sHs[cyi * MM_CC + cxi];
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // operand stores -> visible to the MMA (async proxy)
            MM_BAR_SYNC(1, MM_THREADS);
            MM_STAMP(2);
            // ---- per-lane constants of this block: N(p) from the two planes, b(p) = floor((N(p) - 33791) / 2)
            const int pi = (r_abs + 10) * MM_NPR + cbase + c_rel + 10;
            const int np = -2 * sNA[pi] + (int)(sNQI[pi] >> 8);
            const int bp = (np - MM_SSD_MAX) >> 1;
            const int cp = np - 2 * bp;                                     // 33791 or 33792
            const int nb_ok = -bp;
            unsigned sw = 0, swp = 0;
            for (int i = 0; i < MM_CHUNKS; ++i, ++itw) {
                const uint32_t buf = itw & 1u, use = itw >> 1;
                const int crow0 = 4 * (int)((MM_ORDER >> (4 * i)) & 15ull);
                { const unsigned long long w0_ = (prof && tid == 0) ? clock64() : 0ull;
                mm_bar_wait(bar_full + 8 * buf, use & 1u);
                if (prof && tid == 0) pt[6] += clock64() - w0_; }
                MM_FENCE_AFTER();
                if (crow0 >= rbase && crow0 < rbase + 28) {
                    // two candidate rows per warp and chunk: both rows' sign tests first, then the accumulator buffer goes back to the
                    // MMA warp, then the survivors (a latency-bound per-lane loop) - they stay off the buffer hand-over's critical path
                    uint32_t g[24]; int4 na[6]; uint32_t pk[12], pk2[12]; uint32_t pm, pm2;
                    const int cy_a = crow0 + 2 * half;
#define MM_ROW_LOAD(cyi_, rr_) do { \
                        const uint32_t taddr = tbase + buf * 128 + (rr_) * 32; \
                        MM_LD16(taddr, g, 0); MM_LD8(taddr + 16, g, 16); \
                        const int4* nap = reinterpret_cast<const int4*>(sNA + (cyi_) * MM_NPR + cbase); \
                        _Pragma("unroll") for (int q = 0; q < 6; ++q) na[q] = nap[q]; } while (0)
                    // e = G - (N(q) >> 1) - b(p): two of three adds are written as multiply-adds by a run-time 1 (FMA pipe), the funnel
                    // shifts that collect the sign bits and the remaining three-input adds go to the ALU pipe; three independent
                    // sign chains keep the shifts from serialising
#define MM_ROW_CORE(cyi_, pm, pk) do { \
                        const int nbr = ((unsigned)((cyi_) - r_abs) <= 20u) ? nb_ok : -(1 << 30); \
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); \
                        uint32_t s0 = 0, s1 = 0, s2 = 0; int e[24]; \
                        _Pragma("unroll") for (int q = 0; q < 6; ++q) { \
                            const int a4[4] = {na[q].x, na[q].y, na[q].z, na[q].w}; \
                            _Pragma("unroll") for (int u = 0; u < 4; ++u) { \
                                const int j = 4 * q + u; \
                                if (j % 3 == 0) e[j] = (int)g[j] + a4[u] + nbr; \
                                else e[j] = ((int)g[j] * fma_one + a4[u]) * fma_one + nbr; \
                                if (j < 8) s0 = __funnelshift_l((uint32_t)e[j], s0, 1); \
                                else if (j < 16) s1 = __funnelshift_l((uint32_t)e[j], s1, 1); \
                                else s2 = __funnelshift_l((uint32_t)e[j], s2, 1); \
                            } } \
                        pm = ~((s0 << 16) | (s1 << 8) | s2) & colmask_rev;       /* bit 23-j: column j survives for this lane */ \
                        _Pragma("unroll") for (int q = 0; q < 12; ++q) pk[q] = __byte_perm(e[2 * q], e[2 * q + 1], 0x5410); } while (0)
#define MM_ROW_SURVIVORS(cyi_, pm, pk) do { \
                        if (__any_sync(0xffffffffu, pm != 0u)) { \
                            uint4* sp = reinterpret_cast<uint4*>(scr); \
                            sp[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]); sp[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]); \
                            sp[2] = make_uint4(pk[8], pk[9], pk[10], pk[11]); \
                            __syncwarp(); \
                            const uint32_t nq_row = nqi_base + (uint32_t)(cyi_) * (MM_NPR * 4); \
                            while (pm) { \
                                uint32_t bit, ev, v, w;                            /* column j = 23 - bit */ \
                                asm("bfind.u32 %0, %1;" : "=r"(bit) : "r"(pm)); \
                                pm ^= 1u << bit; \
                                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(ev) : "r"(scr_top - 2u * bit) : "memory"); \
                                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(nq_row - 4u * bit) : "memory"); \
                                const int ssd = cp + (int)(v >> 8) - 2 * (int)ev;      /* = N(p) + N(q) - 2 G exactly */ \
                                const int idx = min(ssd >> 6, MM_NW - 1); \
                                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(lut_base + 4u * (uint32_t)idx) : "memory"); \
                                sw += w; swp += w * (v & 255u); \
                            } \
                            __syncwarp(); \
                        } } while (0)
                    MM_ROW_LOAD(cy_a, 2 * half);
                    MM_ROW_CORE(cy_a, pm, pk);
                    MM_ROW_LOAD(cy_a + 1, 2 * half + 1);
                    MM_ROW_CORE(cy_a + 1, pm2, pk2);
                    MM_FENCE_BEFORE();
                    __syncwarp();
                    if (lane == 0) mm_bar_arrive(bar_empty + 8 * buf);      // survivors are finished off the accumulator's critical path
                    MM_ROW_SURVIVORS(cy_a, pm, pk);
                    MM_ROW_SURVIVORS(cy_a + 1, pm2, pk2);
                } else {
                    MM_FENCE_BEFORE();
                    __syncwarp();
                    if (lane == 0) mm_bar_arrive(bar_empty + 8 * buf);
                }
            }
            MM_STAMP(3);
            // ---- the two halves of a quadrant hold partial sums of the same 32 pixels
            if (half == 1) { sComb[(quad * 32 + lane) * 2] = sw; sComb[(quad * 32 + lane) * 2 + 1] = swp; }
            MM_BAR_SYNC(2, MM_WORKERS);
            if (half == 0) {
                sw += sComb[(quad * 32 + lane) * 2]; swp += sComb[(quad * 32 + lane) * 2 + 1];
                const int gy = y0 + r_abs, gx = x0 + cbase + c_rel;
                if (gy < H && gx < W) dst[(size_t)b * W * H + (size_t)gy * W + gx] = (uint8_t)min((swp + sw / 2u) / sw, 255u);
            }
            MM_BAR_SYNC(2, MM_WORKERS);                     // N planes / scratch / combine buffer are rebuilt for the next block
            MM_STAMP(4);
        }
    }
    if (prof && tid == 0) for (int q = 0; q < 8; ++q) atomicAdd(&prof[q], pt[q]);
    MM_FENCE_BEFORE();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256));
}

// cuTensorMapEncodeTiled through the runtime's driver-entry-point query (no link-time dependency on libcuda)
typedef CUresult (*MmEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static MmEncodeTiledFn mm_encode_tiled() {
    static MmEncodeTiledFn fn = nullptr; static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess) fn = (MmEncodeTiledFn)ptr;
        (void)cudaGetLastError();
    }
    return fn;
}

bool fpb_nlm_mma(FpbLaunch L, const uint8_t* src, int n, int W, int H, uint8_t* dst) {
    static int sms[64]; int dev = 0; cudaGetDevice(&dev);
    if (!sms[dev & 63]) cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    FPB_OPT_IN_SMEM(k_nlm_mma<true>, MM_SMEM_BYTES);
    FPB_OPT_IN_SMEM(k_nlm_mma<false>, MM_SMEM_BYTES);
    const long long total = (long long)n * ((W + MM_BW - 1) / MM_BW) * ((H + MM_BH - 1) / MM_BH);
    if (total >= (1ll << 31)) return false;
    long long grid = 2LL * sms[dev & 63];                // persistent: two CTAs per SM (110 KB of shared memory, 256 TMEM columns each)
    if (grid > total) grid = total;
    // the image batch as a 3-D u8 tensor (W, H, n); one TMA box = one 48 x 42 tile (out-of-image bytes arrive as zeros and are
    // patched to BORDER_REFLECT_101 by the threads).  TMA needs a 16-byte aligned base and row pitch; other widths load the tile
    // with plain loads.
    CUtensorMap tmap; memset(&tmap, 0, sizeof(tmap));
    bool use_tma = false;
    static const bool no_tma = getenv("FPB_NO_TMA") != nullptr;
    MmEncodeTiledFn enc = no_tma ? nullptr : mm_encode_tiled();
    if (enc && (W % 16) == 0 && (((uintptr_t)src) % 16) == 0) {
        const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
        const cuuint64_t gstr[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
        const cuuint32_t box[3] = {MM_TS, MM_TR, 1}, estr[3] = {1, 1, 1};
        use_tma = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    static const bool want_prof = getenv("FPB_NLM_PROF") != nullptr;
    unsigned long long* d_prof = nullptr;
    if (want_prof) { cudaMalloc(&d_prof, 64); cudaMemsetAsync(d_prof, 0, 64, L.st); }
    if (use_tma) k_nlm_mma<true><<<(unsigned)grid, MM_THREADS, MM_SMEM_BYTES, L.st>>>(src, W, H, n, dst, 1, tmap, d_prof);
    else k_nlm_mma<false><<<(unsigned)grid, MM_THREADS, MM_SMEM_BYTES, L.st>>>(src, W, H, n, dst, 1, tmap, d_prof);
    if (want_prof) {                                       // diagnostics only: synchronous
        unsigned long long hp[8]; cudaStreamSynchronize(L.st);
        cudaMemcpy(hp, d_prof, 64, cudaMemcpyDeviceToHost); cudaFree(d_prof);
        const double nb = (double)total;
        fprintf(stderr, "[k_nlm_mma] cycles/block (worker warp 0): tile %.0f  build %.0f  nq+ready %.0f  chunks %.0f (of which waiting on full %.0f)  combine %.0f\n",
                hp[0] / nb, hp[1] / nb, hp[2] / nb, hp[3] / nb, hp[6] / nb, hp[4] / nb);
    }
    LAUNCH_COUNT(L);
    return true;
}
