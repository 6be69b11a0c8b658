// K2 on the tensor cores (formulation v4): cv2.fastNlMeansDenoising(h=10, template 7, search 21) of
// /root/reference/src/preprocessing/fingerprint_preprocess.py:36, bit-exact (same integers as k_nlm of k_front.cu).
//
//   SSD(p,q) = sum_7x7 (a - b)^2 = N(p) - D(q) - 2 <a - 128, b>,     N(p) = sum a^2,   D(q) = sum b (256 - b) >= 0
//
// <a - 128, b> over a block of pixels and its candidate window is a GEMM of 49-byte patch vectors, exact in s8 x u8 -> s32.
// The per-candidate term is folded INTO the GEMM: the operands' K = 64 holds the 7 patch rows padded to 8 bytes (56) + 8
// tail bytes, and the 15 spare bytes carry, on the candidate side, the digits of Q(q) = ceil(D(q)/2) in units of 127
// (pad byte of every patch row = one shared digit, six tail digits, one remainder) against the constants 127 / 1 on the
// pixel side.  What the MMA leaves in tensor memory is therefore
//       e(p,q) = <a - 128, b> + Q(q),          SSD = N(p) + (D(q) & 1) - 2 e,
// and "weight != 0" (SSD <= 33791) implies e >= thr(p) = ceil((N(p) - 33791)/2), a per-lane constant: the test of a
// (pixel, candidate) pair is half a VIMNMX3 on the value as it arrives from tcgen05.ld - no per-pair add, no candidate-side
// table in the inner loop.  The 2 % of the pairs that pass are finished per lane (exact SSD, weight table, accumulate).
//
// One persistent CTA per SM, three roles (warp-specialised, mbarrier hand-offs, everything double-buffered):
//   producers (6 warps)  TMA tile (42 x 64 bytes, reflect-101 patched) -> im2col of the B operand for the block's 36 x 28
//                        candidates into the canonical no-swizzle K-major UMMA layout, 7-row sliding sums of sum b^2 and
//                        sum b -> D(q), digits, (D & 1) << 8 | I(q) table; A operand (128 pixels of a 16 x 8 block)
//   MMA warp             5 chunks of 8 candidate rows (N = 256; the last N = 128): 2 x tcgen05.mma kind::i8 (K = 32 each)
//                        into one of two 256-column accumulators, tcgen05.commit -> `tfull`
//   epilogue (8 warps)   warp (quad, parity): the 32 TMEM lanes of an 8 x 4 pixel tile, the candidate rows of its parity:
//                        tcgen05.ld of the 24 columns the tile can reach, max over them, vote, survivors
// The build of block i+1 runs under the epilogue of block i; the MMA of chunk c+1 under the epilogue of chunk c.
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
#define MM_NCHUNK 5                    // 8 candidate rows per chunk (the last chunk holds rows 32..35)
#ifndef MM_EPI_WARPS
#define MM_EPI_WARPS 12
#endif
#define MM_SUBS (MM_EPI_WARPS / 4)       // epilogue warps per TMEM lane quadrant: warp (quad, sub) takes the candidate rows cy % MM_SUBS == sub
#define MM_PROD_WARPS 6
#define MM_PROD (MM_PROD_WARPS * 32)
#define MM_THREADS (32 * (MM_EPI_WARPS + 1 + MM_PROD_WARPS))
#define MM_SSD_MAX 33791               // largest SSD with a non-zero weight: (528 << 6) - 1

#define MM_B_BYTES (MM_CR * MM_NPR * 64)                 // 73 728
#define MM_A_BYTES (128 * 64)                            //  8 192
#define MM_TILE_BYTES (MM_TR * MM_TS)                    //  2 688
#define MM_NQ_WORDS (MM_CR * MM_NPR)                     //  1 152
#define MM_HS_WORDS (MM_TR * MM_CC)                      //  1 176
#define MM_SCR_STRIDE 96                                 // per-lane scratch: 24 words
#define MM_OFF_B 0
#define MM_OFF_A (2 * MM_B_BYTES)
#define MM_OFF_TILE (MM_OFF_A + 2 * MM_A_BYTES)
#define MM_OFF_NQI (MM_OFF_TILE + 2 * MM_TILE_BYTES)
#define MM_OFF_NAP (MM_OFF_NQI + 2 * MM_NQ_WORDS * 4)
#define MM_OFF_HQ (MM_OFF_NAP + 2 * 128 * 4)
#define MM_OFF_HB (MM_OFF_HQ + MM_HS_WORDS * 4)
#define MM_OFF_LUT (MM_OFF_HB + MM_HS_WORDS * 4)
#define MM_OFF_SCR (MM_OFF_LUT + 2176)
#define MM_OFF_COMB (MM_OFF_SCR + MM_EPI_WARPS * 32 * MM_SCR_STRIDE)
#define MM_OFF_BAR (MM_OFF_COMB + 2 * (MM_SUBS - 1) * 128 * 8)
#define MM_SMEM_BYTES (MM_OFF_BAR + 128)                // bars: ready[2] free[2] tfull[2] tempty[2] tile[2]

__constant__ int c_nlm_w_mma[MM_NW];

void fpb_upload_nlm_table_mma(const int* tab, cudaStream_t st) {
    cudaMemcpyToSymbolAsync(c_nlm_w_mma, tab, sizeof(int) * MM_NW, 0, cudaMemcpyHostToDevice, st);
}

__device__ __forceinline__ uint32_t mm_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle, version 1 (sm_100): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | 1 << 46
__device__ __forceinline__ uint64_t mm_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46);
}
// kind::i8: D = s32 (2 << 4), A = s8 (1 << 7), B = u8 (0 << 10), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
#define MM_IDESC(N) ((2u << 4) | (1u << 7) | (((uint32_t)(N) >> 3) << 17) | ((128u >> 4) << 24))

__device__ __forceinline__ void mm_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
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
// `backoff_ns`: sleep between polls - a spinning waiter takes issue slots from the warps that do the work
__device__ __forceinline__ void mm_bar_wait(uint32_t bar, uint32_t parity, unsigned backoff_ns = 0) {
    uint32_t done = 0;
    for (unsigned spin = 0; ; ++spin) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");       // suspend-time hint (ns)
        if (done) break;
        if (backoff_ns) __nanosleep(backoff_ns);
        if (spin > (1u << 22)) __trap();         // a lost arrival must fail loudly, never hang the GPU
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

__device__ __forceinline__ int mm_max3(int a, int b, int c) { return max(max(a, b), c); }     // VIMNMX3

template <bool PROF>
__global__ void __launch_bounds__(MM_THREADS, 1)
k_nlm_mma(const uint8_t* __restrict__ src, int W, int H, int n_img, uint8_t* __restrict__ dst,
          const __grid_constant__ CUtensorMap tmap, unsigned long long* __restrict__ prof) {
    // optional phase timing (FPB_NLM_PROF=1): clock64 deltas of lane 0 of one warp per role, summed over the CTA's blocks
    unsigned long long pt_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pc_ = 0;
#define MM_STAMP(k) do { if (PROF && lane == 0) { const unsigned long long c_ = clock64(); pt_[k] += c_ - pc_; pc_ = c_; } } while (0)
    extern __shared__ __align__(1024) uint8_t sm[];
    uint32_t* sHq = reinterpret_cast<uint32_t*>(sm + MM_OFF_HQ);     // 7-tap row sums of b^2 [42][28]
    uint32_t* sHb = reinterpret_cast<uint32_t*>(sm + MM_OFF_HB);     // 7-tap row sums of b   [42][28]
    int* sLut = reinterpret_cast<int*>(sm + MM_OFF_LUT);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + MM_OFF_BAR);
    uint32_t& tmem_base_sh = *reinterpret_cast<uint32_t*>(sm + MM_OFF_BAR + 96);   // after the ten mbarriers

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_ready = mm_smem(&bars[0]), bar_free = mm_smem(&bars[2]), bar_tfull = mm_smem(&bars[4]),
                   bar_tempty = mm_smem(&bars[6]), bar_tile = mm_smem(&bars[8]);
    const int bx_n = (W + MM_BW - 1) / MM_BW, by_n = (H + MM_BH - 1) / MM_BH;
    const unsigned total = (unsigned)n_img * (unsigned)bx_n * (unsigned)by_n;

    // ---- one-time setup
    if (warp == MM_EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(mm_smem(&tmem_base_sh)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        if (lane == 0) {
            for (int i = 0; i < 2; ++i) {
                mm_bar_init(bar_ready + 8 * i, 1); mm_bar_init(bar_free + 8 * i, MM_EPI_WARPS);
                mm_bar_init(bar_tfull + 8 * i, 1); mm_bar_init(bar_tempty + 8 * i, MM_EPI_WARPS);
                mm_bar_init(bar_tile + 8 * i, 1);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    } else if (warp < MM_EPI_WARPS) {
        for (int i = tid; i < MM_NW; i += MM_EPI_WARPS * 32) sLut[i] = c_nlm_w_mma[i];
    } else {
        // tails of the two A operands (k = 56..63 of every row): six digit slots x 127, the remainder slot x 1, zero
        const int pt = tid - (MM_EPI_WARPS + 1) * 32;
        for (int i = pt; i < 2 * 128; i += MM_PROD) {
            const int m = i & 127;
            *reinterpret_cast<uint2*>(sm + MM_OFF_A + (i >> 7) * MM_A_BYTES + (m >> 3) * 512 + 3 * 128 + (m & 7) * 16 + 8) =
                make_uint2(0x7F7F7F7Fu, 0x00017F7Fu);
        }
    }
    MM_FENCE_BEFORE();
    __syncthreads();
    MM_FENCE_AFTER();
    const uint32_t tmem = tmem_base_sh;

    if (warp == MM_EPI_WARPS) {
        // =========================================================== MMA issuer
        uint32_t it = 0, nb = 0;
        for (unsigned blk = blockIdx.x; blk < total; blk += gridDim.x, ++nb) {
            const uint32_t buf = nb & 1u;
            mm_bar_wait(bar_ready + 8 * buf, (nb >> 1) & 1u, 500);        // operands of this block are in shared memory
            MM_FENCE_AFTER();
            const uint64_t adesc = mm_desc(mm_smem(sm + MM_OFF_A + buf * MM_A_BYTES));
            const uint32_t b_addr = mm_smem(sm + MM_OFF_B + buf * MM_B_BYTES);
            for (int c = 0; c < MM_NCHUNK; ++c, ++it) {
                const uint32_t tb = it & 1u, use = it >> 1;
                mm_bar_wait(bar_tempty + 8 * tb, (use & 1u) ^ 1u, 200);   // the epilogue has drained this accumulator
                MM_FENCE_AFTER();
                if (lane == 0) {
                    const uint64_t bdesc = mm_desc(b_addr + (uint32_t)c * 16384u);
                    const uint32_t idesc = c < MM_NCHUNK - 1 ? MM_IDESC(256) : MM_IDESC(128);
                    mm_mma(tmem + tb * 256, adesc, bdesc, idesc, 0u);
                    mm_mma(tmem + tb * 256, adesc + 16, bdesc + 16, idesc, 1u);     // + 256 bytes: k 32..63
                    mm_commit(bar_tfull + 8 * tb);
                }
                __syncwarp();
            }
        }
    } else if (warp > MM_EPI_WARPS) {
        // =========================================================== producers: tile, B operand, digits, tables, A operand
        const int pt = tid - (MM_EPI_WARPS + 1) * 32, pw = pt >> 5;
        const uint32_t tile_addr = mm_smem(sm + MM_OFF_TILE);
        if (pt == 0) {
            for (unsigned k = 0; k < 2; ++k) {
                const unsigned long long blk = (unsigned long long)blockIdx.x + (unsigned long long)k * gridDim.x;
                if (blk < total) {
                    const MmBlock kb = mm_block((unsigned)blk, bx_n, by_n);
                    mm_tile_tma(tile_addr + k * MM_TILE_BYTES, &tmap, bar_tile + 8 * k, (kb.x0 & ~15) - 16, kb.y0 - 13, kb.b);
                }
            }
        }
        uint32_t nb = 0;
        for (unsigned blk = blockIdx.x; blk < total; blk += gridDim.x, ++nb) {
            const uint32_t buf = nb & 1u, use = nb >> 1;
            const MmBlock k = mm_block(blk, bx_n, by_n);
            const int x0 = k.x0, y0 = k.y0;
            const uint8_t* img = src + (size_t)k.b * W * H;
            uint8_t* tile = sm + MM_OFF_TILE + buf * MM_TILE_BYTES;
            uint8_t* sB = sm + MM_OFF_B + buf * MM_B_BYTES;
            uint8_t* sA = sm + MM_OFF_A + buf * MM_A_BYTES;
            uint32_t* sNQI = reinterpret_cast<uint32_t*>(sm + MM_OFF_NQI) + buf * MM_NQ_WORDS;
            int* sNAp = reinterpret_cast<int*>(sm + MM_OFF_NAP) + buf * 128;
            const int ox = (x0 & ~15) - 16, tx0 = 3 + (x0 & 8);            // image column of tile column 0; tile column of candidate column 0
            if (PROF && lane == 0) pc_ = clock64();
            mm_bar_wait(bar_free + 8 * buf, (use & 1u) ^ 1u, 1000);         // the epilogue is done with what block nb-2 left in this buffer
            MM_STAMP(0);
            mm_bar_wait(bar_tile + 8 * buf, use & 1u, 100);
            MM_STAMP(1);
            // ---- tile: rows y0-13 .. y0+28, columns ox .. ox+63; where the box left the image zeros arrived: BORDER_REFLECT_101
            if (ox < 0 || ox + MM_TS > W || y0 < 13 || y0 + 29 > H) {
                for (int i = pt; i < MM_TR * MM_TS; i += MM_PROD) {
                    const int r = i / MM_TS, c = i - r * MM_TS;
                    const int gy = y0 - 13 + r, gx = ox + c;
                    if ((unsigned)gy >= (unsigned)H || (unsigned)gx >= (unsigned)W)
                        tile[i] = img[(size_t)fpb_reflect101(gy, H) * W + fpb_reflect101(gx, W)];
                }
            }
            MM_BAR_SYNC(2, MM_PROD);
            MM_STAMP(2);
            // ---- im2col of the B operand.  Patch rows (2j, 2j+1) of a candidate are the 16 bytes of one K chunk: a lane walks
            //      down its band of tile rows with the previous row in registers and stores {row R-1, row R} for the three
            //      candidates that have them as patch rows (0,1), (2,3), (4,5), and row R alone as patch row 6.  Byte 7 of every
            //      8-byte piece is cleared here and receives the candidate's pad digit below.
            if (lane < MM_CC) {
                const int Rs = pw * 7, Re = Rs + 7;
                uint2 prev = make_uint2(0u, 0u);
                if (Rs > 0) { prev = mm_window(tile + (Rs - 1) * MM_TS, tx0 + lane); prev.y &= 0x00FFFFFFu; }
                uint8_t* bp0 = sB + (lane >> 3) * 512 + (lane & 7) * 16;
#pragma unroll
                for (int rr = 0; rr < 7; ++rr) {
                    const int R = Rs + rr;
                    uint2 cur = mm_window(tile + R * MM_TS, tx0 + lane);
                    cur.y &= 0x00FFFFFFu;
                    sHq[R * MM_CC + lane] = __dp4a(cur.x, cur.x, __dp4a(cur.y, cur.y, 0u));
                    sHb[R * MM_CC + lane] = __dp4a(cur.x, 0x01010101u, __dp4a(cur.y, 0x01010101u, 0u));
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
                (void)Re;
            }
            MM_BAR_SYNC(2, MM_PROD);
            MM_STAMP(3);
            // ---- per candidate: N(b) = sum b^2, S(b) = sum b sliding down six candidate rows per thread;
            //      D = 256 S - N, Q = ceil(D/2) = 127 (7 d_pad + sum tail) + r -> digit bytes; (D & 1) << 8 | I(q) for the survivors
            if (pt < 6 * MM_CC) {
                const int cxi = pt % MM_CC, c0 = (pt / MM_CC) * 6;
                uint32_t nbs = 0, sbs = 0;
#pragma unroll
                for (int pr = 0; pr < 7; ++pr) { nbs += sHq[(c0 + pr) * MM_CC + cxi]; sbs += sHb[(c0 + pr) * MM_CC + cxi]; }
#pragma unroll
                for (int t = 0; t < 6; ++t) {
                    const int cyi = c0 + t, n = cyi * MM_NPR + cxi;
                    const uint32_t dq = 256u * sbs - nbs, Q = (dq + 1u) >> 1;
                    const uint32_t T = Q / 127u, r = Q - 127u * T;
                    const uint32_t dpad = min(255u, T / 7u), t2 = T - 7u * dpad;
                    const uint32_t f = t2 / 255u, g2 = t2 - 255u * f;          // tail digits: 255 x f, then g2, then zeros (f <= 5)
                    const uint32_t lo = __funnelshift_lc(0xFFFFFFFFu, 0u, 8u * min(f, 4u)) | (f < 4u ? g2 << (8u * f) : 0u);
                    const uint32_t hi = (f == 5u ? 0xFFu | (g2 << 8) : (f == 4u ? g2 : 0u)) | (r << 16);
                    uint8_t* base = sB + (n >> 3) * 512 + (n & 7) * 16;
#pragma unroll
                    for (int pr = 0; pr < 7; ++pr) base[(pr >> 1) * 128 + (pr & 1) * 8 + 7] = (uint8_t)dpad;
                    *reinterpret_cast<uint2*>(base + 3 * 128 + 8) = make_uint2(lo, hi);
                    sNQI[n] = ((dq & 1u) << 8) | tile[(cyi + 3) * MM_TS + tx0 + cxi + 3];
                    const unsigned pr_ = (unsigned)(cyi - 10), pc_ = (unsigned)(cxi - 10);
                    if (pr_ < (unsigned)MM_BH && pc_ < (unsigned)MM_BW)
                        sNAp[(((pr_ >> 3) * 2 + (pc_ >> 2)) << 5) + ((pr_ & 7) << 2) + (pc_ & 3)] = (int)nbs;
                    if (t < 5) {
                        nbs += sHq[(cyi + 7) * MM_CC + cxi] - sHq[cyi * MM_CC + cxi];
                        sbs += sHb[(cyi + 7) * MM_CC + cxi] - sHb[cyi * MM_CC + cxi];
                    }
                }
            }
            MM_STAMP(4);
            // ---- A operand: the 128 pixels' own patches as a - 128 (s8), pad byte of every patch row = 127
            for (int i = pt; i < 128 * 7; i += MM_PROD) {
                const int m = i & 127, pr = i >> 7;
                const int r = ((m >> 6) << 3) + ((m & 31) >> 2), c = (((m >> 5) & 1) << 2) + (m & 3);
                uint2 v = mm_window(tile + (r + 10 + pr) * MM_TS, tx0 + c + 10);
                v.x ^= 0x80808080u;
                v.y = ((v.y ^ 0x80808080u) & 0x00FFFFFFu) | 0x7F000000u;
                *reinterpret_cast<uint2*>(sA + mm_op_off(m, pr)) = v;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // operand stores -> visible to the MMA (async proxy)
            MM_BAR_SYNC(2, MM_PROD);
            MM_STAMP(5);
            if (pt == 0) {
                const unsigned long long nxt = (unsigned long long)blk + 2ull * gridDim.x;      // this tile buffer's next block
                if (nxt < total) {
                    const MmBlock kb = mm_block((unsigned)nxt, bx_n, by_n);
                    mm_tile_tma(tile_addr + buf * MM_TILE_BYTES, &tmap, bar_tile + 8 * buf, (kb.x0 & ~15) - 16, kb.y0 - 13, kb.b);
                }
                mm_bar_arrive(bar_ready + 8 * buf);
            }
        }
    } else {
        // =========================================================== epilogue warps
        const int quad = warp & 3, sub = warp >> 2;
        const int rbase = (quad >> 1) * 8, cbase = (quad & 1) * 4;
        const int r_abs = rbase + (lane >> 2), c_rel = lane & 3;         // this lane's pixel inside the block / its tile
        const uint32_t colmask = ((1u << 21) - 1u) << c_rel;             // bit j: candidate column cbase + j is in this pixel's reach
        uint32_t* scr = reinterpret_cast<uint32_t*>(sm + MM_OFF_SCR + (warp * 32 + lane) * MM_SCR_STRIDE);
        const uint32_t tbase = tmem + ((uint32_t)(quad * 32) << 16) + cbase;
        uint32_t it = 0, nb = 0;
        for (unsigned blk = blockIdx.x; blk < total; blk += gridDim.x, ++nb) {
            const uint32_t buf = nb & 1u;
            const MmBlock k = mm_block(blk, bx_n, by_n);
            const uint32_t* sNQI = reinterpret_cast<const uint32_t*>(sm + MM_OFF_NQI) + buf * MM_NQ_WORDS;
            const int* sNAp = reinterpret_cast<const int*>(sm + MM_OFF_NAP) + buf * 128;
            uint32_t* sComb = reinterpret_cast<uint32_t*>(sm + MM_OFF_COMB) + buf * (MM_SUBS - 1) * 256;
            if (PROF && lane == 0) pc_ = clock64();
            mm_bar_wait(bar_ready + 8 * buf, (nb >> 1) & 1u, 100);
            MM_STAMP(0);
            const int na = sNAp[quad * 32 + lane];
            const int thr = (na - MM_SSD_MAX + 1) >> 1;                   // e >= thr  <=  SSD <= 33791
            unsigned sw = 0, swp = 0;
            for (int c = 0; c < MM_NCHUNK; ++c, ++it) {
                const uint32_t tb = it & 1u, use = it >> 1;
                mm_bar_wait(bar_tfull + 8 * tb, use & 1u, 60);
                MM_STAMP(1);
                MM_FENCE_AFTER();
                // this warp's rows of the chunk: cy = sub (mod MM_SUBS) inside [8c, 8c+7] and inside the tile's reach [rbase, rbase+27]
                const int lo_ = max(8 * c, rbase), hi_ = min(min(8 * c + 7, MM_CR - 1), rbase + 27);
                for (int cyi = lo_ + (sub + MM_SUBS - lo_ % MM_SUBS) % MM_SUBS; cyi <= hi_; cyi += MM_SUBS) {
                    const int rr = cyi - 8 * c;
                    int e[24];
                    const uint32_t taddr = tbase + tb * 256 + rr * 32;
                    MM_LD16(taddr, e, 0); MM_LD8(taddr + 16, e, 16);
                    const bool rel = (unsigned)(cyi - r_abs) <= 20u;              // this lane's pixel row reaches candidate row cy
                    const int thr_row = rel ? thr : 0x7FFFFFFF;
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    MM_STAMP(4);
                    int m0 = mm_max3(e[0], e[1], e[2]), m1 = mm_max3(e[3], e[4], e[5]), m2 = mm_max3(e[6], e[7], e[8]),
                        m3 = mm_max3(e[9], e[10], e[11]);
                    m0 = mm_max3(m0, e[12], e[13]); m1 = mm_max3(m1, e[14], e[15]); m2 = mm_max3(m2, e[16], e[17]);
                    m3 = mm_max3(m3, e[18], e[19]);
                    m0 = mm_max3(m0, e[20], e[21]); m1 = mm_max3(m1, e[22], e[23]);
                    m0 = mm_max3(m0, m1, max(m2, m3));
                    const bool any_hit = __any_sync(0xffffffffu, m0 >= thr_row);
                    MM_STAMP(5);
                    if (any_hit) {
                        // bit j = (e[j] >= thr_row): the sign of thr_row - 1 - e[j], collected by funnel shifts (three chains)
                        const int tm1 = thr - 1;
                        uint32_t s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
                        for (int j = 7; j >= 0; --j) {
                            s0 = __funnelshift_l((uint32_t)(tm1 - e[j]), s0, 1);
                            s1 = __funnelshift_l((uint32_t)(tm1 - e[8 + j]), s1, 1);
                            s2 = __funnelshift_l((uint32_t)(tm1 - e[16 + j]), s2, 1);
                        }
                        uint32_t pm = (s0 | (s1 << 8) | (s2 << 16)) & (rel ? colmask : 0u);
                        MM_STAMP(6);
                        if (pm) {
                            uint4* sp = reinterpret_cast<uint4*>(scr);
#pragma unroll
                            for (int q = 0; q < 6; ++q) sp[q] = make_uint4((uint32_t)e[4 * q], (uint32_t)e[4 * q + 1], (uint32_t)e[4 * q + 2], (uint32_t)e[4 * q + 3]);
                            const uint32_t* nq_row = sNQI + cyi * MM_NPR + cbase;
                            while (pm) {
                                const int j = __ffs((int)pm) - 1;
                                pm &= pm - 1u;
                                const int ev = (int)scr[j];
                                const uint32_t v = nq_row[j];
                                const int ssd = na + (int)(v >> 8) - 2 * ev;          // = N(p) - D(q) - 2 <a - 128, b> exactly
                                const unsigned w = (unsigned)sLut[ssd >> 6];
                                sw += w; swp += w * (v & 255u);
                            }
                        }
                        __syncwarp();
                        MM_STAMP(7);
                    }
                }
                MM_FENCE_BEFORE();
                __syncwarp();
                if (lane == 0) mm_bar_arrive(bar_tempty + 8 * tb);
                MM_STAMP(2);
            }
            // ---- the MM_SUBS warps of a quadrant hold partial sums of the same 32 pixels
            if (sub > 0) *reinterpret_cast<uint2*>(sComb + (sub - 1) * 256 + (quad * 32 + lane) * 2) = make_uint2(sw, swp);
            MM_BAR_SYNC(3, MM_EPI_WARPS * 32);
            if (sub == 0) {
#pragma unroll
                for (int q = 0; q < MM_SUBS - 1; ++q) {
                    const uint2 v = *reinterpret_cast<const uint2*>(sComb + q * 256 + (quad * 32 + lane) * 2);
                    sw += v.x; swp += v.y;
                }
                const int gy = k.y0 + r_abs, gx = k.x0 + cbase + c_rel;
                if (gy < H && gx < W) dst[(size_t)k.b * W * H + (size_t)gy * W + gx] = (uint8_t)min((swp + sw / 2u) / sw, 255u);
            }
            __syncwarp();
            if (lane == 0) mm_bar_arrive(bar_free + 8 * buf);             // tables / operands of this buffer may be rebuilt
            MM_STAMP(3);
        }
    }
    if (PROF && lane == 0 && (warp == 0 || warp == MM_EPI_WARPS + 1))
        for (int q = 0; q < 8; ++q) atomicAdd(&prof[(warp == 0 ? 0 : 8) + q], pt_[q]);
    MM_FENCE_BEFORE();
    __syncthreads();
    if (warp == MM_EPI_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
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

// false: shape / alignment outside what this kernel takes (the caller falls back to the integer-ALU kernel)
bool fpb_nlm_mma(FpbLaunch L, const uint8_t* src, int n, int W, int H, uint8_t* dst) {
    static int sms[64]; int dev = 0; cudaGetDevice(&dev);
    if (!sms[dev & 63]) cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    const long long total = (long long)n * ((W + MM_BW - 1) / MM_BW) * ((H + MM_BH - 1) / MM_BH);
    if (total >= (1ll << 31) || W < 64 || H < 48) return false;
    // the image batch as a 3-D u8 tensor (W, H, n); one TMA box = one 64 x 42 tile (out-of-image bytes arrive as zeros and are
    // patched to BORDER_REFLECT_101 by the producers).  TMA needs a 16-byte aligned base and row pitch.
    MmEncodeTiledFn enc = mm_encode_tiled();
    if (!enc || (W % 16) != 0 || (((uintptr_t)src) % 16) != 0) return false;
    CUtensorMap tmap; memset(&tmap, 0, sizeof(tmap));
    const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    const cuuint64_t gstr[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
    const cuuint32_t box[3] = {MM_TS, MM_TR, 1}, estr[3] = {1, 1, 1};
    if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    FPB_OPT_IN_SMEM(k_nlm_mma<false>, MM_SMEM_BYTES);
    FPB_OPT_IN_SMEM(k_nlm_mma<true>, MM_SMEM_BYTES);
    long long grid = sms[dev & 63];                        // persistent: one CTA per SM (all 512 TMEM columns, 217 KB of shared memory)
    if (grid > total) grid = total;
    static const bool want_prof = getenv("FPB_NLM_PROF") != nullptr;
    unsigned long long* d_prof = nullptr;
    if (want_prof) { cudaMalloc(&d_prof, 128); cudaMemsetAsync(d_prof, 0, 128, L.st); }
    if (want_prof) k_nlm_mma<true><<<(unsigned)grid, MM_THREADS, MM_SMEM_BYTES, L.st>>>(src, W, H, n, dst, tmap, d_prof);
    else k_nlm_mma<false><<<(unsigned)grid, MM_THREADS, MM_SMEM_BYTES, L.st>>>(src, W, H, n, dst, tmap, d_prof);
    if (want_prof) {                                       // diagnostics only: synchronous
        unsigned long long hp[16]; cudaStreamSynchronize(L.st);
        cudaMemcpy(hp, d_prof, 128, cudaMemcpyDeviceToHost); cudaFree(d_prof);
        const double nb = (double)total;
        fprintf(stderr, "[k_nlm_mma] cycles/block  epilogue warp 0 rows: tmem load %.0f  max+vote %.0f  mask %.0f  survivors %.0f\n",
                hp[4] / nb, hp[5] / nb, hp[6] / nb, hp[7] / nb);
        fprintf(stderr, "[k_nlm_mma] cycles/block  epilogue warp 0: wait ready %.0f  wait tfull %.0f  rows(rest) %.0f  combine %.0f | "
                        "producer warp 0: wait free %.0f  wait tile %.0f  patch %.0f  im2col %.0f  digits %.0f  A+fence %.0f\n",
                hp[0] / nb, hp[1] / nb, hp[2] / nb, hp[3] / nb, hp[8] / nb, hp[9] / nb, hp[10] / nb, hp[11] / nb, hp[12] / nb, hp[13] / nb);
    }
    LAUNCH_COUNT(L);
    return true;
}
