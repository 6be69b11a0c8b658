// K7b + K8: skeletonize -> isolated-pixel clean-up -> crossing-number minutiae.
//   /root/reference/src/preprocessing/fingerprint_preprocess.py:170-177   mask & gate, skimage.skeletonize,
//       convolve(skel, ones(3,3)) 'reflect' > 1
//   /root/reference/src/features/extract_features.py:38-69                 crossing number, row-major emission
//
// BIT-EXACT stage.  One CTA (1024 threads) per image; the image lives bit-packed in shared memory
// (32 pixels per word: 10 KB for 320x240, 128 KB for 1024x1024) for the whole iteration:
//   * each thinning sub-iteration reads a snapshot and writes a copy (scikit-image's fully parallel
//     passes): every thread forms its new words in registers, barrier, store, barrier;
//   * the 256-entry deletion table (DATA - scikit-image's neighbour coding NW=1 N=2 NE=4 E=8 SE=16 S=32
//     SW=64 W=128, value 1/2/3) sits in shared memory;
//   * convergence is a __syncthreads_or over "some pixel was deleted";
//   * the minutiae are emitted in the reference's row-major order by a block-wide exclusive scan of
//     per-thread counts over CONTIGUOUS word ranges (ballot/popc inside the word, no atomics, because
//     atomics would scramble the order the reference's list - and its JSON - has).
#include "fpb_kernels.h"
#include "hd_scalar.h"
#include "ccl_bits.cuh"


#define THIN_THREADS 1024
#define THIN_MAX_WPT 32          // words per thread (max image: 32768 words = 1024 x 1024 pixels)

__global__ void k_gate(const uint8_t* __restrict__ cleaned, const float* __restrict__ rel_smooth, int W, int H,
                       const int4* __restrict__ roi, float thresh, uint8_t* __restrict__ gate) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x >= d.w || y >= d.h) return;
    const size_t o = (size_t)b * W * H + (size_t)y * W + x;
    gate[o] = (cleaned[o] != 0 && rel_smooth[o] > thresh) ? 255 : 0;
}

void fpb_gate(FpbLaunch L, const uint8_t* cleaned, const float* rel_smooth, int n, int W, int H, const int4* roi,
              float thresh, uint8_t* gate) {
    dim3 blk(32, 8), grid((W + 31) / 32, (H + 7) / 8, n);
    k_gate<<<grid, blk, 0, L.st>>>(cleaned, rel_smooth, W, H, roi, thresh, gate);
    LAUNCH_COUNT(L);
}

__global__ void k_thresh_u8(const uint8_t* __restrict__ src, int W, int H, const int4* __restrict__ roi, int thr,
                            uint8_t* __restrict__ dst) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x >= d.w || y >= d.h) return;
    const size_t o = (size_t)b * W * H + (size_t)y * W + x;
    dst[o] = src[o] > thr ? 255 : 0;
}

void fpb_thresh_u8(FpbLaunch L, const uint8_t* src, int n, int W, int H, const int4* roi, int thr, uint8_t* dst) {
    dim3 blk(32, 8), grid((W + 31) / 32, (H + 7) / 8, n);
    k_thresh_u8<<<grid, blk, 0, L.st>>>(src, W, H, roi, thr, dst);
    LAUNCH_COUNT(L);
}

// three-row window around word k of row y: 34-bit strips (bit 0 = pixel x-1 of the word's first pixel)
struct Strip3 { unsigned long long t, m, b; };

__device__ __forceinline__ unsigned long long strip(const uint32_t* row, int k, int wpr) {
    const uint32_t prev = k > 0 ? row[k - 1] : 0u, cur = row[k], next = k + 1 < wpr ? row[k + 1] : 0u;
    return ((unsigned long long)cur << 1) | (unsigned long long)(prev >> 31) | ((unsigned long long)(next & 1u) << 33);
}

__device__ __forceinline__ Strip3 load_strips(const uint32_t* bits, int wpr, int h, int y, int k) {
    Strip3 s;
    s.t = y > 0 ? strip(bits + (y - 1) * wpr, k, wpr) : 0ull;
    s.m = strip(bits + y * wpr, k, wpr);
    s.b = y + 1 < h ? strip(bits + (y + 1) * wpr, k, wpr) : 0ull;
    return s;
}

// neighbour code of pixel j of the word: NW=1 N=2 NE=4 E=8 SE=16 S=32 SW=64 W=128
__device__ __forceinline__ unsigned nb_code(const Strip3& s, int j) {
    const unsigned t3 = (unsigned)(s.t >> j) & 7u, m3 = (unsigned)(s.m >> j) & 7u, b3 = (unsigned)(s.b >> j) & 7u;
    return t3 | ((m3 >> 2) << 3) | ((b3 >> 2) << 4) | (((b3 >> 1) & 1u) << 5) | ((b3 & 1u) << 6) | ((m3 & 1u) << 7);
}

// ending / bifurcation masks of word i (interior pixels only, extract_features.py:50)
__device__ __forceinline__ void cn_masks(const uint32_t* bits, int wpr, int w, int h, int i, uint32_t* endm, uint32_t* bifm) {
    uint32_t e = 0, f = 0;
    const uint32_t cur = bits[i];
    if (cur) {
        const int y = i / wpr, k = i - y * wpr;
        if (y >= 1 && y <= h - 2) {
            const Strip3 s = load_strips(bits, wpr, h, y, k);
            uint32_t rem = cur;
            while (rem) {
                const int j = __ffs(rem) - 1; rem &= rem - 1;
                const int x = k * 32 + j;
                if (x < 1 || x > w - 2) continue;
                const unsigned c = nb_code(s, j);
                const unsigned rot = ((c << 1) | (c >> 7)) & 255u;
                const int cn = __popc(c ^ rot) >> 1;       // = sum |P_i - P_i+1| / 2 around the ring
                if (cn == 1) e |= 1u << j;
                else if (cn == 3) f |= 1u << j;
            }
        }
    }
    *endm = e; *bifm = f;
}

template <int THIN_WPT>
__global__ void __launch_bounds__(THIN_THREADS, (THIN_WPT <= 8 ? 2 : 1))
k_thin_extract(const uint8_t* __restrict__ gate, int W, int H, const int4* __restrict__ roi,
               const uint8_t* __restrict__ table, uint8_t* __restrict__ skeleton, int* __restrict__ raw_count,
               uint32_t* __restrict__ raw, int do_thin, uint32_t* gscratch, FpbThinPre pre, int raw_cap, int pack_thr) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint8_t lut[256];
    __shared__ int scan[THIN_THREADS];
    const int b = blockIdx.x, tid = threadIdx.x;
    const FpbDims d = fpb_dims(roi, b, W, H);
    const int w = d.w, h = d.h;
    const int wpr = (w + 31) >> 5, nw = wpr * h;
    uint32_t* bits = gscratch ? gscratch + (size_t)b * (((W + 31) >> 5) * H) : smem;
    if (tid < 256) lut[tid] = table[tid];
    const bool zs_fast = table[256] != 0;          // the installed table is the built-in Zhang-Suen one: closed form, 32 pixels at a time
    const uint8_t* g = (pre.smooth ? pre.smooth : gate) + (size_t)b * W * H;
    cb_pack_u8(g, W, w, h, wpr, bits, pack_thr);
    __syncthreads();
    if (pre.smooth) {
        // ---- fused K7a (fingerprint_preprocess.py:166-170): remove_small_objects(64), remove_small_holes(80) on bit rows
        //      in shared memory (ccl_bits.cuh), then AND with gaussian_filter(reliability, 2.0) > rel_thresh
        uint32_t* Bq = smem + nw; uint32_t* Cq = Bq + nw; uint32_t* wb = Cq + nw;
        int* const gparent = pre.labels + (size_t)b * W * H;
        int* const gattr = pre.sizes + (size_t)b * W * H;
        int* uf = reinterpret_cast<int*>(wb + nw);                   // shared-memory union-find scratch, 2*sm_cap ints
        int *parent = gparent, *attr = gattr;
        cb_label(bits, wpr, w, h, false, nullptr, wb, parent, attr, scan, pre.sm_cap ? uf : nullptr, pre.sm_cap);
        for (int i = tid; i < nw; i += THIN_THREADS) Bq[i] = cb_select_word(bits, wb, parent, attr, i, i % wpr, pre.min_obj, false);
        __syncthreads();
        for (int i = tid; i < nw; i += THIN_THREADS) Cq[i] = ~Bq[i] & cb_valid_mask(i % wpr, w);
        __syncthreads();
        parent = gparent; attr = gattr;
        cb_label(Cq, wpr, w, h, false, nullptr, wb, parent, attr, scan, pre.sm_cap ? uf : nullptr, pre.sm_cap);
        const float* rs = pre.rel_smooth + (size_t)b * W * H;
        uint8_t* go = pre.gate_out ? pre.gate_out + (size_t)b * W * H : nullptr;
        for (int i = tid; i < nw; i += THIN_THREADS)
            Bq[i] |= cb_select_word(Cq, wb, parent, attr, i, i % wpr, pre.max_hole, true);
        __syncthreads();
        {   // gate: one warp per word, lanes = pixels (coalesced float reads, one ballot)
            const int lane = tid & 31, wid = tid >> 5;
            for (int y = wid; y < h; y += THIN_THREADS / 32)            // warp = row: no division per word
                for (int k = 0; k < wpr; ++k) {
                    const int i = y * wpr + k, x = k * 32 + lane;
                    const bool ok = x < w && rs[(size_t)y * W + x] > pre.thresh;
                    const uint32_t word = Bq[i] & __ballot_sync(0xffffffffu, ok);
                    if (lane == 0) bits[i] = word;
                    if (go && x < w) go[(size_t)y * W + x] = ((word >> lane) & 1u) ? 255 : 0;
                }
        }
        __syncthreads();
    }

    uint32_t nwords[THIN_WPT];
    // ---- thinning: sub-iteration 1 deletes table values {1,3}, sub-iteration 2 {2,3}; repeat until
    //      a full double pass deletes nothing.  Pixels whose eight neighbours are all set have code 255: they are found
    //      with eight shifted ANDs per word and decided by table[255] at once, so the per-pixel look-ups only run over
    //      the border pixels of the ridges.  In the fused form the sub-iterations ping-pong between `bits` and the
    //      (now free) K7a buffer - one barrier per sub-iteration; without a second buffer the new words wait in
    //      registers across a barrier.
    if (do_thin) {
        uint32_t* alt = pre.smooth ? smem + nw : nullptr;
        const unsigned v255 = lut[255];
        for (;;) {
            int any = 0;
            for (int pass = 1; pass <= 2; ++pass) {
                const bool del255 = (v255 == 3u) || (v255 == (unsigned)pass);
#pragma unroll
                for (int q = 0; q < THIN_WPT; ++q) {
                    const int i = tid + q * THIN_THREADS;
                    uint32_t out = 0;
                    if (i < nw) {
                        const uint32_t cur = bits[i];
                        out = cur;
                        if (cur) {
                            const int y = i / wpr, k = i - y * wpr;
                            const Strip3 s = load_strips(bits, wpr, h, y, k);
                            const uint32_t inner = cur & (uint32_t)(s.t & (s.t >> 1) & (s.t >> 2) & s.m & (s.m >> 2) &
                                                                    s.b & (s.b >> 1) & (s.b >> 2));
                            if (zs_fast) {
                                out = cur & ~fpb_zs_delete_mask((uint32_t)s.t, (uint32_t)(s.t >> 1), (uint32_t)(s.t >> 2), (uint32_t)(s.m >> 2),
                                                                (uint32_t)(s.b >> 2), (uint32_t)(s.b >> 1), (uint32_t)s.b, (uint32_t)s.m, pass);
                            } else {
                                if (del255) out &= ~inner;
                                uint32_t rem = cur & ~inner;
                                while (rem) {
                                    const int j = __ffs(rem) - 1; rem &= rem - 1;
                                    const unsigned v = lut[nb_code(s, j)];
                                    if (v == 3u || v == (unsigned)pass) out &= ~(1u << j);
                                }
                            }
                            any |= (out != cur);
                        }
                        if (alt) alt[i] = out;
                    }
                    nwords[q] = out;
                }
                __syncthreads();
                if (alt) { uint32_t* t = bits; bits = alt; alt = t; }
                else {
#pragma unroll
                    for (int q = 0; q < THIN_WPT; ++q) {
                        const int i = tid + q * THIN_THREADS;
                        if (i < nw) bits[i] = nwords[q];
                    }
                    __syncthreads();
                }
            }
            if (!__syncthreads_or(any)) break;
        }
       
        // ---- clean-up (:174-176): keep a pixel iff its 3x3 sum with 'reflect' border exceeds 1, i.e. it has a
        //      set 8-neighbour or lies on the image border (the reflected centre then counts twice)
#pragma unroll
        for (int q = 0; q < THIN_WPT; ++q) {
            const int i = tid + q * THIN_THREADS;
            uint32_t out = 0;
            if (i < nw) {
                const uint32_t cur = bits[i];
                out = cur;
                if (cur) {
                    const int y = i / wpr, k = i - y * wpr;
                    const Strip3 s = load_strips(bits, wpr, h, y, k);
                    uint32_t rem = cur;
                    while (rem) {
                        const int j = __ffs(rem) - 1; rem &= rem - 1;
                        const int x = k * 32 + j;
                        const bool border = (x == 0) || (y == 0) || (x == w - 1) || (y == h - 1);
                        if (!border && nb_code(s, j) == 0u) out &= ~(1u << j);
                    }
                }
            }
            nwords[q] = out;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < THIN_WPT; ++q) {
            const int i = tid + q * THIN_THREADS;
            if (i < nw) bits[i] = nwords[q];
        }
        __syncthreads();
    }
    // ---- skeleton plane out
    if (skeleton) {
        uint8_t* sk = skeleton + (size_t)b * W * H;
        for (int y = tid >> 5; y < h; y += THIN_THREADS / 32)            // warp = row, lane = column: no division per pixel
            for (int x = tid & 31; x < w; x += 32)
                sk[(size_t)y * W + x] = ((bits[y * wpr + (x >> 5)] >> (x & 31)) & 1u) ? 255 : 0;
    }
   
    if (!raw_count) return;
    // ---- K8: crossing numbers.  Thread t owns the contiguous words [t*cpt, (t+1)*cpt): count, block
    //      exclusive scan, then recompute the masks and write in order.
    const int cpt = (nw + THIN_THREADS - 1) / THIN_THREADS;      // <= THIN_WPT
    int cnt = 0;
    for (int q = 0; q < cpt; ++q) {
        const int i = tid * cpt + q;
        if (i < nw) { uint32_t e, f; cn_masks(bits, wpr, w, h, i, &e, &f); cnt += __popc(e | f); }
    }
    scan[tid] = cnt;
    __syncthreads();
    for (int off = 1; off < THIN_THREADS; off <<= 1) {
        const int add = tid >= off ? scan[tid - off] : 0;
        __syncthreads();
        scan[tid] += add;
        __syncthreads();
    }
    int pos = scan[tid] - cnt;
    if (tid == THIN_THREADS - 1) raw_count[b] = scan[tid];          // the TRUE count: the host fails loudly when it exceeds raw_cap
    if (cnt == 0) return;
    uint32_t* out = raw + (size_t)b * raw_cap;
    for (int q = 0; q < cpt; ++q) {
        const int i = tid * cpt + q;
        if (i >= nw) break;
        uint32_t e, f; cn_masks(bits, wpr, w, h, i, &e, &f);
        uint32_t m = e | f;
        const int y = i / wpr, k = i - y * wpr;
        while (m) {
            const int j = __ffs(m) - 1; m &= m - 1;
            if (pos < raw_cap) out[pos] = fpb_pack_raw(k * 32 + j, y, (f >> j) & 1u);
            ++pos;
        }
    }
}

template <int WPT>
static void launch_thin(FpbLaunch L, size_t smem, bool big, const uint8_t* gate, int n, int W, int H, const int4* roi,
                        const uint8_t* table, uint8_t* skeleton, int* raw_count, uint32_t* raw, int do_thin, uint32_t* bitscratch,
                        FpbThinPre pre, int raw_cap, int pack_thr) {
    FPB_OPT_IN_SMEM(k_thin_extract<WPT>, 200 * 1024);
    k_thin_extract<WPT><<<n, THIN_THREADS, big ? 0 : smem, L.st>>>(gate, W, H, roi, table, skeleton, raw_count, raw, do_thin,
                                                                 big ? bitscratch : nullptr, pre, raw_cap, pack_thr);
}

static void thin_dispatch(FpbLaunch L, const uint8_t* gate, int n, int W, int H, const int4* roi, const uint8_t* table,
                          uint8_t* skeleton, int* raw_count, uint32_t* raw, int do_thin, uint32_t* bitscratch, FpbThinPre pre,
                          int raw_cap, int pack_thr) {
    const int nw = ((W + 31) / 32) * H;
    size_t smem = (size_t)nw * 4 * (pre.smooth ? 4 : 1);
    pre.sm_cap = 0;
    if (pre.smooth && smem + 8 * 1024 <= 100 * 1024) {              // room for the union-find arrays at two CTAs per SM
        pre.sm_cap = (int)((100 * 1024 - smem) / 8);
        if (pre.sm_cap > 8192) pre.sm_cap = 8192;
        smem += (size_t)pre.sm_cap * 8;
    }
    const bool big = smem > 200 * 1024;
#define ARGS L, smem, big, gate, n, W, H, roi, table, skeleton, raw_count, raw, do_thin, bitscratch, pre, raw_cap, pack_thr
    if (nw <= 4 * THIN_THREADS) launch_thin<4>(ARGS);
    else if (nw <= 8 * THIN_THREADS) launch_thin<8>(ARGS);
    else if (nw <= 16 * THIN_THREADS) launch_thin<16>(ARGS);
    else launch_thin<THIN_MAX_WPT>(ARGS);
#undef ARGS
    LAUNCH_COUNT(L);
}

void fpb_thin_extract(FpbLaunch L, const uint8_t* gate, int n, int W, int H, const int4* roi, const uint8_t* table,
                      uint8_t* skeleton, int* raw_count, uint32_t* raw, int raw_cap, int do_thin, uint32_t* bitscratch, int pack_thr) {
    FpbThinPre none; none.smooth = nullptr; none.rel_smooth = nullptr; none.gate_out = nullptr; none.labels = none.sizes = nullptr;
    none.thresh = 0.f; none.min_obj = none.max_hole = 0;
    thin_dispatch(L, gate, n, W, H, roi, table, skeleton, raw_count, raw, do_thin, bitscratch, none, raw_cap, pack_thr);
}

bool fpb_thin_fused(FpbLaunch L, FpbThinPre pre, int n, int W, int H, const int4* roi, const uint8_t* table,
                    uint8_t* skeleton, int* raw_count, uint32_t* raw, int raw_cap) {
    const size_t nw = (size_t)((W + 31) / 32) * H;
    if (nw * 16 > 160 * 1024) return false;              // four bit/word buffers must fit in shared memory
    thin_dispatch(L, nullptr, n, W, H, roi, table, skeleton, raw_count, raw, 1, nullptr, pre, raw_cap, 0);
    return true;
}

