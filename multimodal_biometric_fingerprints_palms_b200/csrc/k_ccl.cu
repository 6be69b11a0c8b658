// Connected-component passes of K4 and K7 (scikit-image calls of the reference):
//   remove_small_objects(mask, min_size, connectivity=1)   fingerprint_preprocess.py:73,167   (4-connected foreground)
//   remove_small_holes(mask, area_threshold)                fingerprint_preprocess.py:74,168   (= the same on the background)
//   reconstruction(marker, opened, 'dilation')              fingerprint_preprocess.py:80       (8-connected components holding a marker)
//
// Label-equivalence union-find in global memory over the whole batch at once (one thread per
// pixel, atomicMin hooking, then path flattening), followed by a per-root size count / marker flag.
// Component identity never leaves the device.
#include <stdlib.h>
#include "fpb_kernels.h"
#include "ccl_bits.cuh"


__device__ __forceinline__ int uf_find(const int* L, int x) {
    int p = __ldcg(L + x);                  // L2 reads: other blocks are hooking roots concurrently
    while (p != x) { x = p; p = __ldcg(L + x); }
    return x;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    for (;;) {
        a = uf_find(L, a); b = uf_find(L, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }       // hook the larger root under the smaller
        const int old = atomicMin(&L[b], a);
        if (old == b) return;
        b = old;
    }
}

// pixel is "on" when (src != 0) == polarity.
// Run-based initialisation: a warp covers 32 consecutive pixels of a row; every "on" pixel starts out pointing at
// the first pixel of its horizontal run INSIDE that 32-pixel segment (ballot + clz, no atomics), so horizontal
// connectivity inside segments costs nothing.
__global__ void k_ccl_init(const uint8_t* __restrict__ src, int W, int H, const int4* __restrict__ roi,
                           int polarity, int* __restrict__ labels, int* __restrict__ aux) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    const bool inside = x < d.w && y < d.h;
    const int o = (b * H + y) * W + x;
    const bool on = inside && ((((src[o] != 0) ? 1 : 0) == polarity));
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (!inside) return;
    int lab = -1;
    if (on) {
        const unsigned lane = threadIdx.x;                       // blockDim.x == 32
        const unsigned zeros_below = ~m & ((1u << lane) - 1u);
        const int start = zeros_below ? 32 - __clz(zeros_below) : 0;
        lab = o - (int)lane + start;
    }
    labels[o] = lab;
    aux[o] = 0;
}

// Unions only where a new adjacency appears:
//   * a segment-run that continues the run of the previous 32-pixel segment;
//   * vertically, at the leftmost pixel of every overlap between a run and the run above it;
//   * (8-connectivity) diagonally, only when neither the pixel above nor the horizontal neighbour already links them.
__global__ void k_ccl_merge(int W, int H, const int4* __restrict__ roi, int conn8, int* __restrict__ labels) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x >= d.w || y >= d.h) return;
    const int o = (b * H + y) * W + x;
    if (labels[o] < 0) return;
    const bool left = x > 0 && labels[o - 1] >= 0;
    if (left && (threadIdx.x == 0)) uf_union(labels, o, o - 1);             // run crosses a segment boundary
    if (y > 0) {
        const bool up = labels[o - W] >= 0;
        const bool upleft = x > 0 && labels[o - W - 1] >= 0;
        if (up) {
            if (!(left && upleft)) uf_union(labels, o, o - W);
        } else if (conn8) {
            if (upleft && !left) uf_union(labels, o, o - W - 1);
            const bool right = x + 1 < d.w && labels[o + 1] >= 0;
            const bool upright = x + 1 < d.w && labels[o - W + 1] >= 0;
            if (upright && !right) uf_union(labels, o, o - W + 1);
        }
    }
}

// flatten + per-root statistic: mode 0 = size count, mode 1 = "holds a marker pixel" flag
__global__ void k_ccl_flatten(int W, int H, const int4* __restrict__ roi, int* __restrict__ labels,
                              int* __restrict__ aux, int mode, const uint8_t* __restrict__ marker) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x >= d.w || y >= d.h) return;
    const int o = (b * H + y) * W + x;
    if (labels[o] < 0) return;
    const int r = uf_find(labels, o);
    labels[o] = r;      // only ever shortens paths towards the root: safe while others still walk
    if (mode == 0) {
        // warp-aggregated count: lanes with the same root elect a leader that adds their population once
        const unsigned peers = __match_any_sync(__activemask(), r);
        if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&aux[r], __popc(peers));
    } else if (marker[o]) aux[r] = 1;
}

// remove_small: dst = src, with "on" components smaller than min_size turned off (polarity 1) / on (polarity 0)
__global__ void k_ccl_apply_small(const uint8_t* __restrict__ src, int W, int H, const int4* __restrict__ roi,
                                  int polarity, int min_size, const int* __restrict__ labels,
                                  const int* __restrict__ sizes, uint8_t* __restrict__ dst) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x >= d.w || y >= d.h) return;
    const int o = (b * H + y) * W + x;
    int on = src[o] != 0;
    const int l = labels[o];
    if (l >= 0 && sizes[uf_find(labels, l)] < min_size) on = !polarity;
    dst[o] = on ? 255 : 0;
}

__global__ void k_ccl_apply_flag(int W, int H, const int4* __restrict__ roi, const int* __restrict__ labels,
                                 const int* __restrict__ flags, uint8_t* __restrict__ dst) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x >= d.w || y >= d.h) return;
    const int o = (b * H + y) * W + x;
    const int l = labels[o];
    dst[o] = (l >= 0 && flags[uf_find(labels, l)]) ? 255 : 0;
}

static inline dim3 px_grid(int n, int W, int H) { return dim3((W + 31) / 32, (H + 7) / 8, n); }

void fpb_remove_small(FpbLaunch L, const uint8_t* src, int n, int W, int H, const int4* roi, int polarity,
                      int min_size, int* labels, int* sizes, uint8_t* dst) {
    const dim3 blk(32, 8), grid = px_grid(n, W, H);
    k_ccl_init<<<grid, blk, 0, L.st>>>(src, W, H, roi, polarity, labels, sizes);            LAUNCH_COUNT(L);
    k_ccl_merge<<<grid, blk, 0, L.st>>>(W, H, roi, 0, labels);                              LAUNCH_COUNT(L);
    k_ccl_flatten<<<grid, blk, 0, L.st>>>(W, H, roi, labels, sizes, 0, nullptr);            LAUNCH_COUNT(L);
    k_ccl_apply_small<<<grid, blk, 0, L.st>>>(src, W, H, roi, polarity, min_size, labels, sizes, dst); LAUNCH_COUNT(L);
}

void fpb_reconstruct(FpbLaunch L, const uint8_t* src, const uint8_t* marker, int n, int W, int H, const int4* roi,
                     int* labels, int* flags, uint8_t* dst) {
    const dim3 blk(32, 8), grid = px_grid(n, W, H);
    k_ccl_init<<<grid, blk, 0, L.st>>>(src, W, H, roi, 1, labels, flags);                   LAUNCH_COUNT(L);
    k_ccl_merge<<<grid, blk, 0, L.st>>>(W, H, roi, 1, labels);                              LAUNCH_COUNT(L);
    k_ccl_flatten<<<grid, blk, 0, L.st>>>(W, H, roi, labels, flags, 1, marker);             LAUNCH_COUNT(L);
    k_ccl_apply_flag<<<grid, blk, 0, L.st>>>(W, H, roi, labels, flags, dst);                LAUNCH_COUNT(L);
}

// ------------------------------------------------------------------------------------------------
// K4 tail, fused (fingerprint_preprocess.py:73-81): remove_small_objects(80) -> remove_small_holes(150) -> opening with
// the 3x3 cross -> marker = erode(opened) -> reconstruction -> {0,255}.  One CTA per image, everything between the
// u8 input plane and the u8 output plane happens on bit rows in shared memory (ccl_bits.cuh).
// ------------------------------------------------------------------------------------------------
#define BF_THREADS 512
__global__ void __launch_bounds__(BF_THREADS)
k_bin_finish(const uint8_t* __restrict__ bin0, int W, int H, const int4* __restrict__ roi, int min_obj, int max_hole,
             int* __restrict__ labels, int* __restrict__ sizes, uint8_t* __restrict__ dst, int full_nw, int sm_cap) {
    extern __shared__ __align__(16) uint32_t bf_sm[];
    __shared__ int s_warp[33];
    const int b = blockIdx.x, tid = threadIdx.x;
    const FpbDims d = fpb_dims(roi, b, W, H);
    const int w = d.w, h = d.h, wpr = (w + 31) >> 5, nw = wpr * h;
    uint32_t* A = bf_sm; uint32_t* B = A + nw; uint32_t* Cb = B + nw; uint32_t* M = Cb + nw; uint32_t* wb = M + nw;
    int* const gparent = labels + (size_t)b * W * H;
    int* const gattr = sizes + (size_t)b * W * H;
    int* uf = sm_cap ? reinterpret_cast<int*>(bf_sm + 5 * (size_t)full_nw) : nullptr;   // shared-memory union-find scratch
    int *parent = gparent, *attr = gattr;
    const uint8_t* src = bin0 + (size_t)b * W * H;
    cb_pack_u8(src, W, w, h, wpr, A);
    __syncthreads();
    // remove_small_objects(min_obj), 4-connected
    cb_label(A, wpr, w, h, false, nullptr, wb, parent, attr, s_warp, uf, sm_cap);
    for (int i = tid; i < nw; i += BF_THREADS) B[i] = cb_select_word(A, wb, parent, attr, i, i % wpr, min_obj, false);
    __syncthreads();
    // remove_small_holes(max_hole): small 4-connected background components become foreground
    for (int i = tid; i < nw; i += BF_THREADS) Cb[i] = ~B[i] & cb_valid_mask(i % wpr, w);
    __syncthreads();
    parent = gparent; attr = gattr;
    cb_label(Cb, wpr, w, h, false, nullptr, wb, parent, attr, s_warp, uf, sm_cap);
    for (int i = tid; i < nw; i += BF_THREADS) A[i] = B[i] | cb_select_word(Cb, wb, parent, attr, i, i % wpr, max_hole, true);
    __syncthreads();
    // opening with the cross, marker = erode(opened)
    for (int i = tid; i < nw; i += BF_THREADS) B[i] = cb_cross_word(A, wpr, w, h, i / wpr, i % wpr, true);
    __syncthreads();
    for (int i = tid; i < nw; i += BF_THREADS) Cb[i] = cb_cross_word(B, wpr, w, h, i / wpr, i % wpr, false);
    __syncthreads();
    for (int i = tid; i < nw; i += BF_THREADS) M[i] = cb_cross_word(Cb, wpr, w, h, i / wpr, i % wpr, true);
    __syncthreads();
    // reconstruction by dilation: 8-connected components of `opened` that hold a marker pixel
    parent = gparent; attr = gattr;
    cb_label(Cb, wpr, w, h, true, M, wb, parent, attr, s_warp, uf, sm_cap);
    for (int i = tid; i < nw; i += BF_THREADS) A[i] = cb_select_word(Cb, wb, parent, attr, i, i % wpr, 1, false);
    __syncthreads();
    uint8_t* out = dst + (size_t)b * W * H;
    for (int y = tid >> 5; y < h; y += BF_THREADS / 32)                  // warp = row, lane = column: no division per pixel
        for (int x = tid & 31; x < w; x += 32)
            out[(size_t)y * W + x] = ((A[y * wpr + (x >> 5)] >> (x & 31)) & 1u) ? 255 : 0;
}

// returns false when the image is too large for the shared-memory path (caller falls back to the per-pixel kernels)
bool fpb_bin_finish(FpbLaunch L, const uint8_t* bin0, int n, int W, int H, const int4* roi, int min_obj, int max_hole,
                    int* labels, int* sizes, uint8_t* dst) {
    const size_t nw = (size_t)((W + 31) / 32) * H;
    size_t smem = nw * 5 * sizeof(uint32_t);
    // bit rows beyond one CTA's shared memory (e.g. 1024 x 1024): bands of rows over a thread-block cluster (k_cluster.cu);
    // FPB_BIN_CLUSTER=<2|4|8> forces that kernel at any size (A/B and tests)
    static const int force_cl = getenv("FPB_BIN_CLUSTER") ? atoi(getenv("FPB_BIN_CLUSTER")) : 0;
    if (force_cl == 2 || force_cl == 4 || force_cl == 8)
        return fpb_bin_finish_cluster(L, bin0, n, W, H, roi, min_obj, max_hole, labels, sizes, dst, force_cl);
    if (smem > 160 * 1024) return fpb_bin_finish_cluster(L, bin0, n, W, H, roi, min_obj, max_hole, labels, sizes, dst, 8);
    // mid-size images (512 x 512: 3.1 ms on one CTA per image, the runs no longer fit its shared-memory union-find) go to
    // clusters of four when the batch leaves SMs idle anyway
    static const bool no_cl = getenv("FPB_NO_CLUSTER") != nullptr;
    if (!no_cl && (size_t)W * H >= 384 * 384 && fpb_bin_finish_cluster(L, bin0, n, W, H, roi, min_obj, max_hole, labels, sizes, dst, 4)) return true;
    int sm_cap = 0;                                   // union-find arrays in shared memory when two CTAs per SM still fit
    if (smem + 16 * 1024 <= 110 * 1024) {
        sm_cap = (int)((110 * 1024 - smem) / 8);
        if (sm_cap > 8192) sm_cap = 8192;
        smem += (size_t)sm_cap * 8;
    }
    FPB_OPT_IN_SMEM(k_bin_finish, 160 * 1024);
    k_bin_finish<<<n, BF_THREADS, smem, L.st>>>(bin0, W, H, roi, min_obj, max_hole, labels, sizes, dst, (int)nw, sm_cap);
    LAUNCH_COUNT(L);
    return true;
}
