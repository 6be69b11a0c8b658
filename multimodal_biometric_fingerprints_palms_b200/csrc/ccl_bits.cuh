// Connected components of a bit-packed image held in SHARED memory, one CTA per image.
//
// Replaces the per-pixel union-find over int32 label planes in HBM for the five component passes of K4 / K7
// (skimage remove_small_objects / remove_small_holes / reconstruction; fingerprint_preprocess.py:73-80,167-168):
//   * the unit is the horizontal RUN (maximal stretch of set pixels in a row), found with bit tricks on 32-pixel
//     words; a 320x240 ridge image has ~8 k runs instead of 76.8 k pixels;
//   * run ids come from a block-wide exclusive scan of per-word run-start counts (`wordbase`, shared memory);
//   * runs of adjacent rows that overlap (4-conn) or touch diagonally (8-conn) are united with atomicMin hooking on a
//     parent array (global scratch, L2 resident: a few accesses per run);
//   * a per-component attribute (pixel count, or "contains a marker pixel") is accumulated on the roots and each
//     run is kept / dropped as a whole.
// All functions are block-collective: every thread of the CTA must call them with the same arguments.
#pragma once
#include <stdint.h>
#include "fpb_common.cuh"

__device__ __forceinline__ uint32_t cb_valid_mask(int k, int w) {
    const int rem = w - k * 32;
    return rem >= 32 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
}
__device__ __forceinline__ uint32_t cb_le_mask(int j) { return j >= 31 ? 0xFFFFFFFFu : ((2u << j) - 1u); }   // bits 0..j
__device__ __forceinline__ uint32_t cb_ge_mask(int j) { return 0xFFFFFFFFu << j; }                           // bits j..31

// Pack a u8 plane (value > thr = set; thr 0 = non-zero) into bit rows: one warp per 32-pixel word, lanes = pixels (coalesced 32-byte
// reads, one ballot per word).  Block-collective; blockDim.x must be a multiple of 32.
__device__ __forceinline__ void cb_pack_u8(const uint8_t* __restrict__ src, int W, int w, int h, int wpr, uint32_t* bits, int thr = 0) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    // warp = row (no division per word), two words of the row per trip
    for (int y = wid; y < h; y += nwarp) {
        const uint8_t* row = src + (size_t)y * W;
        uint32_t* out = bits + y * wpr;
        for (int k = 0; k < wpr; k += 2) {
            const int xa = k * 32 + lane, xb = xa + 32;
            const bool a = xa < w && (int)row[xa] > thr;
            const bool c = (k + 1 < wpr) && xb < w && (int)row[xb] > thr;
            const uint32_t wa = __ballot_sync(0xffffffffu, a), wc = __ballot_sync(0xffffffffu, c);
            if (lane == 0) { out[k] = wa; if (k + 1 < wpr) out[k + 1] = wc; }
        }
    }
}

__device__ __forceinline__ uint32_t cb_starts(const uint32_t* bits, int i, int k) {
    const uint32_t m = bits[i];
    const uint32_t carry = k > 0 ? (bits[i - 1] >> 31) : 0u;
    return m & ~((m << 1) | carry);
}

// The union-find arrays live in global memory (any number of runs) or, when the caller offers shared-memory scratch and
// the runs fit, in shared memory: a hop then costs a shared-memory access instead of an L2 round trip.  Loads that race
// with other threads' atomics bypass L1 (global) / are volatile (shared).
__device__ __forceinline__ int cb_ld(const int* p) {
    return __isShared(p) ? *reinterpret_cast<const volatile int*>(p) : __ldcg(p);
}
__device__ __forceinline__ int cb_find(const int* L, int x) {
    int p = cb_ld(L + x);
    while (p != x) { x = p; p = cb_ld(L + x); }
    return x;
}
__device__ __forceinline__ void cb_st(int* p, int v) {
    if (__isShared(p)) *reinterpret_cast<volatile int*>(p) = v; else __stcg(p, v);
}
// find with path splitting: every node passed on the way is re-pointed at its grandparent.  Parents only ever decrease
// towards the root and stay inside the component, so the plain store is safe next to the other threads' atomicMin hooks
// (a hook it overwrites is re-established by the hooking thread's own retry, which continues from the value it displaced).
// Without it a vertical structure of h rows builds a chain of h hops and the finds cost O(h^2) L2 round trips per image.
// Loads of the hooking phase may come from L1 (ld.ca): a stale parent is still an ancestor (pointers only move up the
// tree), so the walk merely takes a hop more, and the atomicMin that decides a hook always sees L2.  The flatten / select
// passes after the barrier keep the L1-bypassing cb_ld.
__device__ __forceinline__ int cb_ld_hook(const int* p) {
    return __isShared(p) ? *reinterpret_cast<const volatile int*>(p) : __ldca(p);
}
__device__ __forceinline__ int cb_find_split(int* L, int x) {
    int p = cb_ld_hook(L + x);
    while (p != x) {
        const int gp = cb_ld_hook(L + p);
        if (gp == p) return p;
        cb_st(L + x, gp);
        x = p; p = gp;
    }
    return x;
}
__device__ __forceinline__ void cb_union(int* L, int a, int b) {
    for (;;) {
        a = cb_find_split(L, a); b = cb_find_split(L, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&L[b], a);
        if (old == b) return;
        b = old;
    }
}

// block-wide exclusive scan of one int per thread (blockDim.x <= 1024); s_warp: 33 ints of shared memory
__device__ __forceinline__ int cb_block_scan_excl(int v, int* s_warp, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += t; }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < nwarp ? s_warp[lane] : 0, winc = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, winc, off); if (lane >= off) winc += t; }
        s_warp[lane] = winc - w;
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    const int res = inc - v + s_warp[wid];
    *total = s_warp[32];
    __syncthreads();
    return res;
}

// id of the run of row-word `i` (column word k) that contains set pixel j
__device__ __forceinline__ int cb_run_id(const uint32_t* bits, const uint32_t* wordbase, int i, int k, int j) {
    return (int)wordbase[i] + __popc(cb_starts(bits, i, k) & cb_le_mask(j)) - 1;
}

// Label the runs of `bits` (wpr words per row, h rows; no set bits beyond column w-1).
//   parent[r] = root run id afterwards; attr[r] (r = root) = component pixel count (marker == nullptr) or
//   1/0 "holds a marker pixel" (marker != nullptr, same layout as bits).
// Returns the number of runs (uniform over the block).
// `sm_scratch` (optional): 2*sm_cap ints of shared memory; when the image has at most sm_cap runs, parent/attr are placed
// there and *parent_io / *attr_io are redirected to them (callers pass the same pointers on to cb_select_word).
static __device__ int cb_label(const uint32_t* bits, int wpr, int w, int h, bool conn8, const uint32_t* marker,
                        uint32_t* wordbase, int*& parent, int*& attr, int* s_warp, int* sm_scratch = nullptr, int sm_cap = 0) {
    const int nw = wpr * h, T = blockDim.x, tid = threadIdx.x;
    // ---- run ids: exclusive scan of run-start counts over the words in raster order (contiguous chunk per thread)
    const int cpt = (nw + T - 1) / T;
    int local = 0;
    for (int q = 0; q < cpt; ++q) { const int i = tid * cpt + q; if (i < nw) local += __popc(cb_starts(bits, i, i % wpr)); }
    int nruns = 0;
    int base = cb_block_scan_excl(local, s_warp, &nruns);
    for (int q = 0; q < cpt; ++q) {
        const int i = tid * cpt + q;
        if (i < nw) { wordbase[i] = (uint32_t)base; base += __popc(cb_starts(bits, i, i % wpr)); }
    }
    if (sm_scratch && nruns <= sm_cap) { parent = sm_scratch; attr = sm_scratch + sm_cap; }
    for (int r = tid; r < nruns; r += T) parent[r] = r;
    __syncthreads();
    __threadfence_block();
    // ---- one pass over the run starts: extent, attribute seed, unions with the row above
    const int c = conn8 ? 1 : 0;
    // Every thread takes the same NUMBER OF RUNS (a contiguous range of run ids), not the same number of words: with words dealt
    // round-robin a thread that met busy words kept the whole CTA at the barrier below (24 % of k_bin_finish's warp time).
    // The word of the first run is the last one whose word base does not exceed the run id (binary search), the rest follows
    // in raster order.
    const int per_thread = (nruns + T - 1) / T;
    {
        int r = tid * per_thread;
        const int r_end = min(nruns, r + per_thread);
        int i = 0;
        uint32_t st = 0u;
        if (r < r_end) {
            int lo_i = 0, hi_i = nw - 1;
            while (lo_i < hi_i) { const int mid = (lo_i + hi_i + 1) >> 1; if ((int)wordbase[mid] <= r) lo_i = mid; else hi_i = mid - 1; }
            i = lo_i;
            st = cb_starts(bits, i, i % wpr);
            for (int skip = r - (int)wordbase[i]; skip > 0; --skip) st &= st - 1;
        }
        for (; r < r_end; ++r) {
            while (!st) { ++i; st = cb_starts(bits, i, i % wpr); }      // the next word that starts a run (one exists: r < nruns)
            const int y = i / wpr, k = i - y * wpr;
            const uint32_t* row = bits + y * wpr;
            const int j = __ffs(st) - 1; st &= st - 1;
            // extent [s, e]
            const int s = k * 32 + j;
            int e, kk = k;
            int hit = 0;
            {
                uint32_t m = row[kk] >> j;
                const uint32_t inv = ~m;                       // zeros shifted in at the top read as "not set"
                int len = __ffs(inv) - 1;                      // ones from bit j upwards (inv != 0 because of the shift, or j == 0)
                if (j == 0 && inv == 0u) len = 32;
                if (marker) { const uint32_t rm = (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << j; hit |= (marker[y * wpr + kk] & rm) != 0u; }
                e = s + len - 1;
                while (((e & 31) == 31) && kk + 1 < wpr && (row[kk + 1] & 1u)) {      // continues into the next word
                    ++kk;
                    const uint32_t m2 = row[kk];
                    const int l2 = (m2 == 0xFFFFFFFFu) ? 32 : (__ffs(~m2) - 1);
                    if (marker) { const uint32_t rm = l2 >= 32 ? 0xFFFFFFFFu : ((1u << l2) - 1u); hit |= (marker[y * wpr + kk] & rm) != 0u; }
                    e += l2;
                    if (l2 < 32) break;
                }
            }
            attr[r] = marker ? hit : (e - s + 1);
            if (y == 0) continue;
            // runs of the row above that meet [s-c, e+c]: consecutive ids from the run holding the lowest set bit
            // in that range to the run holding the highest one
            const int lo = max(s - c, 0), hi = min(e + c, w - 1);
            const uint32_t* prow = bits + (y - 1) * wpr;
            const int k1 = lo >> 5, k2 = hi >> 5;
            int first = -1, last = -1;
            for (int q = k1; q <= k2; ++q) {
                uint32_t mm = prow[q];
                if (q == k1) mm &= cb_ge_mask(lo & 31);
                if (q == k2) mm &= cb_le_mask(hi & 31);
                if (!mm) continue;
                if (first < 0) first = cb_run_id(bits, wordbase, (y - 1) * wpr + q, q, __ffs(mm) - 1);
                last = cb_run_id(bits, wordbase, (y - 1) * wpr + q, q, 31 - __clz(mm));
            }
            for (int q = first; q >= 0 && q <= last; ++q) cb_union(parent, r, q);
        }
    }
    __syncthreads();
    __threadfence_block();
    // ---- flatten, then push every non-root run's attribute onto its root
    for (int r = tid; r < nruns; r += T) parent[r] = cb_find(parent, r);
    __syncthreads();
    for (int r = tid; r < nruns; r += T) {
        const int root = cb_ld(parent + r);
        if (root != r) {
            const int a = cb_ld(attr + r);
            if (marker) { if (a) atomicOr(&attr[root], 1); } else atomicAdd(&attr[root], a);
        }
    }
    __syncthreads();
    return nruns;
}

// out word i = the pixels of word i whose component satisfies: (size mode) count >= min_size ; (marker mode, min_size = 1)
// flag != 0.  `keep_small` inverts the choice (returns the pixels of the components that FAIL the test).
__device__ __forceinline__ uint32_t cb_select_word(const uint32_t* bits, const uint32_t* wordbase, const int* parent,
                                                   const int* attr, int i, int k, int min_size, bool keep_small) {
    uint32_t m = bits[i], out = 0;
    if (!m) return 0u;
    const uint32_t st = cb_starts(bits, i, k);
    int id = (int)wordbase[i] - ((m & 1u) && !(st & 1u) ? 1 : 0);     // first run of the word: continues from the left?
    while (m) {
        const int j = __ffs(m) - 1;
        const uint32_t low = 1u << j;
        const uint32_t rm = (m ^ (m + low)) & m;                      // the carry ripples through exactly the lowest run
        const int a = cb_ld(attr + cb_ld(parent + id));
        const bool pass = a >= min_size;
        if (pass != keep_small) out |= rm;
        m &= ~rm;
        ++id;
    }
    return out;
}

// 3x3 cross (cv2.getStructuringElement(MORPH_ELLIPSE,(3,3))) on bit rows: erosion treats out-of-image neighbours as set,
// dilation as clear (cv2 border semantics of morphologyEx / erode).
__device__ __forceinline__ uint32_t cb_cross_word(const uint32_t* bits, int wpr, int w, int h, int y, int k, bool erode) {
    const uint32_t fill = erode ? 0xFFFFFFFFu : 0u;
    const uint32_t* row = bits + y * wpr;
    auto ext = [&](const uint32_t* r, int kk) -> uint32_t {           // word kk of row r with out-of-image bits = fill
        if (kk < 0 || kk >= wpr) return fill;
        const uint32_t vm = cb_valid_mask(kk, w);
        return (r[kk] & vm) | (fill & ~vm);
    };
    const uint32_t cur = ext(row, k), prev = ext(row, k - 1), next = ext(row, k + 1);
    const uint32_t left = (cur << 1) | (prev >> 31), right = (cur >> 1) | (next << 31);
    const uint32_t up = y > 0 ? ext(row - wpr, k) : fill, dn = y + 1 < h ? ext(row + wpr, k) : fill;
    const uint32_t v = erode ? (cur & left & right & up & dn) : (cur | left | right | up | dn);
    return v & cb_valid_mask(k, w);
}
