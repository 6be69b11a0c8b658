// Large images (bit rows that do not fit one CTA's shared memory, e.g. 1024 x 1024): the K4 tail of k_bin_finish on a
// THREAD-BLOCK CLUSTER.  The five bit planes are cut into bands of 2^k rows, CTA r of the cluster keeps band r of every
// plane in its own shared memory, and the only traffic between CTAs is what the algorithms need at a band boundary:
//   * the row above / below of the 3x3 cross and of the run-adjacency test - read through distributed shared memory
//     (cluster.map_shared_rank);
//   * one integer per CTA for the run numbering (raster-order ids = a scan over the bands);
//   * the union-find of the runs: two levels - every band is labelled on its own in shared memory, then the band roots are
//     merged through a small global union-find over the runs that touch a band boundary.
// Same arithmetic as k_bin_finish (k_ccl.cu) / ccl_bits.cuh, so the result is bit-identical; the per-pixel union-find kernels
// in HBM (k_ccl_*) remain only as the path for images larger than a cluster's shared memory.
// Reference: fingerprint_preprocess.py:73-80 (remove_small_objects, remove_small_holes, opening, reconstruction).
#include "fpb_kernels.h"
#include "ccl_bits.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

struct ClBands {
    uint32_t* sm;          // this CTA's shared memory: plane p of the band at sm + p * band_words, rows contiguous
    int band_words;        // rows per band x words per row
    int rpb_log2, wpr, w, h, rank;
    __device__ __forceinline__ uint32_t* band(int p) const { return sm + p * band_words; }
    // row y of plane p, wherever it lives in the cluster
    __device__ __forceinline__ const uint32_t* row(int p, int y) const {
        uint32_t* q = sm + p * band_words + (y & ((1 << rpb_log2) - 1)) * wpr;
        const int r = y >> rpb_log2;
        return r == rank ? q : cg::this_cluster().map_shared_rank(q, r);
    }
};

__device__ __forceinline__ uint32_t cl_starts(const uint32_t* row, int k) {
    const uint32_t m = row[k];
    const uint32_t carry = k > 0 ? (row[k - 1] >> 31) : 0u;
    return m & ~((m << 1) | carry);
}

// cb_label of ccl_bits.cuh after the run numbering, cluster-collective (every thread of every CTA of the cluster calls it with the
// same arguments); union-find arrays in global memory.  The path for bands with more runs than the shared-memory scratch holds.
static __device__ void cl_label_flat(const ClBands& P, int pb, bool conn8, int pm, int pwb, int* parent, int* attr, int nruns) {
    cg::cluster_group cl = cg::this_cluster();
    const int T = blockDim.x, tid = threadIdx.x, CL = (int)cl.num_blocks(), GT = T * CL, gt = P.rank * T + tid;
    const int wpr = P.wpr, rpb = 1 << P.rpb_log2, y0 = P.rank * rpb;
    const int rows = max(0, min(P.h, y0 + rpb) - y0), nwl = rows * wpr;
    const uint32_t* bits = P.band(pb);
    const uint32_t* marker = pm >= 0 ? P.band(pm) : nullptr;
    const uint32_t* wordbase = P.band(pwb);
    for (int r = gt; r < nruns; r += GT) parent[r] = r;
    __threadfence();
    cl.sync();
    // ---- one pass over the run starts of the band: extent, attribute seed, unions with the row above
    const int c = conn8 ? 1 : 0;
    for (int i = tid; i < nwl; i += T) {
        uint32_t st = cb_starts(bits, i, i % wpr);
        if (!st) continue;
        const int yl = i / wpr, k = i - yl * wpr, y = y0 + yl;
        const uint32_t* row = bits + yl * wpr;
        int rank = 0;
        while (st) {
            const int j = __ffs(st) - 1; st &= st - 1;
            const int r = (int)wordbase[i] + rank; ++rank;
            const int s = k * 32 + j;
            int e, kk = k, hit = 0;
            {
                const uint32_t m = row[kk] >> j;
                const uint32_t inv = ~m;
                int len = __ffs(inv) - 1;
                if (j == 0 && inv == 0u) len = 32;
                if (marker) { const uint32_t rm = (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << j; hit |= (marker[yl * wpr + kk] & rm) != 0u; }
                e = s + len - 1;
                while (((e & 31) == 31) && kk + 1 < wpr && (row[kk + 1] & 1u)) {
                    ++kk;
                    const uint32_t m2 = row[kk];
                    const int l2 = (m2 == 0xFFFFFFFFu) ? 32 : (__ffs(~m2) - 1);
                    if (marker) { const uint32_t rm = l2 >= 32 ? 0xFFFFFFFFu : ((1u << l2) - 1u); hit |= (marker[yl * wpr + kk] & rm) != 0u; }
                    e += l2;
                    if (l2 < 32) break;
                }
            }
            attr[r] = marker ? hit : (e - s + 1);
            if (y == 0) continue;
            const int lo = max(s - c, 0), hi = min(e + c, P.w - 1);
            const uint32_t* prow = P.row(pb, y - 1);            // the neighbouring CTA's last row when yl == 0
            const uint32_t* pwb_row = P.row(pwb, y - 1);
            const int k1 = lo >> 5, k2 = hi >> 5;
            int first = -1, last = -1;
            for (int q = k1; q <= k2; ++q) {
                uint32_t mm = prow[q];
                if (q == k1) mm &= cb_ge_mask(lo & 31);
                if (q == k2) mm &= cb_le_mask(hi & 31);
                if (!mm) continue;
                const uint32_t pst = cl_starts(prow, q);
                if (first < 0) first = (int)pwb_row[q] + __popc(pst & cb_le_mask(__ffs(mm) - 1)) - 1;
                last = (int)pwb_row[q] + __popc(pst & cb_le_mask(31 - __clz(mm))) - 1;
            }
            for (int q = first; q >= 0 && q <= last; ++q) cb_union(parent, r, q);
        }
    }
    __threadfence();
    cl.sync();
    for (int r = gt; r < nruns; r += GT) parent[r] = cb_find(parent, r);
    __threadfence();
    cl.sync();
    for (int r = gt; r < nruns; r += GT) {
        const int root = cb_ld(parent + r);
        if (root != r) {
            const int a = cb_ld(attr + r);
            if (marker) { if (a) atomicOr(&attr[root], 1); } else atomicAdd(&attr[root], a);
        }
    }
    __threadfence();
    cl.sync();
}

// Two-level form of the same labelling, used when every band's runs fit the CTA's shared-memory scratch (2 * cap ints):
//   1. each band is labelled on its own with LOCAL union-find arrays in shared memory (a hop costs a shared-memory access, not an
//      L2 round trip - the flat version above spends its time in ~500 dependent L2 round trips per thread);
//   2. the band forests are exported to the global arrays as depth-1 trees and only the runs of a band's first row are united
//      with the last row of the band above (a few hundred global unions per boundary);
//   3. band roots are flattened / their attributes merged globally, and every run's final attribute is written back to the local
//      array, so that the selection is a shared-memory look-up.
// Afterwards  attrL[gid - base]  (gid = the run id stored in the word-base plane) is the component's pixel count / marker flag.
// Returns this band's first run id (`base`); *nruns_out = runs of the whole image.
static __device__ int cl_label_h(const ClBands& P, int pb, bool conn8, int pm, int pwb, int* parentG, int* attrG, int* parentL,
                                 int* attrL, int base, int nloc, int nruns, int* s_warp) {
    cg::cluster_group cl = cg::this_cluster();
    const int T = blockDim.x, tid = threadIdx.x;
    const int wpr = P.wpr, rpb = 1 << P.rpb_log2, y0 = P.rank * rpb;
    const int rows = max(0, min(P.h, y0 + rpb) - y0), nwl = rows * wpr;
    const uint32_t* bits = P.band(pb);
    const uint32_t* marker = pm >= 0 ? P.band(pm) : nullptr;
    const uint32_t* wordbase = P.band(pwb);
    for (int r = tid; r < nloc; r += T) parentL[r] = r;
    __syncthreads();
    const int c = conn8 ? 1 : 0;
    // ---- level 1: extent, attribute seed, unions with the row above INSIDE the band (shared memory)
    for (int i = tid; i < nwl; i += T) {
        uint32_t st = cb_starts(bits, i, i % wpr);
        if (!st) continue;
        const int yl = i / wpr, k = i - yl * wpr;
        const uint32_t* row = bits + yl * wpr;
        int rank = 0;
        while (st) {
            const int j = __ffs(st) - 1; st &= st - 1;
            const int r = (int)wordbase[i] - base + rank; ++rank;
            const int s = k * 32 + j;
            int e, kk = k, hit = 0;
            {
                const uint32_t m = row[kk] >> j;
                const uint32_t inv = ~m;
                int len = __ffs(inv) - 1;
                if (j == 0 && inv == 0u) len = 32;
                if (marker) { const uint32_t rm = (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << j; hit |= (marker[yl * wpr + kk] & rm) != 0u; }
                e = s + len - 1;
                while (((e & 31) == 31) && kk + 1 < wpr && (row[kk + 1] & 1u)) {
                    ++kk;
                    const uint32_t m2 = row[kk];
                    const int l2 = (m2 == 0xFFFFFFFFu) ? 32 : (__ffs(~m2) - 1);
                    if (marker) { const uint32_t rm = l2 >= 32 ? 0xFFFFFFFFu : ((1u << l2) - 1u); hit |= (marker[yl * wpr + kk] & rm) != 0u; }
                    e += l2;
                    if (l2 < 32) break;
                }
            }
            attrL[r] = marker ? hit : (e - s + 1);
            if (yl == 0) continue;
            const int lo = max(s - c, 0), hi = min(e + c, P.w - 1);
            const uint32_t* prow = row - wpr;
            const uint32_t* pwb_row = wordbase + (yl - 1) * wpr;
            const int k1 = lo >> 5, k2 = hi >> 5;
            int first = -1, last = -1;
            for (int q = k1; q <= k2; ++q) {
                uint32_t mm = prow[q];
                if (q == k1) mm &= cb_ge_mask(lo & 31);
                if (q == k2) mm &= cb_le_mask(hi & 31);
                if (!mm) continue;
                const uint32_t pst = cl_starts(prow, q);
                if (first < 0) first = (int)pwb_row[q] - base + __popc(pst & cb_le_mask(__ffs(mm) - 1)) - 1;
                last = (int)pwb_row[q] - base + __popc(pst & cb_le_mask(31 - __clz(mm))) - 1;
            }
            for (int q = first; q >= 0 && q <= last; ++q) cb_union(parentL, r, q);
        }
    }
    __syncthreads();
    for (int r = tid; r < nloc; r += T) parentL[r] = cb_find(parentL, r);
    __syncthreads();
    for (int r = tid; r < nloc; r += T) {
        const int root = parentL[r];
        if (root != r) { const int a = attrL[r]; if (marker) { if (a) atomicOr(&attrL[root], 1); } else atomicAdd(&attrL[root], a); }
    }
    __syncthreads();
    // ---- export the band forest (depth 1) and the band roots' attributes
    for (int r = tid; r < nloc; r += T) {
        const int root = parentL[r];
        parentG[base + r] = base + root;
        if (root == r) attrG[base + r] = attrL[r];
    }
    __threadfence();
    cl.sync();
    // ---- level 2: the first row of the band against the last row of the band above
    if (P.rank > 0 && rows > 0) {
        const uint32_t* prow = P.row(pb, y0 - 1);
        const uint32_t* pwb_row = P.row(pwb, y0 - 1);
        for (int k = tid; k < wpr; k += T) {
            uint32_t st = cb_starts(bits, k, k);
            int rank = 0;
            while (st) {
                const int j = __ffs(st) - 1; st &= st - 1;
                const int r = (int)wordbase[k] + rank; ++rank;          // global id
                const int s = k * 32 + j;
                int e, kk = k;
                {
                    const uint32_t m = bits[kk] >> j;
                    const uint32_t inv = ~m;
                    int len = __ffs(inv) - 1;
                    if (j == 0 && inv == 0u) len = 32;
                    e = s + len - 1;
                    while (((e & 31) == 31) && kk + 1 < wpr && (bits[kk + 1] & 1u)) {
                        ++kk;
                        const uint32_t m2 = bits[kk];
                        const int l2 = (m2 == 0xFFFFFFFFu) ? 32 : (__ffs(~m2) - 1);
                        e += l2;
                        if (l2 < 32) break;
                    }
                }
                const int lo = max(s - c, 0), hi = min(e + c, P.w - 1);
                const int k1 = lo >> 5, k2 = hi >> 5;
                int first = -1, last = -1;
                for (int q = k1; q <= k2; ++q) {
                    uint32_t mm = prow[q];
                    if (q == k1) mm &= cb_ge_mask(lo & 31);
                    if (q == k2) mm &= cb_le_mask(hi & 31);
                    if (!mm) continue;
                    const uint32_t pst = cl_starts(prow, q);
                    if (first < 0) first = (int)pwb_row[q] + __popc(pst & cb_le_mask(__ffs(mm) - 1)) - 1;
                    last = (int)pwb_row[q] + __popc(pst & cb_le_mask(31 - __clz(mm))) - 1;
                }
                for (int q = first; q >= 0 && q <= last; ++q) cb_union(parentG, r, q);
            }
        }
    }
    __threadfence();
    cl.sync();
    // ---- band roots: global root, attribute merge, final attribute back into the local array
    for (int r = tid; r < nloc; r += T)
        if (parentL[r] == r) {
            const int g = cb_find(parentG, base + r);
            if (g != base + r) {
                parentG[base + r] = g;
                const int a = attrL[r];
                if (marker) { if (a) atomicOr(&attrG[g], 1); } else atomicAdd(&attrG[g], a);
            }
        }
    __threadfence();
    cl.sync();
    for (int r = tid; r < nloc; r += T)
        if (parentL[r] == r) attrL[r] = cb_ld(attrG + cb_ld(parentG + base + r));
    __syncthreads();
    for (int r = tid; r < nloc; r += T) { const int root = parentL[r]; if (root != r) attrL[r] = attrL[root]; }
    __syncthreads();
    (void)nruns; (void)s_warp;
    return base;
}

// run ids of the band in raster order (cluster-wide): fills the word-base plane, returns the band's first id
static __device__ int cl_number_runs(const ClBands& P, int pb, int pwb, int* s_warp, int* s_tot, int* nloc_out, int* nruns_out, int* maxloc_out) {
    cg::cluster_group cl = cg::this_cluster();
    const int T = blockDim.x, tid = threadIdx.x, CL = (int)cl.num_blocks();
    const int wpr = P.wpr, rpb = 1 << P.rpb_log2, y0 = P.rank * rpb;
    const int rows = max(0, min(P.h, y0 + rpb) - y0), nwl = rows * wpr;
    const uint32_t* bits = P.band(pb);
    uint32_t* wordbase = P.band(pwb);
    const int cpt = (nwl + T - 1) / T;
    int local = 0;
    for (int q = 0; q < cpt; ++q) { const int i = tid * cpt + q; if (i < nwl) local += __popc(cb_starts(bits, i, i % wpr)); }
    int tot = 0;
    int excl = cb_block_scan_excl(local, s_warp, &tot);
    if (tid == 0) *s_tot = tot;
    cl.sync();
    int base = 0, nruns = 0, mx = 0;
    for (int r = 0; r < CL; ++r) { const int t = *cl.map_shared_rank(s_tot, r); if (r < P.rank) base += t; nruns += t; mx = max(mx, t); }
    int id = base + excl;
    for (int q = 0; q < cpt; ++q) {
        const int i = tid * cpt + q;
        if (i < nwl) { wordbase[i] = (uint32_t)id; id += __popc(cb_starts(bits, i, i % wpr)); }
    }
    *nloc_out = tot; *nruns_out = nruns; *maxloc_out = mx;
    __syncthreads();
    return base;
}

// out word = the pixels of word i whose run's final attribute (attrL, indexed by run id - base) passes the test
__device__ __forceinline__ uint32_t cl_select_word_h(const uint32_t* bits, const uint32_t* wordbase, const int* attrL, int base,
                                                     int i, int k, int min_size, bool keep_small) {
    uint32_t m = bits[i], out = 0;
    if (!m) return 0u;
    const uint32_t st = cb_starts(bits, i, k);
    int id = (int)wordbase[i] - base - ((m & 1u) && !(st & 1u) ? 1 : 0);
    while (m) {
        const int j = __ffs(m) - 1;
        const uint32_t low = 1u << j;
        const uint32_t rm = (m ^ (m + low)) & m;
        const bool pass = attrL[id] >= min_size;
        if (pass != keep_small) out |= rm;
        m &= ~rm;
        ++id;
    }
    return out;
}

// cb_cross_word with the rows above / below taken from wherever they live in the cluster
__device__ __forceinline__ uint32_t cl_cross_word(const ClBands& P, int p, int y, int k, bool erode) {
    const uint32_t fill = erode ? 0xFFFFFFFFu : 0u;
    const int wpr = P.wpr, w = P.w;
    auto ext = [&](const uint32_t* r, int kk) -> uint32_t {
        if (kk < 0 || kk >= wpr) return fill;
        const uint32_t vm = cb_valid_mask(kk, w);
        return (r[kk] & vm) | (fill & ~vm);
    };
    const uint32_t* row = P.row(p, y);
    const uint32_t cur = ext(row, k), prev = ext(row, k - 1), next = ext(row, k + 1);
    const uint32_t left = (cur << 1) | (prev >> 31), right = (cur >> 1) | (next << 31);
    const uint32_t up = y > 0 ? ext(P.row(p, y - 1), k) : fill, dn = y + 1 < P.h ? ext(P.row(p, y + 1), k) : fill;
    const uint32_t v = erode ? (cur & left & right & up & dn) : (cur | left | right | up | dn);
    return v & cb_valid_mask(k, w);
}

// numbering + labelling of plane pb; afterwards select with cl_pick()
struct ClLab { int base; bool local; };
static __device__ ClLab cl_label(const ClBands& P, int pb, bool conn8, int pm, int pwb, int* parentG, int* attrG, int* ufL, int cap,
                                 int* s_warp, int* s_tot) {
    int nloc, nruns, mx;
    ClLab L;
    L.base = cl_number_runs(P, pb, pwb, s_warp, s_tot, &nloc, &nruns, &mx);
    L.local = cap > 0 && mx <= cap;                     // uniform over the cluster
    if (L.local) cl_label_h(P, pb, conn8, pm, pwb, parentG, attrG, ufL, ufL + cap, L.base, nloc, nruns, s_warp);
    else cl_label_flat(P, pb, conn8, pm, pwb, parentG, attrG, nruns);
    return L;
}
__device__ __forceinline__ uint32_t cl_pick(const ClLab& L, const uint32_t* bits, const uint32_t* wb, const int* parentG, const int* attrG,
                                            const int* attrL, int i, int k, int min_size, bool keep_small) {
    return L.local ? cl_select_word_h(bits, wb, attrL, L.base, i, k, min_size, keep_small)
                   : cb_select_word(bits, wb, parentG, attrG, i, k, min_size, keep_small);
}

#define BFC_MAX_THREADS 1024
__global__ void __launch_bounds__(BFC_MAX_THREADS)
k_bin_finish_cl(const uint8_t* __restrict__ bin0, int W, int H, const int4* __restrict__ roi, int min_obj, int max_hole,
                int* __restrict__ labels, int* __restrict__ sizes, uint8_t* __restrict__ dst, int rpb_log2, int plane_words, int cap,
                int filters_only) {
    extern __shared__ __align__(16) uint32_t bfc_sm[];
    __shared__ int s_warp[33];
    __shared__ int s_tot;
    cg::cluster_group cl = cg::this_cluster();
    const int CL = (int)cl.num_blocks(), b = blockIdx.x / CL, tid = threadIdx.x, T = blockDim.x;
    const FpbDims d = fpb_dims(roi, b, W, H);
    ClBands P;
    P.sm = bfc_sm; P.rpb_log2 = rpb_log2; P.w = d.w; P.h = d.h; P.wpr = (d.w + 31) >> 5; P.rank = (int)cl.block_rank();
    P.band_words = (1 << rpb_log2) * P.wpr;
    enum { A = 0, B = 1, C = 2, M = 3, WB = 4 };
    const int wpr = P.wpr, y0 = P.rank << rpb_log2, rows = max(0, min(d.h, y0 + (1 << rpb_log2)) - y0), nwl = rows * wpr;
    int* const parent = labels + (size_t)b * W * H;
    int* const attr = sizes + (size_t)b * W * H;
    int* const ufL = reinterpret_cast<int*>(bfc_sm + 5 * (size_t)plane_words);          // local union-find scratch: 2 * cap ints
    const int* const attrL = ufL + cap;
    uint32_t *a = P.band(A), *bb = P.band(B), *cc = P.band(C), *mm = P.band(M), *wb = P.band(WB);
    if (rows > 0) cb_pack_u8(bin0 + (size_t)b * W * H + (size_t)y0 * W, W, d.w, rows, wpr, a);
    cl.sync();
    // remove_small_objects(min_obj), 4-connected
    ClLab L = cl_label(P, A, false, -1, WB, parent, attr, ufL, cap, s_warp, &s_tot);
    for (int i = tid; i < nwl; i += T) bb[i] = cl_pick(L, a, wb, parent, attr, attrL, i, i % wpr, min_obj, false);
    // remove_small_holes(max_hole): small 4-connected background components become foreground
    __syncthreads();
    for (int i = tid; i < nwl; i += T) cc[i] = ~bb[i] & cb_valid_mask(i % wpr, d.w);
    cl.sync();
    L = cl_label(P, C, false, -1, WB, parent, attr, ufL, cap, s_warp, &s_tot);
    for (int i = tid; i < nwl; i += T) a[i] = bb[i] | cl_pick(L, cc, wb, parent, attr, attrL, i, i % wpr, max_hole, true);
    uint8_t* out = dst + (size_t)b * W * H + (size_t)y0 * W;
    if (filters_only) {     // K7a of the thinning stage: remove_small_objects + remove_small_holes only (fingerprint_preprocess.py:167-168)
        __syncthreads();
        for (int y = tid >> 5; y < rows; y += T / 32)
            for (int x = tid & 31; x < d.w; x += 32)
                out[(size_t)y * W + x] = ((a[y * wpr + (x >> 5)] >> (x & 31)) & 1u) ? 255 : 0;
        cl.sync();
        return;
    }
    cl.sync();
    // opening with the cross, marker = erode(opened)
    for (int i = tid; i < nwl; i += T) bb[i] = cl_cross_word(P, A, y0 + i / wpr, i % wpr, true);
    cl.sync();
    for (int i = tid; i < nwl; i += T) cc[i] = cl_cross_word(P, B, y0 + i / wpr, i % wpr, false);
    cl.sync();
    for (int i = tid; i < nwl; i += T) mm[i] = cl_cross_word(P, C, y0 + i / wpr, i % wpr, true);
    cl.sync();
    // reconstruction by dilation: 8-connected components of `opened` that hold a marker pixel
    L = cl_label(P, C, true, M, WB, parent, attr, ufL, cap, s_warp, &s_tot);
    for (int i = tid; i < nwl; i += T) a[i] = cl_pick(L, cc, wb, parent, attr, attrL, i, i % wpr, 1, false);
    __syncthreads();
    for (int y = tid >> 5; y < rows; y += T / 32)
        for (int x = tid & 31; x < d.w; x += 32)
            out[(size_t)y * W + x] = ((a[y * wpr + (x >> 5)] >> (x & 31)) & 1u) ? 255 : 0;
    cl.sync();          // no CTA may exit while a neighbour can still read its shared memory
}

// returns false when even a cluster of eight cannot hold the image (caller falls back to the per-pixel kernels)
bool fpb_bin_finish_cluster(FpbLaunch L, const uint8_t* bin0, int n, int W, int H, const int4* roi, int min_obj, int max_hole,
                            int* labels, int* sizes, uint8_t* dst, int cluster, bool filters_only) {
    const int wpr = (W + 31) / 32;
    int rpb_log2 = 0;
    while ((cluster << rpb_log2) < H) ++rpb_log2;
    const size_t plane_words = ((size_t)1 << rpb_log2) * wpr;
    size_t smem = 5 * plane_words * sizeof(uint32_t);
    if (smem > 200 * 1024) return false;
    // the rest of the CTA's shared memory (up to 200 KB, two CTAs per SM while that keeps >= 4 k runs) holds the band's union-find
    static const bool no_local = getenv("FPB_CLUSTER_FLAT") != nullptr;
    const bool two_per_sm = smem + 2 * 4096 * 4 <= 100 * 1024;
    const size_t budget = two_per_sm ? 100 * 1024 : 200 * 1024;
    const int threads = two_per_sm ? 512 : 1024;       // one CTA per SM: twice the warps hide the same latencies
    int cap = no_local ? 0 : (int)((budget - smem) / 8);
    if (cap > 32768) cap = 32768;
    smem += (size_t)cap * 8;
    FPB_OPT_IN_SMEM(k_bin_finish_cl, 200 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n * cluster)); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = L.st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, k_bin_finish_cl, bin0, W, H, roi, min_obj, max_hole, labels, sizes, dst, rpb_log2, (int)plane_words, cap,
                           filters_only ? 1 : 0) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    LAUNCH_COUNT(L);
    return true;
}
