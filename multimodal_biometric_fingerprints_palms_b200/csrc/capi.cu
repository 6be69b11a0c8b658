// C ABI of libfpb200 (include/fpb200.h): handle, workspace in HBM, pipeline orchestration.
// No CPU implementation of any stage lives here: every entry point enqueues CUDA kernels and fails
// with FPB_E_CUDA when there is no usable device.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fpb200.h"
#include "../../include/fpb200_io.h"
#include <algorithm>
#include "fpb_kernels.h"
#include "fpb_jpeg.h"

static char g_create_error[512] = "";

#define FPB_MAX_SPLIT 10
struct fpb_handle {
    int device, maxB, H, W;
    cudaStream_t st; bool own_stream;
    long long launches;
    char err[512];
    int last_n;
    FpbPost post;
    // ---- device workspace
    uint8_t* u8pool; float* f32pool; int* i32pool;
    uint8_t *in, *normalized, *nlm, *denoised, *eq, *blur, *segmented, *mask, *img_eq, *bin0, *bA, *bB, *bC,
            *binary, *smooth, *gate, *skeleton, *aux_u8, *skel_file;
    int raw_cap;                 // raw minutiae kept per image (fpb_raw_cap_for(H, W)); more is FPB_E_OVERFLOW
    int32_t* stage_wh; int stage_n;   // per-image (w', h') of the next stage-entry calls (fpb_set_stage_dims), NULL = H x W
    float rel_thresh;            // thinning_and_cleaning(rel_thresh): 0.1 on the reference's path (fpb_set_rel_threshold)
    int handoff;                 // 1 = K8/K9 read the skeleton through the reference's JPEG file hand-off (default), 0 = in memory
    float *t[6], *orient_img, *rel_img, *skel_orient, *skel_coher, *dens, *orient_blocks, *skel_blocks;
    int *labels, *sizes;
    unsigned *hist, *stdmax, *dmax;
    uint8_t *lut, *tilelut, *thin_table;
    float *flut, *blk, *blk_rel, *blk_scratch;
    double *pct, *post_scratch;
    int* post_idx;
    cudaStream_t split_st[FPB_MAX_SPLIT]; cudaEvent_t split_ev[FPB_MAX_SPLIT + 1]; int split;   // number of sub-batches of the fused run
    int4* roi;
    int *raw_count, *out_count;
    uint32_t *raw, *bitscratch;
    FpbMinutiaDev* out;
    // ---- pinned host results
    int4* h_roi; int *h_raw_count, *h_out_count; uint32_t* h_raw; FpbMinutiaDev* h_out;
    bool results_valid, raw_valid;
    // ---- optional stage timing (CUDA events on h->st)
    bool profile; cudaEvent_t ev[12];
    FpbProf prof;
    // ---- EXTENSION (rows G1/G2): allocated by fpb_enable_enhanced / fpb_enhance_gabor
    bool gabor_on, gabor_ready; FpbGaborParams gabor_prm; FpbGaborBank gabor_bank;
    uint8_t* enhanced; float *gabor_resp, *freq_blocks;
    // ---- JPEG input path (fpb200_io.h): coefficient staging, allocated by the first fpb_decode_jpeg_batch
    int16_t *h_coef, *d_coef; uint16_t *h_qt, *d_qt;
};
#define FPB_NSTAGES 9   /* K1 K2 K3 K4 K5 K6 K7+K8 K9 | NLM kernel alone */

static int fail(fpb_handle* h, int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(h ? h->err : g_create_error, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(h, FPB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

static const uint8_t* zhang_suen_table() {
    // Zhang & Suen (CACM 1984) in scikit-image's coding: NW=1 N=2 NE=4 E=8 SE=16 S=32 SW=64 W=128;
    // 2 <= B <= 6, A == 1; pass 1: N*E*S == 0 && E*S*W == 0 ; pass 2: N*E*W == 0 && N*S*W == 0
    static uint8_t tab[256];
    static bool done = false;
    if (!done) {
        const int bit[8] = {2, 4, 8, 16, 32, 64, 128, 1};       // N NE E SE S SW W NW = P2..P9
        for (int c = 0; c < 256; ++c) {
            int p[8], B = 0, A = 0;
            for (int i = 0; i < 8; ++i) { p[i] = (c & bit[i]) ? 1 : 0; B += p[i]; }
            for (int i = 0; i < 8; ++i) A += (p[i] == 0 && p[(i + 1) & 7] == 1);
            uint8_t v = 0;
            if (B >= 2 && B <= 6 && A == 1) {
                const int n = p[0], e = p[2], s = p[4], w = p[6];
                if (n * e * s == 0 && e * s * w == 0) v |= 1;
                if (n * e * w == 0 && n * s * w == 0) v |= 2;
            }
            tab[c] = v;
        }
        done = true;
    }
    return tab;
}

static FpbPost default_post() {
    FpbPost p; p.quality_window = 25; p.quality_threshold = 0.15; p.coherence_threshold = 0.2;
    p.min_distance = 8.0; p.margin = 30; p.max_minutiae = 60; p.patch_radius = 15;
    return p;
}

extern "C" int fpb_abi_version(void) { return FPB_ABI_VERSION; }

extern "C" const char* fpb_last_error(const fpb_handle* h) { return h ? h->err : g_create_error; }

extern "C" long long fpb_launch_count(const fpb_handle* h) { return h ? h->launches : 0; }

extern "C" void fpb_destroy(fpb_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->st) cudaStreamSynchronize(h->st);
    void* dev[] = {h->u8pool, h->f32pool, h->i32pool, h->hist, h->stdmax, h->dmax, h->lut, h->tilelut, h->thin_table,
                   h->flut, h->blk, h->pct, h->post_scratch, h->post_idx, h->roi, h->raw_count, h->out_count, h->raw, h->bitscratch, h->out};
    for (int i = 0; i < FPB_MAX_SPLIT; ++i) if (h->split_st[i]) cudaStreamDestroy(h->split_st[i]);
    for (int i = 0; i <= FPB_MAX_SPLIT; ++i) if (h->split_ev[i]) cudaEventDestroy(h->split_ev[i]);
    for (void* p : dev) if (p) cudaFree(p);
    { void* ext[] = {h->enhanced, h->gabor_resp, h->freq_blocks, h->gabor_bank.d_taps, h->gabor_bank.d_offset, h->gabor_bank.d_radius};
      for (void* p : ext) if (p) cudaFree(p); }
    if (h->d_coef) cudaFree(h->d_coef);
    if (h->d_qt) cudaFree(h->d_qt);
    if (h->h_coef) cudaFreeHost(h->h_coef);
    if (h->h_qt) cudaFreeHost(h->h_qt);
    void* host[] = {h->h_roi, h->h_raw_count, h->h_out_count, h->h_raw, h->h_out};
    for (void* p : host) if (p) cudaFreeHost(p);
    for (int i = 0; i < 12; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i <= FPB_PROF_MAX; ++i) if (h->prof.ev[i]) cudaEventDestroy(h->prof.ev[i]);
    if (h->own_stream && h->st) cudaStreamDestroy(h->st);
    free(h->stage_wh);
    delete h;
}

extern "C" int fpb_create(fpb_handle** out, int device, int max_batch, int height, int width, void* cuda_stream) {
    if (!out) return fail(nullptr, FPB_E_ARG, "fpb_create: out is NULL");
    *out = nullptr;
    if (max_batch < 1 || height < 3 || width < 3) return fail(nullptr, FPB_E_ARG, "fpb_create: bad shape %d x %d x %d", max_batch, height, width);
    if (height > 16383 || width > 16383) return fail(nullptr, FPB_E_SHAPE, "fpb_create: image larger than 16383 pixels per side");
    if (((width + 31) / 32) * (long long)height > 32768) return fail(nullptr, FPB_E_SHAPE, "fpb_create: image larger than 1024x1024 pixels is not supported by the thinning kernel");
    const long long P = (long long)height * width;
    if ((long long)max_batch * P >= (1ll << 31)) return fail(nullptr, FPB_E_SHAPE, "fpb_create: max_batch*H*W must stay below 2^31 (chunk the batch)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, FPB_E_CUDA, "fpb_create: no CUDA device (%s) - this library has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, FPB_E_ARG, "fpb_create: device %d out of range (%d devices)", device, ndev);
    fpb_handle* h = new (std::nothrow) fpb_handle();
    if (!h) return fail(nullptr, FPB_E_NOMEM, "fpb_create: out of host memory");
    memset(h, 0, sizeof(*h));
    h->device = device; h->maxB = max_batch; h->H = height; h->W = width; h->post = default_post();
    h->raw_cap = fpb_raw_cap_for(height, width); h->handoff = 1; h->rel_thresh = 0.1f;
#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
        fail(nullptr, FPB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); fpb_destroy(h); return FPB_E_CUDA; } } while (0)
    CUC(cudaSetDevice(device));
    if (cuda_stream) { h->st = (cudaStream_t)cuda_stream; h->own_stream = false; }
    else { CUC(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking)); h->own_stream = true; }
    const size_t B = (size_t)max_batch, NP = B * (size_t)P;
    const int NU8 = 19, NF32 = 11;
    CUC(cudaMalloc(&h->u8pool, NP * NU8));
    CUC(cudaMalloc(&h->f32pool, NP * NF32 * sizeof(float)));
    CUC(cudaMalloc(&h->i32pool, NP * 2 * sizeof(int)));
    uint8_t** u8s[] = {&h->in, &h->normalized, &h->nlm, &h->denoised, &h->eq, &h->blur, &h->segmented, &h->mask, &h->img_eq,
                       &h->bin0, &h->bA, &h->bB, &h->bC, &h->binary, &h->smooth, &h->gate, &h->skeleton, &h->aux_u8, &h->skel_file};
    for (int i = 0; i < NU8; ++i) *u8s[i] = h->u8pool + NP * i;
    float** f32s[] = {&h->t[0], &h->t[1], &h->t[2], &h->t[3], &h->t[4], &h->t[5], &h->orient_img, &h->rel_img,
                      &h->skel_orient, &h->skel_coher, &h->dens};
    for (int i = 0; i < NF32; ++i) *f32s[i] = h->f32pool + NP * i;
    h->labels = h->i32pool; h->sizes = h->i32pool + NP;
    const size_t NB = (size_t)(width / 16) * (height / 16) + 1;
    CUC(cudaMalloc(&h->hist, B * 256 * sizeof(unsigned)));
    CUC(cudaMalloc(&h->stdmax, B * sizeof(unsigned)));
    CUC(cudaMalloc(&h->dmax, B * sizeof(unsigned)));
    CUC(cudaMalloc(&h->lut, B * 256));
    CUC(cudaMalloc(&h->tilelut, B * 64 * 256));
    CUC(cudaMalloc(&h->thin_table, 260));                 // 256 entries + [256] = 1 when they are the built-in Zhang-Suen table
    CUC(cudaMalloc(&h->flut, B * 256 * sizeof(float)));
    CUC(cudaMalloc(&h->blk, B * NB * 7 * sizeof(float)));           // orient_blocks, skel_blocks, blk_rel, 4 scratch
    CUC(cudaMalloc(&h->pct, B * 2 * sizeof(double)));
    CUC(cudaMalloc(&h->post_scratch, B * FPB_POST_SCRATCH_DOUBLES(h->raw_cap) * sizeof(double)));
    CUC(cudaMalloc(&h->post_idx, B * FPB_POST_IDX_INTS(h->raw_cap) * sizeof(int)));
    for (int i = 0; i < FPB_MAX_SPLIT; ++i) CUC(cudaStreamCreateWithFlags(&h->split_st[i], cudaStreamNonBlocking));
    for (int i = 0; i <= FPB_MAX_SPLIT; ++i) CUC(cudaEventCreateWithFlags(&h->split_ev[i], cudaEventDisableTiming));
    h->split = getenv("FPB_NO_SPLIT") ? 1 : 2;
    if (getenv("FPB_SPLIT")) { const int k = atoi(getenv("FPB_SPLIT")); if (k >= 1 && k <= FPB_MAX_SPLIT) h->split = k; }
    CUC(cudaMalloc(&h->roi, B * sizeof(int4)));
    CUC(cudaMalloc(&h->raw_count, B * sizeof(int)));
    CUC(cudaMalloc(&h->out_count, B * sizeof(int)));
    CUC(cudaMalloc(&h->raw, B * (size_t)h->raw_cap * sizeof(uint32_t)));
    CUC(cudaMalloc(&h->out, B * FPB_MAX_REFINED * sizeof(FpbMinutiaDev)));
    const size_t bitwords = (size_t)((width + 31) / 32) * height;
    if (bitwords * 4 * 3 + 64 * 1024 > 200 * 1024) CUC(cudaMalloc(&h->bitscratch, B * bitwords * 3 * sizeof(uint32_t)));
    h->orient_blocks = h->blk; h->skel_blocks = h->blk + B * NB; h->blk_rel = h->blk + 2 * B * NB; h->blk_scratch = h->blk + 3 * B * NB;
    CUC(cudaMallocHost(&h->h_roi, B * sizeof(int4)));
    CUC(cudaMallocHost(&h->h_raw_count, B * sizeof(int)));
    CUC(cudaMallocHost(&h->h_out_count, B * sizeof(int)));
    CUC(cudaMallocHost(&h->h_raw, B * (size_t)h->raw_cap * sizeof(uint32_t)));
    CUC(cudaMallocHost(&h->h_out, B * FPB_MAX_REFINED * sizeof(FpbMinutiaDev)));
    {
        static uint8_t tab257[260];
        memcpy(tab257, zhang_suen_table(), 256); tab257[256] = 1;
        CUC(cudaMemcpyAsync(h->thin_table, tab257, 260, cudaMemcpyHostToDevice, h->st));
    }
    fpb_upload_nlm_table(h->st);
    CUC(cudaStreamSynchronize(h->st));
    CUC(cudaGetLastError());
#undef CUC
    *out = h;
    return FPB_OK;
}

extern "C" int fpb_set_profiling(fpb_handle* h, int on) {
    if (!h) return FPB_E_ARG;
    CU(h, cudaSetDevice(h->device));
    if (on && !h->ev[0]) for (int i = 0; i < 12; ++i) CU(h, cudaEventCreate(&h->ev[i]));
    h->profile = on != 0;
    return FPB_OK;
}

// per-launch device times of the last fpb_run_* call as text lines "<source file>:<line> <ms>" (launch order)
extern "C" int fpb_kernel_times(fpb_handle* h, int enable, char* buf, int cap) {
    if (!h) return FPB_E_ARG;
    CU(h, cudaSetDevice(h->device));
    if (enable && !h->prof.ev[0]) for (int i = 0; i <= FPB_PROF_MAX; ++i) CU(h, cudaEventCreate(&h->prof.ev[i]));
    int written = 0;
    if (buf && cap > 0 && h->prof.on && h->prof.n > 0) {
        CU(h, cudaStreamSynchronize(h->st));
        for (int i = 0; i < h->prof.n; ++i) {
            float ms = 0.f;
            CU(h, cudaEventElapsedTime(&ms, h->prof.ev[i], h->prof.ev[i + 1]));
            const char* f = strrchr(h->prof.file[i], '/'); f = f ? f + 1 : h->prof.file[i];
            const int k = snprintf(buf + written, (size_t)(cap - written), "%s:%d %.4f\n", f, h->prof.line[i], ms);
            if (k < 0 || k >= cap - written) break;
            written += k;
        }
    }
    h->prof.on = enable != 0;
    return written;
}

extern "C" int fpb_stage_times(fpb_handle* h, float* ms, int cap) {
    if (!h || !ms) return FPB_E_ARG;
    if (!h->profile || h->last_n < 1) return fail(h, FPB_E_STATE, "profiling is off or nothing ran");
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->st));
    for (int i = 0; i < FPB_NSTAGES && i < cap; ++i) {
        if (i < 8) CU(h, cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]));
        else CU(h, cudaEventElapsedTime(&ms[i], h->ev[9], h->ev[10]));
    }
    return FPB_NSTAGES;
}

extern "C" int fpb_sync(fpb_handle* h) {
    if (!h) return FPB_E_ARG;
    CU(h, cudaStreamSynchronize(h->st));
    CU(h, cudaGetLastError());
    return FPB_OK;
}

extern "C" int fpb_set_stage_dims(fpb_handle* h, const int32_t* wh, int n) {
    if (!h) return FPB_E_ARG;
    free(h->stage_wh); h->stage_wh = nullptr; h->stage_n = 0;
    if (!wh) return FPB_OK;
    if (n < 1 || n > h->maxB) return fail(h, FPB_E_ARG, "batch %d outside [1, %d]", n, h->maxB);
    for (int i = 0; i < n; ++i)
        if (wh[2 * i] < 3 || wh[2 * i] > h->W || wh[2 * i + 1] < 3 || wh[2 * i + 1] > h->H)
            return fail(h, FPB_E_SHAPE, "image %d: %d x %d does not fit this handle's %d x %d planes", i, wh[2 * i + 1], wh[2 * i], h->H, h->W);
    h->stage_wh = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)n);
    if (!h->stage_wh) return fail(h, FPB_E_NOMEM, "out of host memory");
    memcpy(h->stage_wh, wh, sizeof(int32_t) * 2 * (size_t)n);
    h->stage_n = n;
    return FPB_OK;
}

extern "C" int fpb_raw_capacity(const fpb_handle* h) { return h ? h->raw_cap : FPB_E_ARG; }

extern "C" int fpb_set_handoff(fpb_handle* h, int mode) {
    if (!h || (mode != 0 && mode != 1)) return FPB_E_ARG;
    h->handoff = mode;
    return FPB_OK;
}

extern "C" int fpb_set_rel_threshold(fpb_handle* h, double rel_thresh) {
    if (!h || !(rel_thresh == rel_thresh)) return FPB_E_ARG;
    h->rel_thresh = (float)rel_thresh;          // the reference compares a float32 array with the scalar: float32 comparison
    return FPB_OK;
}

extern "C" int fpb_set_thin_table(fpb_handle* h, const uint8_t table[256]) {
    if (!h || !table) return FPB_E_ARG;
    CU(h, cudaSetDevice(h->device));
    // byte 256 tells the kernels whether the table is the built-in one (closed-form, word-parallel deletion test) or data
    uint8_t tab257[260];
    memcpy(tab257, table, 256); memset(tab257 + 256, 0, 4);
    tab257[256] = memcmp(table, zhang_suen_table(), 256) == 0 ? 1 : 0;
    CU(h, cudaMemcpyAsync(h->thin_table, tab257, 260, cudaMemcpyHostToDevice, h->st));
    CU(h, cudaStreamSynchronize(h->st));
    return FPB_OK;
}

extern "C" int fpb_set_post_params(fpb_handle* h, const fpb_post_params* p) {
    if (!h) return FPB_E_ARG;
    if (!p) { h->post = default_post(); return FPB_OK; }
    if (p->quality_window < 1 || p->quality_window > 33)
        return fail(h, FPB_E_ARG, "quality_window %d outside [1, 33] (the density kernel's tile)", p->quality_window);
    // the reference slices [:max_minutiae] without a bound; the result block holds FPB_MAX_REFINED entries per image, so a
    // larger request is refused instead of being clamped silently (negative = Python's "all but the last k" is not offered)
    if (p->max_minutiae < 0 || p->max_minutiae > FPB_MAX_REFINED)
        return fail(h, FPB_E_ARG, "max_minutiae %d outside [0, %d]", p->max_minutiae, FPB_MAX_REFINED);
    h->post.quality_window = p->quality_window; h->post.quality_threshold = p->quality_threshold;
    h->post.coherence_threshold = p->coherence_threshold; h->post.min_distance = p->min_distance;
    h->post.margin = p->margin; h->post.max_minutiae = p->max_minutiae; h->post.patch_radius = p->patch_radius;
    return FPB_OK;
}

// ------------------------------------------------------------------------------------------------
// stage sequences (all asynchronous on h->st)
// ------------------------------------------------------------------------------------------------
static FpbLaunch LN(fpb_handle* h) { FpbLaunch L; L.st = h->st; L.counter = &h->launches; L.prof = &h->prof; return L; }

static FpbOrientWs orient_ws(fpb_handle* h) {
    FpbOrientWs ws;
    ws.t0 = h->t[0]; ws.t1 = h->t[1]; ws.t2 = h->t[2]; ws.t3 = h->t[3]; ws.t4 = h->t[4];
    ws.hist = h->hist; ws.flut = h->flut; ws.pct = h->pct;
    ws.blk_rel = h->blk_rel; ws.blk_scratch = h->blk_scratch;
    return ws;
}

static void seq_normalize(fpb_handle* h, const uint8_t* img, int n, uint8_t* dst) {
    fpb_hist256(LN(h), img, n, h->W, h->H, nullptr, h->hist);
    fpb_stretch_lut(LN(h), h->hist, n, h->W, h->H, h->lut);
    fpb_clahe(LN(h), img, h->lut, n, h->W, h->H, nullptr, 2.5, h->tilelut, dst);
}

static void seq_denoise(fpb_handle* h, const uint8_t* src, int n, uint8_t* nlm, uint8_t* dst) {
    if (h->profile) cudaEventRecord(h->ev[9], h->st);
    fpb_nlm(LN(h), src, n, h->W, h->H, nlm);
    if (h->profile) cudaEventRecord(h->ev[10], h->st);
    fpb_gauss_u8(LN(h), nlm, n, h->W, h->H, 3, dst);
}

static void seq_segment(fpb_handle* h, const uint8_t* gray, int n) {
    fpb_clahe(LN(h), gray, nullptr, n, h->W, h->H, nullptr, 2.0, h->tilelut, h->eq);
    fpb_gauss_u8(LN(h), h->eq, n, h->W, h->H, 5, h->blur);
    fpb_segment_core(LN(h), gray, h->blur, n, h->W, h->H, h->hist, h->roi, h->segmented, h->mask, h->bitscratch, h->labels, h->sizes);
}

static void seq_binarize(fpb_handle* h, const uint8_t* img, int n, uint8_t* dst) {
    const int W = h->W, H = h->H;
    fpb_clahe(LN(h), img, nullptr, n, W, H, h->roi, 2.5, h->tilelut, h->img_eq);
    fpb_binarize_core(LN(h), h->img_eq, n, W, H, h->roi, h->t[0], h->t[1], h->stdmax, h->bin0);
    if (fpb_bin_finish(LN(h), h->bin0, n, W, H, h->roi, 80, 150, h->labels, h->sizes, dst)) return;
    fpb_remove_small(LN(h), h->bin0, n, W, H, h->roi, 1, 80, h->labels, h->sizes, h->bA);
    fpb_remove_small(LN(h), h->bA, n, W, H, h->roi, 0, 150, h->labels, h->sizes, h->bB);
    fpb_cross3(LN(h), h->bB, n, W, H, h->roi, 1, h->bC);
    fpb_cross3(LN(h), h->bC, n, W, H, h->roi, 0, h->bA);            // opened
    fpb_cross3(LN(h), h->bA, n, W, H, h->roi, 1, h->bC);            // marker
    fpb_reconstruct(LN(h), h->bA, h->bC, n, W, H, h->roi, h->labels, h->sizes, dst);
}

static void seq_orientation(fpb_handle* h, const uint8_t* img, const uint8_t* mask, int n, float* blocks,
                            float* orient_img, float* rel_img) {
    fpb_orientation_core(LN(h), img, mask, n, h->W, h->H, h->roi, orient_ws(h), blocks, orient_img, rel_img);
}

// EXTENSION rows G1/G2: block frequencies + Gabor enhancement from the K5 block orientations
static void seq_gabor(fpb_handle* h, const uint8_t* img, const uint8_t* mask, int n) {
    fpb_ridge_frequency(LN(h), img, mask, n, h->W, h->H, h->roi, h->orient_blocks, h->gabor_prm, h->freq_blocks);
    fpb_gabor_apply(LN(h), img, mask, n, h->W, h->H, h->roi, h->orient_blocks, h->freq_blocks, h->gabor_bank, h->gabor_resp, h->enhanced);
}

static void seq_smooth(fpb_handle* h, const uint8_t* binary, int n, uint8_t* dst) {
    // ux == nullptr selects the fused shared-memory kernel (k_smooth_fused); FPB_UNFUSED_SMOOTH=1 keeps the
    // five-kernel sequence for diagnostics
    static const bool unfused = getenv("FPB_UNFUSED_SMOOTH") != nullptr;
    fpb_smooth_core(LN(h), binary, n, h->W, h->H, h->roi, unfused ? h->t[0] : nullptr, h->t[1], h->t[2], h->t[3], h->t[4], dst);
}

static void seq_thin(fpb_handle* h, const uint8_t* smooth, const float* rel_img, int n, uint8_t* skeleton, bool extract) {
    const int W = h->W, H = h->H;
    {
        fpb_gaussian_f32(LN(h), rel_img, n, W, H, h->roi, 2.0, h->t[0], h->t[1]);
        FpbThinPre pre; pre.smooth = smooth; pre.rel_smooth = h->t[1]; pre.gate_out = h->gate; pre.labels = h->labels;
        pre.sizes = h->sizes; pre.thresh = h->rel_thresh; pre.min_obj = 64; pre.max_hole = 80;
        if (fpb_thin_fused(LN(h), pre, n, W, H, h->roi, h->thin_table, skeleton, extract ? h->raw_count : nullptr, h->raw, h->raw_cap)) return;
    }
    // the two component filters: one cluster launch (bands of bit rows per CTA, k_cluster.cu) when the image fits a cluster's
    // shared memory, else the per-pixel union-find passes in HBM
    if (!fpb_bin_finish_cluster(LN(h), smooth, n, W, H, h->roi, 64, 80, h->labels, h->sizes, h->bB, 8, true)) {
        fpb_remove_small(LN(h), smooth, n, W, H, h->roi, 1, 64, h->labels, h->sizes, h->bA);
        fpb_remove_small(LN(h), h->bA, n, W, H, h->roi, 0, 80, h->labels, h->sizes, h->bB);
    }
    fpb_gaussian_f32(LN(h), rel_img, n, W, H, h->roi, 2.0, h->t[0], h->t[1]);
    fpb_gate(LN(h), h->bB, h->t[1], n, W, H, h->roi, h->rel_thresh, h->gate);
    fpb_thin_extract(LN(h), h->gate, n, W, H, h->roi, h->thin_table, skeleton, extract ? h->raw_count : nullptr, h->raw, h->raw_cap, 1, h->bitscratch);
}

// `gray`: the image the orientation / coherence maps are computed from (post_processing.py:93); the reference's caller
// passes the skeleton itself (extract_features.py:92), which is `gray == nullptr` here
static void seq_post(fpb_handle* h, const uint8_t* skeleton, int n, const uint8_t* gray = nullptr) {
    const int W = h->W, H = h->H;
    fpb_density(LN(h), skeleton, n, W, H, h->roi, h->post.quality_window, h->dens, h->dmax);
    seq_orientation(h, gray ? gray : skeleton, nullptr, n, h->skel_blocks, h->skel_orient, h->skel_coher);
    fpb_postprocess_core(LN(h), skeleton, h->dens, h->dmax, h->skel_orient, h->skel_coher, n, W, H, h->roi,
                         h->raw_count, h->raw, h->raw_cap, h->post, h->out_count, h->out, h->post_scratch, h->post_idx);
}

// roi of a stage-entry call: the whole H x W plane, or the top-left w' x h' region fpb_set_stage_dims declared
static int set_full_roi(fpb_handle* h, int n) {
    if (h->stage_wh && n > h->stage_n) return fail(h, FPB_E_ARG, "fpb_set_stage_dims covered %d images, this call has %d", h->stage_n, n);
    for (int i = 0; i < n; ++i)
        h->h_roi[i] = h->stage_wh ? make_int4(0, 0, h->stage_wh[2 * i], h->stage_wh[2 * i + 1]) : make_int4(0, 0, h->W, h->H);
    CU(h, cudaMemcpyAsync(h->roi, h->h_roi, (size_t)n * sizeof(int4), cudaMemcpyHostToDevice, h->st));
    return FPB_OK;
}

static int check_n(fpb_handle* h, int n, const void* p) {
    if (!h) return FPB_E_ARG;
    if (!p) return fail(h, FPB_E_ARG, "null buffer");
    if (n < 1 || n > h->maxB) return fail(h, FPB_E_ARG, "batch %d outside [1, %d]", n, h->maxB);
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return fail(h, FPB_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return FPB_OK;
}

// entries that work on whole H x W frames (K1..K3, the fused run) refuse to run while crop dimensions are declared
static int require_full_frames(fpb_handle* h) {
    if (h->stage_wh) return fail(h, FPB_E_STATE, "fpb_set_stage_dims is active: this entry point works on whole %d x %d images (reset with NULL)", h->H, h->W);
    return FPB_OK;
}

static int finish(fpb_handle* h) {
    CU(h, cudaStreamSynchronize(h->st));
    CU(h, cudaGetLastError());
    return FPB_OK;
}

#define PLANE_BYTES(h, n) ((size_t)(n) * (h)->H * (h)->W)
#define H2D(h, dst, src, bytes) CU(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (h)->st))
#define D2H(h, dst, src, bytes) CU(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (h)->st))

// ------------------------------------------------------------------------------------------------
// whole path
// ------------------------------------------------------------------------------------------------
static void run_all_one(fpb_handle* h, const uint8_t* d_img, int n) {
#define MARK(i) do { if (h->profile) cudaEventRecord(h->ev[i], h->st); } while (0)
    if (h->prof.on) { h->prof.n = 0; cudaEventRecord(h->prof.ev[0], h->st); }
    MARK(0); seq_normalize(h, d_img, n, h->normalized);
    MARK(1); seq_denoise(h, h->normalized, n, h->nlm, h->denoised);
    MARK(2); seq_segment(h, h->denoised, n);
    MARK(3); seq_binarize(h, h->segmented, n, h->binary);
    MARK(4); seq_orientation(h, h->segmented, h->mask, n, h->orient_blocks, h->orient_img, h->rel_img);
    if (h->gabor_on) seq_gabor(h, h->segmented, h->mask, n);
    MARK(5); seq_smooth(h, h->binary, n, h->smooth);
    // K8/K9 input.  The reference hands the skeleton over as a quality-95 JPEG FILE (run_preprocessing.py:137-140 ->
    // extract_features.py:83): extract_minutiae thresholds the decoded grey levels at 127 and postprocess_minutiae
    // (density of `> 0`, orientation map of the grey values) sees the codec's ringing.  `handoff` reproduces that file
    // on the device (k_jpeg_roundtrip: bit-identical to cv2.imwrite + cv2.imread); 0 keeps the skeleton in memory.
    MARK(6); seq_thin(h, h->smooth, h->rel_img, n, h->skeleton, !h->handoff);
    const uint8_t* k9_in = h->skeleton;
    if (h->handoff) {
        fpb_jpeg_roundtrip_q95(LN(h), h->skeleton, n, h->W, h->H, h->roi, h->skel_file);
        fpb_thin_extract(LN(h), h->skel_file, n, h->W, h->H, h->roi, h->thin_table, nullptr, h->raw_count, h->raw, h->raw_cap, 0,
                         h->bitscratch, 127);
        k9_in = h->skel_file;
    }
    MARK(7); seq_post(h, k9_in, n);
    MARK(8);
#undef MARK
}

// A view of the handle whose per-image buffers start at image `first` and whose work goes to stream `st`.
static fpb_handle make_view(const fpb_handle* h, int first, cudaStream_t st) {
    fpb_handle v = *h;
    const size_t P = (size_t)h->H * h->W, f = (size_t)first;
    const size_t NB = (size_t)(h->W / 16) * (h->H / 16) + 1;
    uint8_t** u8s[] = {&v.in, &v.normalized, &v.nlm, &v.denoised, &v.eq, &v.blur, &v.segmented, &v.mask, &v.img_eq, &v.bin0,
                       &v.bA, &v.bB, &v.bC, &v.binary, &v.smooth, &v.gate, &v.skeleton, &v.aux_u8, &v.skel_file};
    for (uint8_t** p : u8s) *p += f * P;
    float** f32s[] = {&v.t[0], &v.t[1], &v.t[2], &v.t[3], &v.t[4], &v.t[5], &v.orient_img, &v.rel_img, &v.skel_orient,
                      &v.skel_coher, &v.dens};
    for (float** p : f32s) *p += f * P;
    v.labels += f * P; v.sizes += f * P;
    v.hist += f * 256; v.stdmax += f; v.dmax += f; v.lut += f * 256; v.tilelut += f * 64 * 256; v.flut += f * 256;
    v.pct += f * 2; v.roi += f; v.raw_count += f; v.out_count += f; v.raw += f * (size_t)h->raw_cap; v.out += f * FPB_MAX_REFINED;
    v.post_scratch += f * FPB_POST_SCRATCH_DOUBLES(h->raw_cap); v.post_idx += f * FPB_POST_IDX_INTS(h->raw_cap);
    v.orient_blocks += f * (NB - 1); v.skel_blocks += f * (NB - 1); v.blk_rel += f * (NB - 1); v.blk_scratch += f * 4 * (NB - 1);
    if (v.bitscratch) v.bitscratch += f * 3 * (size_t)((h->W + 31) / 32) * h->H;
    if (v.enhanced) { v.enhanced += f * P; v.gabor_resp += f * P; v.freq_blocks += f * (NB - 1); }
    v.st = st; v.launches = 0; v.profile = false; v.prof.on = false;
    return v;
}

// The stages alternate between throughput-bound kernels (NLM fills every register file) and latency-bound ones
// (one CTA per image walking a border, thinning to convergence ...).  Two halves of the batch on two streams let one
// half's latency-bound stages run under the other half's NLM.  Stage profiling keeps the single-stream order.
// `host_img` non-null: the images are still on the host - each half copies its own slice on its own stream, so the
// second half's H2D runs under the first half's kernels.
static void run_all(fpb_handle* h, const uint8_t* d_img, int n, const uint8_t* host_img = nullptr) {
    h->last_n = n; h->results_valid = false; h->raw_valid = false;
    const size_t P = (size_t)h->H * h->W;
    // a sub-batch should still fill the machine: at least 32 images of 320 x 240, proportionally fewer large ones
    const long long min_sub = std::max<long long>(1, (32LL * 320 * 240 + (long long)P - 1) / (long long)P);
    if (h->split < 2 || h->profile || h->prof.on || n < min_sub * h->split) {
        if (host_img) { cudaMemcpyAsync(h->in, host_img, (size_t)n * P, cudaMemcpyHostToDevice, h->st); d_img = h->in; }
        run_all_one(h, d_img, n);
        return;
    }
    const int K = h->split;
    cudaEventRecord(h->split_ev[0], h->st);
    for (int k = 0; k < K; ++k) {
        const int first = (int)((long long)n * k / K), cnt = (int)((long long)n * (k + 1) / K) - first;
        fpb_handle v = make_view(h, first, h->split_st[k]);
        cudaStreamWaitEvent(v.st, h->split_ev[0], 0);
        if (host_img) {
            cudaMemcpyAsync(v.in, host_img + (size_t)first * P, (size_t)cnt * P, cudaMemcpyHostToDevice, v.st);
            run_all_one(&v, v.in, cnt);
        } else {
            run_all_one(&v, d_img + (size_t)first * P, cnt);
        }
        cudaEventRecord(h->split_ev[k + 1], v.st);
        cudaStreamWaitEvent(h->st, h->split_ev[k + 1], 0);
        h->launches += v.launches;
    }
}

extern "C" int fpb_run_device(fpb_handle* h, const uint8_t* d_images, int n) {
    int rc = check_n(h, n, d_images); if (rc) return rc;
    rc = require_full_frames(h); if (rc) return rc;
    run_all(h, d_images, n);
    CU(h, cudaGetLastError());
    return FPB_OK;
}

// The reference keeps every crossing-number minutia (extract_features.py:41-69); a list that does not fit the handle's
// capacity is an error, never a silent truncation.
static int check_raw_overflow(fpb_handle* h, int n) {
    for (int b = 0; b < n; ++b)
        if (h->h_raw_count[b] > h->raw_cap)
            return fail(h, FPB_E_OVERFLOW, "image %d has %d raw minutiae, more than this handle's capacity of %d (H*W/8): "
                        "the result would be truncated - refusing", b, h->h_raw_count[b], h->raw_cap);
    return FPB_OK;
}

static int download(fpb_handle* h, bool with_raw) {
    const int n = h->last_n;
    if (n < 1) return fail(h, FPB_E_STATE, "no run to download");
    D2H(h, h->h_roi, h->roi, (size_t)n * sizeof(int4));
    D2H(h, h->h_raw_count, h->raw_count, (size_t)n * sizeof(int));
    D2H(h, h->h_out_count, h->out_count, (size_t)n * sizeof(int));
    D2H(h, h->h_out, h->out, (size_t)n * FPB_MAX_REFINED * sizeof(FpbMinutiaDev));
    if (with_raw) D2H(h, h->h_raw, h->raw, (size_t)n * h->raw_cap * sizeof(uint32_t));
    int rc = finish(h); if (rc) return rc;
    rc = check_raw_overflow(h, n); if (rc) return rc;
    h->results_valid = true; h->raw_valid = with_raw;
    return FPB_OK;
}

extern "C" int fpb_download_results(fpb_handle* h) {
    if (!h) return FPB_E_ARG;
    CU(h, cudaSetDevice(h->device));
    return download(h, true);
}

extern "C" int fpb_download_refined(fpb_handle* h) {
    if (!h) return FPB_E_ARG;
    CU(h, cudaSetDevice(h->device));
    return download(h, false);
}

extern "C" int fpb_result_block(const fpb_handle* h, int32_t* roi4, int32_t* raw_counts, int32_t* out_counts, fpb_minutia* out, int cap) {
    if (!h || !h->results_valid) return FPB_E_STATE;
    const int n = h->last_n;
    if (roi4) memcpy(roi4, h->h_roi, (size_t)n * sizeof(int4));
    if (raw_counts) memcpy(raw_counts, h->h_raw_count, (size_t)n * sizeof(int));
    if (out_counts) memcpy(out_counts, h->h_out_count, (size_t)n * sizeof(int));
    if (out && cap > 0)
        for (int i = 0; i < n; ++i) {
            const int m = h->h_out_count[i] < cap ? h->h_out_count[i] : cap;
            memcpy(out + (size_t)i * cap, h->h_out + (size_t)i * FPB_MAX_REFINED, (size_t)m * sizeof(fpb_minutia));
        }
    return n;
}

// asynchronous form of fpb_run_host: H2D, K1..K9 and the D2H of roi / counts / refined lists are enqueued and the call returns;
// fpb_wait makes the results readable.  `images` must stay valid (and should be pinned) until fpb_wait returns.
extern "C" int fpb_run_host_async(fpb_handle* h, const uint8_t* images, int n) {
    int rc = check_n(h, n, images); if (rc) return rc;
    rc = require_full_frames(h); if (rc) return rc;
    run_all(h, nullptr, n, images);
    D2H(h, h->h_roi, h->roi, (size_t)n * sizeof(int4));
    D2H(h, h->h_raw_count, h->raw_count, (size_t)n * sizeof(int));
    D2H(h, h->h_out_count, h->out_count, (size_t)n * sizeof(int));
    D2H(h, h->h_out, h->out, (size_t)n * FPB_MAX_REFINED * sizeof(FpbMinutiaDev));
    CU(h, cudaGetLastError());
    return FPB_OK;
}

extern "C" int fpb_wait(fpb_handle* h) {
    if (!h) return FPB_E_ARG;
    if (h->last_n < 1) return fail(h, FPB_E_STATE, "nothing was enqueued");
    CU(h, cudaSetDevice(h->device));
    int rc = finish(h); if (rc) return rc;
    rc = check_raw_overflow(h, h->last_n); if (rc) return rc;
    h->results_valid = true; h->raw_valid = false;
    return FPB_OK;
}

extern "C" int fpb_run_host(fpb_handle* h, const uint8_t* images, int n) {
    int rc = check_n(h, n, images); if (rc) return rc;
    rc = require_full_frames(h); if (rc) return rc;
    run_all(h, nullptr, n, images);
    return download(h, false);
}

extern "C" int fpb_result_roi(const fpb_handle* h, int image, int32_t roi[4]) {
    if (!h || !roi || !h->results_valid || image < 0 || image >= h->last_n) return FPB_E_STATE;
    const int4 r = h->h_roi[image];
    roi[0] = r.x; roi[1] = r.y; roi[2] = r.z; roi[3] = r.w;
    return FPB_OK;
}

extern "C" int fpb_result_raw(const fpb_handle* h, int image, int32_t* xyt, int cap) {
    if (!h || !h->results_valid || image < 0 || image >= h->last_n) return FPB_E_STATE;
    const int cnt = h->h_raw_count[image];
    if (xyt && cap > 0) {
        if (!h->raw_valid) return FPB_E_STATE;
        const int m = cnt < cap ? cnt : cap;
        for (int i = 0; i < m && i < h->raw_cap; ++i) {
            const uint32_t pk = h->h_raw[(size_t)image * h->raw_cap + i];
            xyt[3 * i] = pk & 0x3FFF; xyt[3 * i + 1] = (pk >> 14) & 0x3FFF; xyt[3 * i + 2] = (pk >> 28) & 1;
        }
    }
    return cnt;
}

extern "C" int fpb_result_minutiae(const fpb_handle* h, int image, fpb_minutia* out, int cap) {
    if (!h || !h->results_valid || image < 0 || image >= h->last_n) return FPB_E_STATE;
    const int cnt = h->h_out_count[image];
    if (out && cap > 0) {
        const int m = cnt < cap ? cnt : cap;
        memcpy(out, h->h_out + (size_t)image * FPB_MAX_REFINED, (size_t)m * sizeof(fpb_minutia));
    }
    return cnt;
}

extern "C" int fpb_fetch_plane(fpb_handle* h, int plane_id, void* dst, size_t bytes) {
    if (!h || !dst) return FPB_E_ARG;
    if (h->last_n < 1) return fail(h, FPB_E_STATE, "no run to fetch from");
    const void* src = nullptr; size_t el = 1;
    switch (plane_id) {
        case FPB_PLANE_NORMALIZED: src = h->normalized; break;
        case FPB_PLANE_DENOISED: src = h->denoised; break;
        case FPB_PLANE_SEGMENTED: src = h->segmented; break;
        case FPB_PLANE_MASK: src = h->mask; break;
        case FPB_PLANE_BINARY: src = h->binary; break;
        case FPB_PLANE_SMOOTH: src = h->smooth; break;
        case FPB_PLANE_SKELETON: src = h->skeleton; break;
        case FPB_PLANE_GATE: src = h->gate; break;
        case FPB_PLANE_NLM: src = h->nlm; break;
        case FPB_PLANE_ORIENT: src = h->orient_img; el = 4; break;
        case FPB_PLANE_RELIAB: src = h->rel_img; el = 4; break;
        case FPB_PLANE_SKEL_ORIENT: src = h->skel_orient; el = 4; break;
        case FPB_PLANE_SKEL_COHER: src = h->skel_coher; el = 4; break;
        case FPB_PLANE_DENSITY: src = h->dens; el = 4; break;
        case FPB_PLANE_ENHANCED: src = h->enhanced; break;
        case FPB_PLANE_GABOR: src = h->gabor_resp; el = 4; break;
        case FPB_PLANE_SKELETON_FILE: src = h->handoff ? h->skel_file : nullptr; break;
        default: return fail(h, FPB_E_ARG, "unknown plane id %d", plane_id);
    }
    if (!src) return fail(h, FPB_E_STATE, "plane %d is not available (fpb_enable_enhanced not called?)", plane_id);
    if (bytes != PLANE_BYTES(h, h->last_n) * el) return fail(h, FPB_E_ARG, "fetch_plane: expected %zu bytes", PLANE_BYTES(h, h->last_n) * el);
    CU(h, cudaSetDevice(h->device));
    D2H(h, dst, src, bytes);
    return finish(h);
}

// ------------------------------------------------------------------------------------------------
// stage entry points (host buffers, synchronous)
// ------------------------------------------------------------------------------------------------
extern "C" int fpb_normalize(fpb_handle* h, const uint8_t* img, int n, uint8_t* out) {
    int rc = check_n(h, n, img); if (rc) return rc;
    rc = require_full_frames(h); if (rc) return rc;
    if (!out) return fail(h, FPB_E_ARG, "null output");
    H2D(h, h->in, img, PLANE_BYTES(h, n));
    seq_normalize(h, h->in, n, h->normalized);
    D2H(h, out, h->normalized, PLANE_BYTES(h, n));
    return finish(h);
}

extern "C" int fpb_denoise(fpb_handle* h, const uint8_t* img, int n, uint8_t* out, uint8_t* nlm_out) {
    int rc = check_n(h, n, img); if (rc) return rc;
    rc = require_full_frames(h); if (rc) return rc;
    if (!out) return fail(h, FPB_E_ARG, "null output");
    H2D(h, h->in, img, PLANE_BYTES(h, n));
    seq_denoise(h, h->in, n, h->nlm, h->denoised);
    D2H(h, out, h->denoised, PLANE_BYTES(h, n));
    if (nlm_out) D2H(h, nlm_out, h->nlm, PLANE_BYTES(h, n));
    return finish(h);
}

extern "C" int fpb_segment(fpb_handle* h, const uint8_t* img, int n, uint8_t* segmented, uint8_t* mask, int32_t* roi4) {
    int rc = check_n(h, n, img); if (rc) return rc;
    rc = require_full_frames(h); if (rc) return rc;
    if (!segmented || !mask || !roi4) return fail(h, FPB_E_ARG, "null output");
    H2D(h, h->in, img, PLANE_BYTES(h, n));
    seq_segment(h, h->in, n);
    D2H(h, segmented, h->segmented, PLANE_BYTES(h, n));
    D2H(h, mask, h->mask, PLANE_BYTES(h, n));
    D2H(h, h->h_roi, h->roi, (size_t)n * sizeof(int4));
    rc = finish(h); if (rc) return rc;
    for (int i = 0; i < n; ++i) { roi4[4 * i] = h->h_roi[i].x; roi4[4 * i + 1] = h->h_roi[i].y; roi4[4 * i + 2] = h->h_roi[i].z; roi4[4 * i + 3] = h->h_roi[i].w; }
    return FPB_OK;
}

// segment_fingerprint on colour images (fingerprint_preprocess.py:94: cv2.cvtColor(img, COLOR_BGR2GRAY) first): interleaved
// [n, H, W, channels] uint8, channels 3 (BGR) or 4 (BGRA); the grey conversion runs on the device, then fpb_segment's path
extern "C" int fpb_segment_bgr(fpb_handle* h, const uint8_t* bgr, int channels, int n, uint8_t* segmented, uint8_t* mask, int32_t* roi4) {
    int rc = check_n(h, n, bgr); if (rc) return rc;
    rc = require_full_frames(h); if (rc) return rc;
    if (!segmented || !mask || !roi4) return fail(h, FPB_E_ARG, "null output");
    if (channels != 3 && channels != 4) return fail(h, FPB_E_ARG, "%d channels (cv2.COLOR_BGR2GRAY takes 3 or 4)", channels);
    uint8_t* d_bgr = nullptr;
    CU(h, cudaMalloc(&d_bgr, PLANE_BYTES(h, n) * channels));
    cudaError_t e = cudaMemcpyAsync(d_bgr, bgr, PLANE_BYTES(h, n) * channels, cudaMemcpyHostToDevice, h->st);
    if (e == cudaSuccess) {
        fpb_bgr2gray(LN(h), d_bgr, channels, PLANE_BYTES(h, n), h->in);
        seq_segment(h, h->in, n);
        e = cudaMemcpyAsync(segmented, h->segmented, PLANE_BYTES(h, n), cudaMemcpyDeviceToHost, h->st);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(mask, h->mask, PLANE_BYTES(h, n), cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->h_roi, h->roi, (size_t)n * sizeof(int4), cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(d_bgr);
    if (e != cudaSuccess) return fail(h, FPB_E_CUDA, "fpb_segment_bgr: %s", cudaGetErrorString(e));
    for (int i = 0; i < n; ++i) { roi4[4 * i] = h->h_roi[i].x; roi4[4 * i + 1] = h->h_roi[i].y; roi4[4 * i + 2] = h->h_roi[i].z; roi4[4 * i + 3] = h->h_roi[i].w; }
    return FPB_OK;
}

extern "C" int fpb_binarize(fpb_handle* h, const uint8_t* img, int n, uint8_t* out) {
    int rc = check_n(h, n, img); if (rc) return rc;
    if (!out) return fail(h, FPB_E_ARG, "null output");
    if (h->W < 13 || h->H < 13) return fail(h, FPB_E_SHAPE, "binarize needs images of at least 13x13 (25x25 window, reflect-101)");
    for (int i = 0; h->stage_wh && i < n && i < h->stage_n; ++i)
        if (h->stage_wh[2 * i] < 13 || h->stage_wh[2 * i + 1] < 13) return fail(h, FPB_E_SHAPE, "binarize needs images of at least 13x13 (25x25 window, reflect-101)");
    rc = set_full_roi(h, n); if (rc) return rc;
    H2D(h, h->in, img, PLANE_BYTES(h, n));
    seq_binarize(h, h->in, n, h->binary);
    D2H(h, out, h->binary, PLANE_BYTES(h, n));
    return finish(h);
}

extern "C" int fpb_orientation(fpb_handle* h, const uint8_t* img, const uint8_t* mask, int n,
                               float* orient_blocks, float* orient_img, float* rel_img) {
    int rc = check_n(h, n, img); if (rc) return rc;
    if (!orient_blocks || !orient_img || !rel_img) return fail(h, FPB_E_ARG, "null output");
    rc = set_full_roi(h, n); if (rc) return rc;
    H2D(h, h->in, img, PLANE_BYTES(h, n));
    if (mask) H2D(h, h->aux_u8, mask, PLANE_BYTES(h, n));
    seq_orientation(h, h->in, mask ? h->aux_u8 : nullptr, n, h->orient_blocks, h->orient_img, h->rel_img);
    const size_t nb = (size_t)(h->W / 16) * (h->H / 16);
    if (nb) D2H(h, orient_blocks, h->orient_blocks, (size_t)n * nb * sizeof(float));
    D2H(h, orient_img, h->orient_img, PLANE_BYTES(h, n) * sizeof(float));
    D2H(h, rel_img, h->rel_img, PLANE_BYTES(h, n) * sizeof(float));
    return finish(h);
}

// compute_orientation_map with its keyword arguments as data (orientation.py:9-14).  The defaults take fpb_orientation's
// path; other values use the handle's float planes plus - block sizes other than 16 - a block grid allocated for the call.
static int orientation_ex_common(fpb_handle* h, const uint8_t* img, const float* img_f32, const uint8_t* mask, int n, int block_size,
                                 double smooth_sigma, int invert_if_needed, double smooth_orientation_sigma,
                                 float* orient_blocks, float* orient_img, float* rel_img);

extern "C" int fpb_orientation_ex(fpb_handle* h, const uint8_t* img, const uint8_t* mask, int n, int block_size,
                                  double smooth_sigma, int invert_if_needed, double smooth_orientation_sigma,
                                  float* orient_blocks, float* orient_img, float* rel_img) {
    return orientation_ex_common(h, img, nullptr, mask, n, block_size, smooth_sigma, invert_if_needed, smooth_orientation_sigma,
                                 orient_blocks, orient_img, rel_img);
}

// compute_orientation_map on a non-uint8 image (orientation.py:21-24): img is [n,H,W] float32 (`img.astype(np.float32)`),
// rescaled to [0, 1] on the device when it leaves that range, as the reference does
extern "C" int fpb_orientation_f32(fpb_handle* h, const float* img, const uint8_t* mask, int n, int block_size,
                                   double smooth_sigma, int invert_if_needed, double smooth_orientation_sigma,
                                   float* orient_blocks, float* orient_img, float* rel_img) {
    return orientation_ex_common(h, nullptr, img, mask, n, block_size, smooth_sigma, invert_if_needed, smooth_orientation_sigma,
                                 orient_blocks, orient_img, rel_img);
}

static int orientation_ex_common(fpb_handle* h, const uint8_t* img, const float* img_f32, const uint8_t* mask, int n, int block_size,
                                 double smooth_sigma, int invert_if_needed, double smooth_orientation_sigma,
                                 float* orient_blocks, float* orient_img, float* rel_img) {
    int rc = check_n(h, n, img ? (const void*)img : (const void*)img_f32); if (rc) return rc;
    if (!orient_blocks || !orient_img || !rel_img) return fail(h, FPB_E_ARG, "null output");
    if (block_size < 1 || block_size > h->W || block_size > h->H)
        return fail(h, FPB_E_SHAPE, "block_size %d leaves no whole block in a %d x %d image (cv2.resize of an empty grid fails in the reference)", block_size, h->H, h->W);
    // Gaussian radius int(4 sigma + 0.5) <= 31 (64-tap weight table); sigma <= 1e-15 (negative included) = not filtered, as SciPy
    const double pre = smooth_sigma / 2.0 > 0.5 ? smooth_sigma / 2.0 : 0.5;
    if (smooth_sigma >= 7.875 || smooth_orientation_sigma >= 7.875 || pre >= 7.875)
        return fail(h, FPB_E_ARG, "sigma above 7.875 (smooth_sigma %g, smooth_orientation_sigma %g): Gaussian radius over 31", smooth_sigma, smooth_orientation_sigma);
    for (int i = 0; h->stage_wh && i < n && i < h->stage_n; ++i)
        if (h->stage_wh[2 * i] < block_size || h->stage_wh[2 * i + 1] < block_size)
            return fail(h, FPB_E_SHAPE, "block_size %d leaves no whole block in image %d", block_size, i);
    rc = set_full_roi(h, n); if (rc) return rc;
    if (img) H2D(h, h->in, img, PLANE_BYTES(h, n));
    else H2D(h, h->t[5], img_f32, PLANE_BYTES(h, n) * sizeof(float));
    if (mask) H2D(h, h->aux_u8, mask, PLANE_BYTES(h, n));
    FpbOrientPrm prm; prm.block_size = block_size; prm.smooth_sigma = smooth_sigma; prm.invert_if_needed = invert_if_needed != 0;
    prm.smooth_orientation_sigma = smooth_orientation_sigma;
    const size_t nb = (size_t)(h->W / block_size) * (h->H / block_size);
    FpbOrientWs ws = orient_ws(h);
    float* blocks = h->orient_blocks; float* grid_buf = nullptr;
    if (block_size != 16) {                                      // orient_blocks, blk_rel, four scratch grids
        CU(h, cudaMalloc(&grid_buf, (size_t)n * nb * 6 * sizeof(float)));
        blocks = grid_buf; ws.blk_rel = grid_buf + (size_t)n * nb; ws.blk_scratch = grid_buf + 2 * (size_t)n * nb;
    }
    fpb_orientation_core(LN(h), h->in, mask ? h->aux_u8 : nullptr, n, h->W, h->H, h->roi, ws, blocks, h->orient_img, h->rel_img, &prm,
                         img ? nullptr : h->t[5]);
    cudaError_t e = cudaMemcpyAsync(orient_blocks, blocks, (size_t)n * nb * sizeof(float), cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(orient_img, h->orient_img, PLANE_BYTES(h, n) * sizeof(float), cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(rel_img, h->rel_img, PLANE_BYTES(h, n) * sizeof(float), cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (grid_buf) cudaFree(grid_buf);
    if (e != cudaSuccess) return fail(h, FPB_E_CUDA, "fpb_orientation_ex: %s", cudaGetErrorString(e));
    return FPB_OK;
}

// ---- EXTENSION rows G1/G2 ------------------------------------------------------------------------
static int gabor_setup(fpb_handle* h, const fpb_gabor_params* p) {
    FpbGaborParams g; g.n_orient = 16; g.min_period = 3; g.max_period = 25; g.sigma_factor = 0.45; g.radius_factor = 2.5;
    g.min_amplitude = 8.0; g.default_period = 9.0;
    if (p) { g.n_orient = p->n_orient; g.min_period = p->min_period; g.max_period = p->max_period; g.sigma_factor = p->sigma_factor;
             g.radius_factor = p->radius_factor; g.min_amplitude = p->min_amplitude; g.default_period = p->default_period; }
    if (g.n_orient < 2 || g.n_orient > 64 || g.min_period < 2 || g.max_period < g.min_period || g.max_period > 31 ||
        !(g.sigma_factor > 0.0) || !(g.radius_factor > 0.0) || !(g.default_period >= g.min_period && g.default_period <= g.max_period))
        return fail(h, FPB_E_ARG, "fpb_gabor_params: need 2 <= n_orient <= 64, 2 <= min_period <= max_period <= 31, positive factors, default_period inside the range");
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->st));
    const size_t NP = (size_t)h->maxB * h->H * h->W, NB = (size_t)(h->W / 16) * (h->H / 16) + 1;
    if (!h->enhanced) {
        CU(h, cudaMalloc(&h->enhanced, NP));
        CU(h, cudaMalloc(&h->gabor_resp, NP * sizeof(float)));
        CU(h, cudaMalloc(&h->freq_blocks, (size_t)h->maxB * NB * sizeof(float)));
        CU(h, cudaMemset(h->freq_blocks, 0, (size_t)h->maxB * NB * sizeof(float)));
    }
    std::vector<float> taps; std::vector<int> off, rad;
    const int rmax = fpb_gabor_build_bank(g, taps, off, rad);
    FpbGaborBank& b = h->gabor_bank;
    if (b.d_taps) { cudaFree(b.d_taps); cudaFree(b.d_offset); cudaFree(b.d_radius); b.d_taps = nullptr; b.d_offset = nullptr; b.d_radius = nullptr; }
    CU(h, cudaMalloc(&b.d_taps, taps.size() * sizeof(float)));
    CU(h, cudaMalloc(&b.d_offset, off.size() * sizeof(int)));
    CU(h, cudaMalloc(&b.d_radius, rad.size() * sizeof(int)));
    CU(h, cudaMemcpy(b.d_taps, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(b.d_offset, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(b.d_radius, rad.data(), rad.size() * sizeof(int), cudaMemcpyHostToDevice));
    b.n_orient = g.n_orient; b.pmin = g.min_period; b.pmax = g.max_period; b.rmax = rmax;
    h->gabor_prm = g; h->gabor_ready = true;
    return FPB_OK;
}

extern "C" int fpb_enable_enhanced(fpb_handle* h, const fpb_gabor_params* p) {
    if (!h) return FPB_E_ARG;
    const int rc = gabor_setup(h, p);
    if (rc) return rc;
    h->gabor_on = true;
    return FPB_OK;
}

extern "C" int fpb_disable_enhanced(fpb_handle* h) {
    if (!h) return FPB_E_ARG;
    h->gabor_on = false;
    return FPB_OK;
}

extern "C" int fpb_enhance_gabor(fpb_handle* h, const uint8_t* img, const uint8_t* mask, int n, const fpb_gabor_params* p,
                                 float* freq_blocks, float* response, uint8_t* enhanced) {
    int rc = check_n(h, n, img); if (rc) return rc;
    if (!enhanced) return fail(h, FPB_E_ARG, "null output");
    if (p || !h->gabor_ready) { rc = gabor_setup(h, p); if (rc) return rc; }
    rc = set_full_roi(h, n); if (rc) return rc;
    H2D(h, h->in, img, PLANE_BYTES(h, n));
    if (mask) H2D(h, h->aux_u8, mask, PLANE_BYTES(h, n));
    const uint8_t* dm = mask ? h->aux_u8 : nullptr;
    seq_orientation(h, h->in, dm, n, h->orient_blocks, h->orient_img, h->rel_img);
    seq_gabor(h, h->in, dm, n);
    h->last_n = n;
    const size_t nb = (size_t)(h->W / 16) * (h->H / 16);
    if (freq_blocks && nb) D2H(h, freq_blocks, h->freq_blocks, (size_t)n * nb * sizeof(float));
    if (response) D2H(h, response, h->gabor_resp, PLANE_BYTES(h, n) * sizeof(float));
    D2H(h, enhanced, h->enhanced, PLANE_BYTES(h, n));
    return finish(h);
}

extern "C" int fpb_fetch_freq_blocks(fpb_handle* h, float* dst, size_t bytes) {
    if (!h || !dst) return FPB_E_ARG;
    if (!h->freq_blocks || h->last_n < 1) return fail(h, FPB_E_STATE, "no block frequencies (fpb_enable_enhanced not called?)");
    const size_t nb = (size_t)(h->W / 16) * (h->H / 16);
    if (bytes != (size_t)h->last_n * nb * sizeof(float)) return fail(h, FPB_E_ARG, "fetch_freq_blocks: expected %zu bytes", (size_t)h->last_n * nb * sizeof(float));
    CU(h, cudaSetDevice(h->device));
    if (nb) D2H(h, dst, h->freq_blocks, bytes);
    return finish(h);
}

extern "C" int fpb_smooth(fpb_handle* h, const uint8_t* binary, int n, uint8_t* out) {
    int rc = check_n(h, n, binary); if (rc) return rc;
    if (!out) return fail(h, FPB_E_ARG, "null output");
    rc = set_full_roi(h, n); if (rc) return rc;
    H2D(h, h->in, binary, PLANE_BYTES(h, n));
    seq_smooth(h, h->in, n, h->smooth);
    D2H(h, out, h->smooth, PLANE_BYTES(h, n));
    return finish(h);
}

// smooth_fingerprint_skeleton with its keyword arguments as data (fingerprint_preprocess.py:141-144): the unfused kernel
// sequence (k_sm_init, diffusion_iter x k_sm_step, Gaussian 0.6, k_sm_final); fpb_smooth's fused kernel holds the defaults.
extern "C" int fpb_smooth_ex(fpb_handle* h, const uint8_t* binary, int n, double sigma, int diffusion_iter,
                             double contrast_boost, uint8_t* out) {
    int rc = check_n(h, n, binary); if (rc) return rc;
    if (!out) return fail(h, FPB_E_ARG, "null output");
    if (diffusion_iter < 0) diffusion_iter = 0;                  // range(negative) is empty
    rc = set_full_roi(h, n); if (rc) return rc;
    H2D(h, h->in, binary, PLANE_BYTES(h, n));
    FpbSmoothPrm prm; prm.sigma = sigma; prm.diffusion_iter = diffusion_iter; prm.contrast_boost = contrast_boost;
    fpb_smooth_core(LN(h), h->in, n, h->W, h->H, h->roi, h->t[0], h->t[1], h->t[2], h->t[3], h->t[4], h->smooth, &prm);
    D2H(h, out, h->smooth, PLANE_BYTES(h, n));
    return finish(h);
}

extern "C" int fpb_thin(fpb_handle* h, const uint8_t* binary_smooth, const float* reliability, int n,
                        uint8_t* skeleton, uint8_t* gate_out) {
    int rc = check_n(h, n, binary_smooth); if (rc) return rc;
    if (!reliability || !skeleton) return fail(h, FPB_E_ARG, "null buffer");
    rc = set_full_roi(h, n); if (rc) return rc;
    H2D(h, h->in, binary_smooth, PLANE_BYTES(h, n));
    H2D(h, h->rel_img, reliability, PLANE_BYTES(h, n) * sizeof(float));
    seq_thin(h, h->in, h->rel_img, n, h->skeleton, false);
    D2H(h, skeleton, h->skeleton, PLANE_BYTES(h, n));
    if (gate_out) D2H(h, gate_out, h->gate, PLANE_BYTES(h, n));
    return finish(h);
}

extern "C" int fpb_skeletonize(fpb_handle* h, const uint8_t* gate, int n, uint8_t* skeleton) {
    int rc = check_n(h, n, gate); if (rc) return rc;
    if (!skeleton) return fail(h, FPB_E_ARG, "null output");
    rc = set_full_roi(h, n); if (rc) return rc;
    H2D(h, h->gate, gate, PLANE_BYTES(h, n));
    fpb_thin_extract(LN(h), h->gate, n, h->W, h->H, h->roi, h->thin_table, h->skeleton, nullptr, h->raw, h->raw_cap, 1, h->bitscratch);
    D2H(h, skeleton, h->skeleton, PLANE_BYTES(h, n));
    return finish(h);
}

extern "C" int fpb_extract_minutiae(fpb_handle* h, const uint8_t* skeleton, int n, int32_t* counts, int32_t* xyt, int cap) {
    int rc = check_n(h, n, skeleton); if (rc) return rc;
    if (!counts || (cap > 0 && !xyt)) return fail(h, FPB_E_ARG, "null output");
    rc = set_full_roi(h, n); if (rc) return rc;
    H2D(h, h->in, skeleton, PLANE_BYTES(h, n));
    fpb_thresh_u8(LN(h), h->in, n, h->W, h->H, h->roi, 127, h->gate);          // clean_skeleton: skel > 127
    fpb_thin_extract(LN(h), h->gate, n, h->W, h->H, h->roi, h->thin_table, nullptr, h->raw_count, h->raw, h->raw_cap, 0, h->bitscratch);
    D2H(h, h->h_raw_count, h->raw_count, (size_t)n * sizeof(int));
    D2H(h, h->h_raw, h->raw, (size_t)n * h->raw_cap * sizeof(uint32_t));
    rc = finish(h); if (rc) return rc;
    rc = check_raw_overflow(h, n); if (rc) return rc;
    for (int b = 0; b < n; ++b) {
        counts[b] = h->h_raw_count[b];
        const int m = counts[b] < cap ? counts[b] : cap;
        for (int i = 0; i < m && i < h->raw_cap; ++i) {
            const uint32_t pk = h->h_raw[(size_t)b * h->raw_cap + i];
            int32_t* o = xyt + ((size_t)b * cap + i) * 3;
            o[0] = pk & 0x3FFF; o[1] = (pk >> 14) & 0x3FFF; o[2] = (pk >> 28) & 1;
        }
    }
    return FPB_OK;
}

static int postprocess_common(fpb_handle* h, const uint8_t* skeleton, const uint8_t* gray, int n, const int32_t* counts,
                              const int32_t* xyt, int cap, int32_t* out_counts, fpb_minutia* out, int cap_out);

extern "C" int fpb_postprocess(fpb_handle* h, const uint8_t* skeleton, int n, const int32_t* counts,
                               const int32_t* xyt, int cap, int32_t* out_counts, fpb_minutia* out, int cap_out) {
    return postprocess_common(h, skeleton, nullptr, n, counts, xyt, cap, out_counts, out, cap_out);
}

// postprocess_minutiae(minutiae, skel, gray): orientation / coherence from `gray` (uint8, same planes as the skeleton),
// density and the intensity term from the skeleton (post_processing.py:85-93, 113).  gray == NULL: fpb_postprocess.
extern "C" int fpb_postprocess_gray(fpb_handle* h, const uint8_t* skeleton, const uint8_t* gray, int n, const int32_t* counts,
                                    const int32_t* xyt, int cap, int32_t* out_counts, fpb_minutia* out, int cap_out) {
    return postprocess_common(h, skeleton, gray, n, counts, xyt, cap, out_counts, out, cap_out);
}

static int postprocess_common(fpb_handle* h, const uint8_t* skeleton, const uint8_t* gray, int n, const int32_t* counts,
                              const int32_t* xyt, int cap, int32_t* out_counts, fpb_minutia* out, int cap_out) {
    int rc = check_n(h, n, skeleton); if (rc) return rc;
    if (!counts || !out_counts || (cap > 0 && !xyt) || (cap_out > 0 && !out)) return fail(h, FPB_E_ARG, "null buffer");
    rc = set_full_roi(h, n); if (rc) return rc;
    for (int b = 0; b < n; ++b) {
        const int m = counts[b];
        if (m < 0 || m > cap) return fail(h, FPB_E_ARG, "image %d: %d raw minutiae (cap %d)", b, m, cap);
        if (m > h->raw_cap) return fail(h, FPB_E_OVERFLOW, "image %d: %d raw minutiae exceed this handle's capacity of %d (H*W/8)", b, m, h->raw_cap);
        h->h_raw_count[b] = m;
        for (int i = 0; i < m; ++i) {
            const int32_t* q = xyt + ((size_t)b * cap + i) * 3;
            if (q[0] < 0 || q[0] >= h->W || q[1] < 0 || q[1] >= h->H) return fail(h, FPB_E_ARG, "image %d: minutia %d outside the image", b, i);
            h->h_raw[(size_t)b * h->raw_cap + i] = (uint32_t)q[0] | ((uint32_t)q[1] << 14) | ((uint32_t)(q[2] != 0) << 28);
        }
    }
    H2D(h, h->skeleton, skeleton, PLANE_BYTES(h, n));
    H2D(h, h->raw_count, h->h_raw_count, (size_t)n * sizeof(int));
    H2D(h, h->raw, h->h_raw, (size_t)n * h->raw_cap * sizeof(uint32_t));
    if (gray) H2D(h, h->aux_u8, gray, PLANE_BYTES(h, n));
    seq_post(h, h->skeleton, n, gray ? h->aux_u8 : nullptr);
    D2H(h, h->h_out_count, h->out_count, (size_t)n * sizeof(int));
    D2H(h, h->h_out, h->out, (size_t)n * FPB_MAX_REFINED * sizeof(FpbMinutiaDev));
    rc = finish(h); if (rc) return rc;
    h->last_n = n;
    for (int b = 0; b < n; ++b) {
        out_counts[b] = h->h_out_count[b];
        const int m = out_counts[b] < cap_out ? out_counts[b] : cap_out;
        if (m > 0) memcpy(out + (size_t)b * cap_out, h->h_out + (size_t)b * FPB_MAX_REFINED, (size_t)m * sizeof(fpb_minutia));
    }
    return FPB_OK;
}

// nms_adaptive (post_processing.py:10-32) / remove_redundant_oriented_adaptive (:37-64) as stand-alone calls
static int select_common(fpb_handle* h, int mode, int n, const int32_t* xy, const double* quality, const double* orientation,
                         const float* density, double p0, double p1, uint8_t* keep) {
    if (!h) return FPB_E_ARG;
    if (n == 0) return FPB_OK;
    if (n < 0 || n > h->raw_cap) return fail(h, FPB_E_OVERFLOW, "list of %d minutiae exceeds this handle's capacity of %d", n, h->raw_cap);
    if (!xy || !quality || !density || !keep || (mode == 2 && !orientation)) return fail(h, FPB_E_ARG, "null buffer");
    CU(h, cudaSetDevice(h->device));
    double* stage = (double*)malloc(sizeof(double) * 5 * (size_t)n);
    if (!stage) return fail(h, FPB_E_NOMEM, "out of host memory");
    for (int i = 0; i < n; ++i) {
        stage[i] = xy[2 * i]; stage[n + i] = xy[2 * i + 1]; stage[2 * n + i] = quality[i];
        stage[3 * n + i] = orientation ? orientation[i] : 0.0; stage[4 * n + i] = (double)density[i];
    }
    cudaError_t e = cudaMemcpyAsync(h->post_scratch, stage, sizeof(double) * 5 * (size_t)n, cudaMemcpyHostToDevice, h->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);          // pageable staging buffer: finish before free
    free(stage);
    if (e != cudaSuccess) return fail(h, FPB_E_CUDA, "upload failed: %s", cudaGetErrorString(e));
    unsigned char* d_keep = (unsigned char*)(h->post_idx + 2 * (size_t)n);
    fpb_minutiae_select(LN(h), mode, n, h->post_scratch, p0, p1, h->post_idx, d_keep);
    D2H(h, keep, d_keep, (size_t)n);
    return finish(h);
}

extern "C" int fpb_nms_adaptive(fpb_handle* h, int n, const int32_t* xy, const double* quality, const float* density,
                                double base_dist, uint8_t* keep) {
    return select_common(h, 1, n, xy, quality, nullptr, density, base_dist, 0.0, keep);
}

extern "C" int fpb_remove_redundant(fpb_handle* h, int n, const int32_t* xy, const double* quality, const double* orientation,
                                    const float* density, double base_radius, double angle_thresh, uint8_t* keep) {
    return select_common(h, 2, n, xy, quality, orientation, density, base_radius, angle_thresh, keep);
}

// ------------------------------------------------------------------------------------------------
// on-disk hand-offs (include/fpb200_io.h)
// ------------------------------------------------------------------------------------------------
static_assert(FPB_E_JPEG_FORMAT == FPB_JPEG_E_FORMAT && FPB_E_JPEG_UNSUPPORTED == FPB_JPEG_E_UNSUPPORTED &&
              FPB_E_JPEG_SHAPE == FPB_JPEG_E_SHAPE, "status codes of fpb200_io.h and fpb_jpeg.h");

extern "C" int fpb_synth_ridge(fpb_handle* h, uint64_t seed, uint64_t first_index, int n, double period, double noise_sigma) {
    if (!h) return FPB_E_ARG;
    if (n < 1 || n > h->maxB) return fail(h, FPB_E_ARG, "batch %d outside [1, %d]", n, h->maxB);
    CU(h, cudaSetDevice(h->device));
    fpb_synth_ridge_launch(LN(h), h->in, n, h->W, h->H, seed, first_index, (float)period, (float)noise_sigma, 10.0f);
    CU(h, cudaGetLastError());
    return FPB_OK;
}

extern "C" int fpb_jpeg_roundtrip(fpb_handle* h, const uint8_t* img, int n, uint8_t* out) {
    int rc = check_n(h, n, img); if (rc) return rc;
    if (!out) return fail(h, FPB_E_ARG, "null output");
    rc = set_full_roi(h, n); if (rc) return rc;
    H2D(h, h->in, img, PLANE_BYTES(h, n));
    fpb_jpeg_roundtrip_q95(LN(h), h->in, n, h->W, h->H, h->roi, h->skel_file);
    D2H(h, out, h->skel_file, PLANE_BYTES(h, n));
    return finish(h);
}

extern "C" int fpb_jpeg_info(const uint8_t* buf, size_t size, int* width, int* height, int* components) {
    FpbJpegInfo i;
    const int rc = fpb_jpeg_parse(buf, size, &i);
    if (rc) return rc;
    if (width) *width = i.width;
    if (height) *height = i.height;
    if (components) *components = i.components;
    return FPB_OK;
}

extern "C" int fpb_jpeg_coefficients(const uint8_t* buf, size_t size, int width, int height, int16_t* coefs, uint16_t* qt) {
    if (!buf || !coefs || !qt || width < 1 || height < 1) return FPB_E_ARG;
    return fpb_jpeg_entropy_decode(buf, size, width, height, coefs, qt);
}

template <class F>
static void parallel_for(int n, int threads, F f) {
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads > n) threads = n;
    std::atomic<int> next(0);
    auto work = [&]() { for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) f(i); };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
}

extern "C" int fpb_decode_jpeg_batch(fpb_handle* h, const uint8_t* const* bufs, const size_t* sizes, int n, int threads, int32_t* status) {
    int rc = check_n(h, n, bufs); if (rc) return rc;
    if (!sizes || !status) return fail(h, FPB_E_ARG, "null sizes / status");
    const int bw = (h->W + 7) / 8, bh = (h->H + 7) / 8;
    const size_t per = (size_t)bw * bh * 64;
    if (!h->d_coef) {
        CU(h, cudaMalloc(&h->d_coef, (size_t)h->maxB * per * sizeof(int16_t)));
        CU(h, cudaMalloc(&h->d_qt, (size_t)h->maxB * 64 * sizeof(uint16_t)));
        CU(h, cudaMallocHost(&h->h_coef, (size_t)h->maxB * per * sizeof(int16_t)));
        CU(h, cudaMallocHost(&h->h_qt, (size_t)h->maxB * 64 * sizeof(uint16_t)));
    }
    CU(h, cudaStreamSynchronize(h->st));                               // the staging buffers may still be in flight
    int16_t* hc = h->h_coef; uint16_t* hq = h->h_qt;
    const int W = h->W, H = h->H;
    parallel_for(n, threads, [&](int i) {
        int r = fpb_jpeg_entropy_decode(bufs[i], sizes[i], W, H, hc + (size_t)i * per, hq + (size_t)i * 64);
        if (r) {                                                       // undecodable here: a zero image, flagged
            memset(hc + (size_t)i * per, 0, per * sizeof(int16_t));
            for (int k = 0; k < 64; ++k) hq[(size_t)i * 64 + k] = 1;
        }
        status[i] = r;
    });
    H2D(h, h->d_coef, h->h_coef, (size_t)n * per * sizeof(int16_t));
    H2D(h, h->d_qt, h->h_qt, (size_t)n * 64 * sizeof(uint16_t));
    fpb_jpeg_idct(LN(h), h->d_coef, h->d_qt, n, W, H, h->in);
    CU(h, cudaGetLastError());
    int ok = 0;
    for (int i = 0; i < n; ++i) ok += status[i] == 0;
    return ok;
}

extern "C" int fpb_fetch_input(fpb_handle* h, uint8_t* dst, int n) {
    int rc = check_n(h, n, dst); if (rc) return rc;
    D2H(h, dst, h->in, PLANE_BYTES(h, n));
    return finish(h);
}

extern "C" const void* fpb_input_plane(const fpb_handle* h) { return h ? h->in : nullptr; }

extern "C" int fpb_run_decoded(fpb_handle* h, int n) {
    int rc = check_n(h, n, h); if (rc) return rc;
    // the split streams of run_all wait on an event recorded on h->st, i.e. on the IDCT kernel
    run_all(h, h->in, n);
    return download(h, false);
}

extern "C" long long fpb_minutiae_json(const fpb_minutia* m, int n, char* buf, size_t cap) {
    if (n > 0 && !m) return FPB_E_ARG;
    std::string s;
    fpb_minutiae_json_string(m, n, s);
    if (buf && cap) {
        const size_t k = s.size() < cap ? s.size() : cap;
        memcpy(buf, s.data(), k);
        if (cap > s.size()) buf[s.size()] = 0;
    }
    return (long long)s.size();
}

extern "C" int fpb_write_minutiae_json_batch(fpb_handle* h, const char* const* paths, int n, int threads) {
    if (!h || !paths) return FPB_E_ARG;
    if (!h->results_valid || n < 1 || n > h->last_n) return fail(h, FPB_E_STATE, "write_json: %d images requested, last run has %d", n, h->results_valid ? h->last_n : 0);
    std::atomic<int> written(0);
    parallel_for(n, threads, [&](int i) {
        if (!paths[i]) return;
        int cnt = h->h_out_count[i];
        if (cnt > FPB_MAX_REFINED) cnt = FPB_MAX_REFINED;
        std::string s, tmp = std::string(paths[i]) + ".tmp";
        fpb_minutiae_json_string(reinterpret_cast<const fpb_minutia*>(h->h_out + (size_t)i * FPB_MAX_REFINED), cnt, s);
        FILE* f = fopen(tmp.c_str(), "wb");
        if (!f) return;
        const bool ok = fwrite(s.data(), 1, s.size(), f) == s.size();
        if (fclose(f) == 0 && ok && rename(tmp.c_str(), paths[i]) == 0) written.fetch_add(1);
    });
    if (written.load() != n) return fail(h, FPB_E_ARG, "write_json: %d of %d files written (missing directory?)", written.load(), n);
    return n;
}
