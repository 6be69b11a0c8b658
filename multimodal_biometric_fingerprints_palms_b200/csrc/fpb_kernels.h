// Host-side launcher declarations of the fpb200 kernels (one per stage of SURVEY.md section 8(a)).
// Every launcher enqueues on `st` and returns immediately; `L` counts kernel launches.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fpb_common.cuh"

#define FPB_PROF_MAX 256
struct FpbProf {                // optional per-launch timing: one event after every kernel launch
    bool on; int n;
    cudaEvent_t ev[FPB_PROF_MAX + 1];
    const char* file[FPB_PROF_MAX]; int line[FPB_PROF_MAX];
};

struct FpbLaunch {
    cudaStream_t st;
    long long* counter;
    FpbProf* prof;
};

static inline void fpb_mark_launch(const FpbLaunch& L, const char* file, int line) {
    if (L.counter) ++*L.counter;
    FpbProf* p = L.prof;
    if (p && p->on && p->n < FPB_PROF_MAX) {
        cudaEventRecord(p->ev[p->n + 1], L.st);
        p->file[p->n] = file; p->line[p->n] = line; ++p->n;
    }
}
#define LAUNCH_COUNT(L) fpb_mark_launch((L), __FILE__, __LINE__)

// cudaFuncSetAttribute is per device: remember which devices already have the opt-in (one process may drive several GPUs)
#define FPB_OPT_IN_SMEM(kernel, bytes) do { \
    static unsigned long long done_mask_ = 0ull; int dev_ = 0; cudaGetDevice(&dev_); \
    if (!((done_mask_ >> (dev_ & 63)) & 1ull)) { \
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); done_mask_ |= 1ull << (dev_ & 63); } \
} while (0)


// ---- k_front.cu : K1 normalise, CLAHE, K2 NLM, fixed-point Gaussians ---------------------------
void fpb_hist256(FpbLaunch L, const uint8_t* src, int n, int W, int H, const int4* roi, unsigned* hist);
void fpb_stretch_lut(FpbLaunch L, const unsigned* hist, int n, int W, int H, uint8_t* lut);
void fpb_clahe(FpbLaunch L, const uint8_t* src, const uint8_t* premap, int n, int W, int H, const int4* roi,
               double clip, uint8_t* tilelut, uint8_t* dst);
void fpb_nlm(FpbLaunch L, const uint8_t* src, int n, int W, int H, uint8_t* dst);
void fpb_gauss_u8(FpbLaunch L, const uint8_t* src, int n, int W, int H, int ntaps, uint8_t* dst);
void fpb_upload_nlm_table(cudaStream_t st);
// k_nlm_mma.cu: the same NLM as a banded Gram GEMM on tcgen05 (kind::i8, accumulators in TMEM)
bool fpb_nlm_mma(FpbLaunch L, const uint8_t* src, int n, int W, int H, uint8_t* dst);
void fpb_upload_nlm_table_mma(const int* tab, cudaStream_t st);

// ---- k_segment.cu : K3 ----------------------------------------------------------------------------
// blur = GaussianBlur5(CLAHE2.0(gray)); writes roi[b], cropped `segmented` and `mask` planes
void fpb_segment_core(FpbLaunch L, const uint8_t* gray, const uint8_t* blur, int n, int W, int H,
                      unsigned* hist, int4* roi, uint8_t* segmented, uint8_t* mask, uint32_t* bitscratch,
                      int* labels, int* sizes);

// cv2.cvtColor(COLOR_BGR2GRAY) of interleaved 8-bit pixels (ch = 3 or 4)
void fpb_bgr2gray(FpbLaunch L, const uint8_t* src, int ch, size_t npx, uint8_t* dst);

// ---- k_ccl.cu : connected components (remove_small_objects / holes, reconstruction) -------------
// dst = src with 4-connected components of `polarity` pixels smaller than min_size flipped
void fpb_remove_small(FpbLaunch L, const uint8_t* src, int n, int W, int H, const int4* roi, int polarity,
                      int min_size, int* labels, int* sizes, uint8_t* dst);
// dst = 255 on 8-connected components of src!=0 that contain a marker!=0 pixel
void fpb_reconstruct(FpbLaunch L, const uint8_t* src, const uint8_t* marker, int n, int W, int H, const int4* roi,
                     int* labels, int* flags, uint8_t* dst);

// fused K4 tail on bit rows in shared memory (false = image too large, use the kernels above)
bool fpb_bin_finish(FpbLaunch L, const uint8_t* bin0, int n, int W, int H, const int4* roi, int min_obj, int max_hole,
                    int* labels, int* sizes, uint8_t* dst);
// k_cluster.cu: the same on a thread-block cluster (bands of bit rows per CTA, distributed shared memory at the band edges)
bool fpb_bin_finish_cluster(FpbLaunch L, const uint8_t* bin0, int n, int W, int H, const int4* roi, int min_obj, int max_hole,
                            int* labels, int* sizes, uint8_t* dst, int cluster, bool filters_only = false);

// ---- k_binarize.cu : K4 -------------------------------------------------------------------------
void fpb_binarize_core(FpbLaunch L, const uint8_t* img_eq, int n, int W, int H, const int4* roi,
                       float* mean, float* stdv, unsigned* stdmax_bits, uint8_t* bin0);
void fpb_cross3(FpbLaunch L, const uint8_t* src, int n, int W, int H, const int4* roi, int erode, uint8_t* dst);

// ---- k_orient.cu : K5 ---------------------------------------------------------------------------
struct FpbOrientWs {            // float planes [n,H,W] unless noted
    float *t0, *t1, *t2, *t3, *t4;
    unsigned* hist;             // [n,256]
    float* flut;                // [n,256]
    double* pct;                // [n,2]
    float* blk_rel;             // [n, (W/16)*(H/16)]     block reliability
    float* blk_scratch;         // [n, 4, (W/16)*(H/16)]  scratch planes of the grid smoothing
};
struct FpbOrientPrm {           // compute_orientation_map's keyword arguments (orientation.py:9-14)
    int block_size; double smooth_sigma; int invert_if_needed; double smooth_orientation_sigma;
};
void fpb_orientation_core(FpbLaunch L, const uint8_t* img, const uint8_t* mask, int n, int W, int H,
                          const int4* roi, FpbOrientWs ws, float* orient_blocks, float* orient_img, float* rel_img,
                          const FpbOrientPrm* prm = nullptr, const float* img_f32 = nullptr);
// scipy gaussian_filter on an f32 plane (axis 0 then axis 1, f64 accumulation, f32 intermediate)
void fpb_gaussian_f32(FpbLaunch L, const float* src, int n, int W, int H, const int4* roi, double sigma,
                      float* tmp, float* dst);

// ---- k_smooth.cu : K6 ---------------------------------------------------------------------------
struct FpbSmoothPrm { double sigma; int diffusion_iter; double contrast_boost; };     // fingerprint_preprocess.py:141-144
void fpb_smooth_core(FpbLaunch L, const uint8_t* binary, int n, int W, int H, const int4* roi,
                     float* ux, float* uy, float* acc, float* acc2, float* tmp, uint8_t* dst,
                     const FpbSmoothPrm* prm = nullptr);

// ---- k_thin.cu : K7 gate/skeletonize/clean-up + K8 crossing numbers ------------------------------
void fpb_gate(FpbLaunch L, const uint8_t* cleaned, const float* rel_smooth, int n, int W, int H, const int4* roi,
              float thresh, uint8_t* gate);
void fpb_thresh_u8(FpbLaunch L, const uint8_t* src, int n, int W, int H, const int4* roi, int thr, uint8_t* dst);
struct FpbThinPre {             // optional fused K7a prologue of k_thin_extract
    const uint8_t* smooth;      // binary_smooth plane (non-null = run the prologue)
    const float* rel_smooth;    // gaussian_filter(reliability, 2.0)
    uint8_t* gate_out;          // optional: the mask entering skeletonize, as a {0,255} plane
    int* labels; int* sizes;    // union-find scratch, W*H ints per image each
    float thresh; int min_obj, max_hole;
    int sm_cap;                 // set by the launcher: runs that fit the shared-memory union-find scratch
};
// K7a + K7b + K8 in one kernel (false = image too large for the shared-memory path)
// raw: [n][raw_cap] packed minutiae; raw_count[b] is always the TRUE count (may exceed raw_cap: the caller must check)
bool fpb_thin_fused(FpbLaunch L, FpbThinPre pre, int n, int W, int H, const int4* roi, const uint8_t* table,
                    uint8_t* skeleton, int* raw_count, uint32_t* raw, int raw_cap);
// pack_thr: pixels > pack_thr are set (0 for {0,255} masks, 127 = clean_skeleton of extract_features.py:38-39)
void fpb_thin_extract(FpbLaunch L, const uint8_t* gate, int n, int W, int H, const int4* roi, const uint8_t* table,
                      uint8_t* skeleton, int* raw_count, uint32_t* raw, int raw_cap, int do_thin, uint32_t* bitscratch,
                      int pack_thr = 0);

// ---- k_post.cu : K9 -----------------------------------------------------------------------------
void fpb_density(FpbLaunch L, const uint8_t* skel, int n, int W, int H, const int4* roi, int win,
                 float* dens, unsigned* dmax_bits);
void fpb_postprocess_core(FpbLaunch L, const uint8_t* skel, const float* dens, const unsigned* dmax_bits,
                          const float* orient, const float* coher, int n, int W, int H, const int4* roi,
                          const int* raw_count, const uint32_t* raw, int raw_cap, FpbPost prm, int* out_count,
                          FpbMinutiaDev* out, double* scratch, int* idx_ws);
#define FPB_POST_SCRATCH_DOUBLES(raw_cap) (1 + (size_t)(raw_cap) * 8)      // per image
#define FPB_POST_IDX_INTS(raw_cap) (3 * (size_t)(raw_cap))                 // per image

// stand-alone nms_adaptive (mode 1) / remove_redundant_oriented_adaptive (mode 2) on one list
void fpb_minutiae_select(FpbLaunch L, int mode, int n, const double* buf, double p0, double p1, int* iws, unsigned char* keep_out);

// ---- k_synth.cu : synthetic ridge prints generated on the device (counter-based, BASELINE configs[3]) ----------------
void fpb_synth_ridge_launch(FpbLaunch L, uint8_t* dst, int n, int W, int H, unsigned long long seed, unsigned long long first_index,
                            float period, float noise_sigma, float jitter);

// ---- k_gabor.cu : EXTENSION rows G1/G2 (not in the reference; opt-in) ------------------------------
#include <vector>
#define FPB_GABOR_RMAX 31
struct FpbGaborParams {         // mirrors fpb_gabor_params of include/fpb200.h
    int n_orient, min_period, max_period;
    double sigma_factor, radius_factor, min_amplitude, default_period;
};
struct FpbGaborBank {           // device-resident filter bank
    float* d_taps; int* d_offset; int* d_radius;
    int n_orient, pmin, pmax, rmax;
};
int fpb_gabor_build_bank(const FpbGaborParams& p, std::vector<float>& taps, std::vector<int>& offset, std::vector<int>& radius);
void fpb_ridge_frequency(FpbLaunch L, const uint8_t* img, const uint8_t* mask, int n, int W, int H, const int4* roi,
                         const float* blk_theta, const FpbGaborParams& p, float* blk_freq);
void fpb_gabor_apply(FpbLaunch L, const uint8_t* img, const uint8_t* mask, int n, int W, int H, const int4* roi,
                     const float* blk_theta, const float* blk_freq, const FpbGaborBank& bank, float* response, uint8_t* enhanced);
