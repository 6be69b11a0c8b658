/*
 * fpb200 - C ABI of the B200-native fingerprint enhance -> minutiae hot path.
 *
 * One handle = one CUDA device + one stream + a workspace sized for `max_batch`
 * images of `height` x `width` uint8 pixels.  No shared mutable state between
 * handles; one handle per host thread / per GPU.  All functions return 0 on success
 * or a negative FPB_E_* code; `fpb_last_error` gives the message.  No exceptions, no
 * torch types, plain pointers and sizes only.
 *
 * The reference (GiovanniIacuzzo/multimodal_biometric_fingerprints_palms) is pure
 * Python and has no FFI; its "plugin interface" for this path is the set of Python
 * callables listed in SURVEY.md section 8(b).  Each entry point below names the reference
 * callable (file:line under /root/reference) whose arithmetic it replaces; the Python
 * package `multimodal_biometric_fingerprints_palms_b200` binds these with ctypes and
 * re-exports the reference's names.
 */
#ifndef FPB200_H
#define FPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPB_ABI_VERSION 2

#define FPB_OK            0
#define FPB_E_ARG        -1   /* bad argument                                  */
#define FPB_E_CUDA       -2   /* CUDA runtime error (message in last_error)    */
#define FPB_E_NOMEM      -3
#define FPB_E_STATE      -4   /* call order / handle state                     */
#define FPB_E_SHAPE      -5   /* image too small / too large for this handle   */
#define FPB_E_OVERFLOW   -6   /* more raw minutiae than fpb_raw_capacity(): the reference keeps them all, so the
                                 run is refused instead of truncated             */

typedef struct fpb_handle fpb_handle;

/* post-processing parameters: defaults reproduce the values hard-coded at
 * src/features/post_processing.py:77-83 (NOT the dead YAML values) */
typedef struct fpb_post_params {
    int    quality_window;       /* 25   ; 1 .. 33, odd or even (cv2.blur's anchor win / 2) */
    double quality_threshold;    /* 0.15 */
    double coherence_threshold;  /* 0.2  */
    double min_distance;         /* 8.0  */
    int    margin;               /* 30   */
    int    max_minutiae;         /* 60   ; 0 .. 128 (capacity of the result block), larger is FPB_E_ARG */
    int    patch_radius;         /* 15   */
} fpb_post_params;

/* one refined minutia - the fields of the reference's JSON record
 * (src/features/extract_features.py:104-105, post_processing.py:122-127) */
typedef struct fpb_minutia {
    int32_t x, y;
    int32_t type;                /* 0 = "ending", 1 = "bifurcation" */
    int32_t _pad;
    double  orientation;         /* rad in [-pi/2, pi/2) */
    double  quality;
    double  coherence;
    double  angular_stability;
} fpb_minutia;

/* identifiers for fpb_fetch_plane: intermediates of the last fpb_run_* call */
enum {
    FPB_PLANE_NORMALIZED = 0,    /* u8  [n,H,W]            normalize_image          */
    FPB_PLANE_DENOISED   = 1,    /* u8  [n,H,W]            denoise_image            */
    FPB_PLANE_SEGMENTED  = 2,    /* u8  [n,H,W] crop at origin, valid [h',w']       */
    FPB_PLANE_MASK       = 3,    /* u8  [n,H,W] crop                                 */
    FPB_PLANE_BINARY     = 4,    /* u8  [n,H,W] crop                                 */
    FPB_PLANE_SMOOTH     = 5,    /* u8  [n,H,W] crop      smooth_fingerprint_skeleton*/
    FPB_PLANE_SKELETON   = 6,    /* u8  [n,H,W] crop                                 */
    FPB_PLANE_ORIENT     = 7,    /* f32 [n,H,W] crop      orient_img                 */
    FPB_PLANE_RELIAB     = 8,    /* f32 [n,H,W] crop      rel_img                    */
    FPB_PLANE_GATE       = 9,    /* u8  [n,H,W] crop      mask entering skeletonize  */
    FPB_PLANE_NLM        = 10,   /* u8  [n,H,W]           NLM output before the blur */
    FPB_PLANE_SKEL_ORIENT = 11,  /* f32 [n,H,W] crop      K9: orientation of skeleton*/
    FPB_PLANE_SKEL_COHER  = 12,  /* f32 [n,H,W] crop      K9: coherence              */
    FPB_PLANE_DENSITY     = 13,  /* f32 [n,H,W] crop      K9: normalised density     */
    FPB_PLANE_ENHANCED    = 14,  /* u8  [n,H,W] crop      EXTENSION: Gabor-enhanced image (fpb_enable_enhanced) */
    FPB_PLANE_GABOR       = 15,  /* f32 [n,H,W] crop      EXTENSION: Gabor response                             */
    FPB_PLANE_SKELETON_FILE = 16,/* u8  [n,H,W] crop      the skeleton as cv2.imread returns it after the reference's
                                    cv2.imwrite(.jpg, quality 95) hand-off - what K8/K9 read (fpb_set_handoff)      */
    FPB_PLANE_COUNT_
};

/* EXTENSION (NOT in the reference, no oracle there): per-block ridge frequency and
 * oriented Gabor enhancement in the style of Hong, Wan & Jain (1998) on the 16x16 block grid of
 * compute_orientation_map.  Defaults in brackets.  The bank holds n_orient x (max_period-min_period+1) filters:
 * exp(-(u^2+v^2)/(2 sigma^2)) cos(2 pi x_n / period), sigma = sigma_factor * period, radius = ceil(radius_factor *
 * sigma) (capped at 31), made zero-mean and L1-normalised. */
typedef struct fpb_gabor_params {
    int    n_orient;          /* [16]   orientation bins over [0, pi)                           */
    int    min_period;        /* [3]    valid ridge periods (pixels), also the bank's range     */
    int    max_period;        /* [25]                                                           */
    double sigma_factor;      /* [0.45]                                                         */
    double radius_factor;     /* [2.5]                                                          */
    double min_amplitude;     /* [8.0]  x-signature swing (grey levels) below which a block is invalid */
    double default_period;    /* [9.0]  used when no block of an image has a valid frequency    */
} fpb_gabor_params;

/* ---- lifetime ------------------------------------------------------------------ */
int  fpb_abi_version(void);
/* `cuda_stream` may be NULL (the handle creates its own non-blocking stream) or a
 * cudaStream_t owned by the caller (e.g. torch.cuda.current_stream().cuda_stream). */
int  fpb_create(fpb_handle** out, int device, int max_batch, int height, int width, void* cuda_stream);
void fpb_destroy(fpb_handle* h);
const char* fpb_last_error(const fpb_handle* h);   /* h may be NULL: error of fpb_create */
int  fpb_sync(fpb_handle* h);                      /* cudaStreamSynchronize             */

/* 256-entry thinning table (skimage `skeletonize` convention: neighbour code
 * NW=1 N=2 NE=4 E=8 SE=16 S=32 SW=64 W=128; value 1/2/3 = deletable in sub-iteration
 * 1 / 2 / either).  Default: Zhang-Suen (1984).  fingerprint_preprocess.py:171 */
int  fpb_set_thin_table(fpb_handle* h, const uint8_t table[256]);
int  fpb_set_post_params(fpb_handle* h, const fpb_post_params* p);   /* NULL = defaults */
/* thinning_and_cleaning(..., rel_thresh) fingerprint_preprocess.py:161-170: smoothed reliability must exceed it.  Default
 * 0.1 (the value the reference's path passes, :202); config_fingerprint.yml general.rel_threshold is the opt-in override. */
int  fpb_set_rel_threshold(fpb_handle* h, double rel_thresh);
/* How the skeleton reaches extract_minutiae / postprocess_minutiae inside fpb_run_*.
 *   1 (default) = the reference's CLI flow: run_preprocessing.py:137-140 writes <base>_skeleton.jpg (JPEG, quality 95)
 *       and extract_features.py:83-92 reads it back, so K8 thresholds the DECODED grey levels at 127 and K9 computes
 *       density / orientation / coherence on the codec's ringing.  The file's pixels are reproduced on the device
 *       (forward islow DCT, quantise, dequantise, inverse DCT: bit-identical to cv2.imwrite + cv2.imread).
 *   0 = in memory: K8/K9 read the clean {0,255} skeleton (calling the two reference functions in one process). */
int  fpb_set_handoff(fpb_handle* h, int mode);
/* raw crossing-number minutiae this handle can hold per image: max(2048, H*W/8).  The reference has no cap; a longer
 * list makes the run fail with FPB_E_OVERFLOW (never truncated). */
int  fpb_raw_capacity(const fpb_handle* h);

/* EXTENSION: when enabled, fpb_run_* also computes the block frequencies and the Gabor-enhanced image of the
 * `segmented` crop (the result key "enhanced" that run_preprocessing.py:133 looks for).  p == NULL: defaults.
 * Everything else the run produces is unchanged.  fpb_disable_enhanced frees the extra planes. */
int  fpb_enable_enhanced(fpb_handle* h, const fpb_gabor_params* p);
int  fpb_disable_enhanced(fpb_handle* h);
/* stage form (host buffers): orientation field of (img, mask) as compute_orientation_map, then frequency + Gabor.
 * freq_blocks [n, H/16, W/16] (0 rows/cols beyond an image's grid), response f32 [n,H,W] (optional), enhanced u8 [n,H,W] */
int  fpb_enhance_gabor(fpb_handle* h, const uint8_t* img, const uint8_t* mask, int n, const fpb_gabor_params* p,
                       float* freq_blocks, float* response, uint8_t* enhanced);
/* block frequencies of the last run / stage call: [n, H/16, W/16] floats */
int  fpb_fetch_freq_blocks(fpb_handle* h, float* dst, size_t bytes);

/* ---- whole hot path: preprocess_fingerprint (fingerprint_preprocess.py:182-225) followed by
 *      extract_minutiae (extract_features.py:41-69) and postprocess_minutiae
 *      (post_processing.py:69-137) on the skeleton as the feature stage reads it (fpb_set_handoff) */
/* device-resident input: d_images = n*H*W bytes on this handle's device; asynchronous */
int  fpb_run_device(fpb_handle* h, const uint8_t* d_images, int n);
/* host input (pinned or pageable): H2D, run, and D2H of the per-image results into the
 * handle's pinned result block; synchronous on return */
int  fpb_run_host(fpb_handle* h, const uint8_t* images, int n);

/* asynchronous form (enqueue / wait): H2D, the run and the D2H of roi / counts / refined lists are enqueued
 * on the handle's streams and the call returns; `images` must stay valid - and should be pinned - until fpb_wait returns.
 * Two handles used alternately overlap one batch's copies and host work with the other batch's kernels. */
int  fpb_run_host_async(fpb_handle* h, const uint8_t* images, int n);
int  fpb_wait(fpb_handle* h);

/* results of the last run (host memory owned by the handle, valid until the next run;
 * after fpb_run_device call fpb_download_results first) */
int  fpb_download_results(fpb_handle* h);
/* roi = x0,y0,w',h' of the crop inside the input image */
int  fpb_result_roi(const fpb_handle* h, int image, int32_t roi[4]);
/* raw crossing-number minutiae in row-major order: returns count (>=0); up to `cap` entries
 * written as (x, y, type) triplets */
int  fpb_result_raw(const fpb_handle* h, int image, int32_t* xyt, int cap);
/* refined minutiae, quality-descending, at most max_minutiae */
int  fpb_result_minutiae(const fpb_handle* h, int image, fpb_minutia* out, int cap);
/* Streaming loops (BASELINE configs[3]): fpb_download_refined copies roi / counts / refined lists only (no raw lists:
 * 48 B x 64 + 24 B per image instead of + 4 B x fpb_raw_capacity), and fpb_result_block hands the whole batch over in one
 * call: roi4 [n][4] int32, raw_counts [n], out_counts [n], out [n][cap] (first min(count, cap) entries of each row
 * written); any of the four may be NULL.  Returns the number of images of the last run. */
int  fpb_download_refined(fpb_handle* h);
int  fpb_result_block(const fpb_handle* h, int32_t* roi4, int32_t* raw_counts, int32_t* out_counts, fpb_minutia* out, int cap);
/* copy an intermediate plane of the last run to host memory (synchronous);
 * `bytes` must be n*H*W*sizeof(element) */
int  fpb_fetch_plane(fpb_handle* h, int plane_id, void* dst, size_t bytes);
/* per-stage device times of the last fpb_run_* call (CUDA events on the handle's stream): enable with
 * fpb_set_profiling(h,1); ms[0..7] = K1,K2,K3,K4,K5,K6,K7+K8,K9 ; ms[8] = the NLM kernel alone.
 * Returns the number of entries defined (9). */
int  fpb_set_profiling(fpb_handle* h, int on);
int  fpb_stage_times(fpb_handle* h, float* ms, int cap);
/* per-launch device times of the last fpb_run_* call: text lines "<file>:<line> <ms>" in launch order (the line is
 * the LAUNCH_COUNT site right after the <<<>>>).  Pass enable=1 once to switch the recording on (buf may be NULL),
 * run, then call again with a buffer.  Returns the number of bytes written. */
int  fpb_kernel_times(fpb_handle* h, int enable, char* buf, int cap);
/* number of kernel launches issued by this handle since creation */
long long fpb_launch_count(const fpb_handle* h);

/* ---- stage entry points (host buffers, batch of n images of the handle's H x W,
 *      synchronous) - one per public function of the reference ------------------- */
/* The reference's per-file flow hands every stage after segment_fingerprint a CROP whose size depends on the image
 * (fingerprint_preprocess.py:125-129).  Instead of one handle per crop size, declare the per-image (w', h') of the next
 * stage calls: their host buffers stay n x H x W planes whose top-left h' x w' region holds image i (row stride W), and
 * only that region is read and written.  wh = [n][2] int32 (w', h'), 3 <= w' <= W, 3 <= h' <= H; NULL resets to H x W.
 * Honoured by fpb_binarize / orientation / smooth / thin / skeletonize / extract_minutiae / postprocess /
 * fpb_jpeg_roundtrip / fpb_enhance_gabor; fpb_normalize / denoise / segment and fpb_run_* take whole frames and return
 * FPB_E_STATE while crop dimensions are declared. */
int fpb_set_stage_dims(fpb_handle* h, const int32_t* wh, int n);
/* normalize_image            fingerprint_preprocess.py:13-29 */
int fpb_normalize(fpb_handle* h, const uint8_t* img, int n, uint8_t* out);
/* denoise_image              fingerprint_preprocess.py:34-38 ; nlm_out optional (may be NULL) */
int fpb_denoise(fpb_handle* h, const uint8_t* img, int n, uint8_t* out, uint8_t* nlm_out);
/* segment_fingerprint        fingerprint_preprocess.py:86-136 ; outputs are H x W planes whose
 * top-left roi[2] x roi[3] region holds the crop */
int fpb_segment(fpb_handle* h, const uint8_t* img, int n, uint8_t* segmented, uint8_t* mask, int32_t* roi4);
/* the same for colour input (fingerprint_preprocess.py:94, cv2.COLOR_BGR2GRAY first): bgr is [n,H,W,channels] uint8,
 * channels 3 or 4 (alpha ignored); the conversion (OpenCV's 15-bit fixed-point weights) runs on the device */
int fpb_segment_bgr(fpb_handle* h, const uint8_t* bgr, int channels, int n, uint8_t* segmented, uint8_t* mask, int32_t* roi4);
/* binarize                   fingerprint_preprocess.py:43-81 */
int fpb_binarize(fpb_handle* h, const uint8_t* img, int n, uint8_t* out);
/* compute_orientation_map    orientation.py:9-85 (block 16, sigmas 3.0/3.0, invert_if_needed);
 * mask may be NULL; orient_blocks is [n, H/16, W/16] */
int fpb_orientation(fpb_handle* h, const uint8_t* img, const uint8_t* mask, int n,
                    float* orient_blocks, float* orient_img, float* rel_img);
/* compute_orientation_map with the reference's keyword arguments (orientation.py:9-14) as data: block_size >= 1 with at
 * least one whole block per image (else FPB_E_SHAPE - the reference fails in cv2.resize), sigmas below 7.875
 * (FPB_E_ARG otherwise; <= 1e-15 = that Gaussian is skipped, as SciPy does), invert_if_needed 0/1.
 * orient_blocks is [n, H/block_size, W/block_size].  The defaults (16, 3.0, 1, 3.0) are fpb_orientation. */
int fpb_orientation_ex(fpb_handle* h, const uint8_t* img, const uint8_t* mask, int n, int block_size,
                       double smooth_sigma, int invert_if_needed, double smooth_orientation_sigma,
                       float* orient_blocks, float* orient_img, float* rel_img);
/* compute_orientation_map on a non-uint8 image (orientation.py:21-24): img = `img.astype(np.float32)`, [n,H,W] float32;
 * rescaled by its own minimum / maximum when it leaves [0, 1], then as fpb_orientation_ex */
int fpb_orientation_f32(fpb_handle* h, const float* img, const uint8_t* mask, int n, int block_size,
                        double smooth_sigma, int invert_if_needed, double smooth_orientation_sigma,
                        float* orient_blocks, float* orient_img, float* rel_img);
/* smooth_fingerprint_skeleton fingerprint_preprocess.py:141-159 */
int fpb_smooth(fpb_handle* h, const uint8_t* binary, int n, uint8_t* out);
/* the same with sigma / diffusion_iter / contrast_boost (fingerprint_preprocess.py:142-144) as data; the defaults
 * (1.4, 3, 1.25) give fpb_smooth's output bit for bit */
int fpb_smooth_ex(fpb_handle* h, const uint8_t* binary, int n, double sigma, int diffusion_iter,
                  double contrast_boost, uint8_t* out);
/* thinning_and_cleaning      fingerprint_preprocess.py:161-177 ; gate_out optional */
int fpb_thin(fpb_handle* h, const uint8_t* binary_smooth, const float* reliability, int n,
             uint8_t* skeleton, uint8_t* gate_out);
/* skeletonize + isolated-pixel clean-up only (fingerprint_preprocess.py:171-177) from a
 * boolean mask - the bit-exact part of the contract */
int fpb_skeletonize(fpb_handle* h, const uint8_t* gate, int n, uint8_t* skeleton);
/* extract_minutiae           extract_features.py:41-69 ; counts[n]; xyt [n, cap, 3] */
int fpb_extract_minutiae(fpb_handle* h, const uint8_t* skeleton, int n, int32_t* counts, int32_t* xyt, int cap);
/* postprocess_minutiae       post_processing.py:69-137 (gray = skel, as extract_features.py:92);
 * in: raw lists as produced by fpb_extract_minutiae; out_counts[n], out [n, cap_out] */
int fpb_postprocess(fpb_handle* h, const uint8_t* skeleton, int n, const int32_t* counts,
                    const int32_t* xyt, int cap, int32_t* out_counts, fpb_minutia* out, int cap_out);
/* the same with the `gray` argument of the reference (post_processing.py:71, 93): the orientation / coherence maps come
 * from `gray` [n,H,W] uint8, density and the intensity term from the skeleton.  gray == NULL is fpb_postprocess.
 * (The reference's gray=None is gray = (skel > 0) as 0/1 uint8 - the Python mirror passes exactly that.) */
int fpb_postprocess_gray(fpb_handle* h, const uint8_t* skeleton, const uint8_t* gray, int n, const int32_t* counts,
                         const int32_t* xyt, int cap, int32_t* out_counts, fpb_minutia* out, int cap_out);

/* nms_adaptive                post_processing.py:10-32 : keep[i] = 1 for survivors.  density[i] = density_map[y_i, x_i]
 * (float32, as indexed by the reference); quality defaults to 1.0 on the caller's side (m.get("quality", 1.0)) */
int fpb_nms_adaptive(fpb_handle* h, int n, const int32_t* xy, const double* quality, const float* density,
                     double base_dist, uint8_t* keep);
/* remove_redundant_oriented_adaptive   post_processing.py:37-64 */
int fpb_remove_redundant(fpb_handle* h, int n, const int32_t* xy, const double* quality, const double* orientation,
                         const float* density, double base_radius, double angle_thresh, uint8_t* keep);

#ifdef __cplusplus
}
#endif
#endif /* FPB200_H */
