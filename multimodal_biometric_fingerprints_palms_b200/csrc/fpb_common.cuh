// Shared definitions of the fpb200 CUDA library (sm_100a).
//
// Data layout in HBM:
//   * every per-image plane of a batch is [n][H][W] with row stride W (the INPUT width), element
//     type u8 / f32 / i32;  after segmentation the per-image crop lives at the plane's origin and
//     its size is read from `roi[b] = (x0, y0, w', h')` in device memory - no host round trip.
//   * kernels take `const int4* roi` (+ a flag): when null the image is the full H x W.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define FPB_HD __host__ __device__ __forceinline__

struct FpbDims { int w, h; };

__device__ __forceinline__ FpbDims fpb_dims(const int4* roi, int b, int W, int H) {
    FpbDims d;
    if (roi) { int4 r = roi[b]; d.w = r.z; d.h = r.w; }
    else { d.w = W; d.h = H; }
    return d;
}

// OpenCV BORDER_REFLECT_101 (gfedcb|abcdefgh|gfedcba), valid for any distance
__device__ __forceinline__ int fpb_reflect101(int i, int n) {
    if ((unsigned)i < (unsigned)n) return i;                 // interior: no arithmetic
    if (n == 1) return 0;
    { const int j = i < 0 ? -i : 2 * (n - 1) - i; if ((unsigned)j < (unsigned)n) return j; }   // one reflection
    const int p = 2 * (n - 1);
    i %= p; if (i < 0) i += p;
    return i >= n ? p - i : i;
}
// scipy.ndimage mode='reflect' (dcba|abcd|dcba), valid for any distance
__device__ __forceinline__ int fpb_reflect_dup(int i, int n) {
    if ((unsigned)i < (unsigned)n) return i;
    { const int j = i < 0 ? -i - 1 : 2 * n - 1 - i; if ((unsigned)j < (unsigned)n) return j; }
    const int p = 2 * n;
    i %= p; if (i < 0) i += p;
    return i >= n ? p - 1 - i : i;
}

// Raw crossing-number minutiae per image: the reference has no cap (extract_features.py:41-69).  The per-handle capacity is
// sized from the image (H*W/8, at least FPB_RAW_CAP_MIN); a list that still does not fit is an ERROR (FPB_E_OVERFLOW),
// never a silent truncation.
#define FPB_RAW_CAP_MIN 2048
#define FPB_RAW_CAP_MAX (1 << 17)
static inline int fpb_raw_cap_for(int H, int W) {
    long long c = (long long)H * W / 8;
    if (c < FPB_RAW_CAP_MIN) c = FPB_RAW_CAP_MIN;
    if (c > FPB_RAW_CAP_MAX) c = FPB_RAW_CAP_MAX;
    return (int)c;
}
#define FPB_MAX_REFINED 128     // refined minutiae kept per image (reference keeps 60)

// packed raw minutia: x | y << 14 | type << 28
__device__ __forceinline__ uint32_t fpb_pack_raw(int x, int y, int t) { return (uint32_t)x | ((uint32_t)y << 14) | ((uint32_t)t << 28); }

struct FpbMinutiaDev {      // mirrors fpb_minutia of include/fpb200.h
    int x, y, type, pad;
    double orientation, quality, coherence, angular_stability;
};

struct FpbPost {            // mirrors fpb_post_params
    int quality_window; double quality_threshold, coherence_threshold, min_distance;
    int margin, max_minutiae, patch_radius;
};
