// K6: smooth_fingerprint_skeleton  (/root/reference/src/preprocessing/fingerprint_preprocess.py:141-159)
//   img = bin/255; (nx,ny) = sobel(img)/(|grad|+1e-6); 3x: smoothed += 1.4*(dx*ny - dy*nx) with fresh
//   sobels of `smoothed`; gaussian_filter(0.6); *1.25, clip; > 0.35.
//
// scipy.ndimage.sobel(axis=a) = correlate1d([-1,0,1]) along a, then correlate1d([1,2,1]) along the other
// axis, float64 accumulation, FLOAT32 intermediate, mode='reflect' - reproduced literally, so this
// stage is bit-exact (float32 ops evaluated in numpy's order, nvcc -fmad=false).
#include "fpb_kernels.h"


struct Sobel2 { float dx, dy; };

// both scipy sobels of plane p at (x,y); T = float (plane) or uint8 {0,255} read as {0,1}
template <typename T> __device__ __forceinline__ float ldv(const T* p, size_t i);
template <> __device__ __forceinline__ float ldv<float>(const float* p, size_t i) { return p[i]; }
template <> __device__ __forceinline__ float ldv<uint8_t>(const uint8_t* p, size_t i) { return (float)p[i] / 255.0f; }

template <typename T>
__device__ __forceinline__ Sobel2 sobel_at(const T* p, int W, int w, int h, int x, int y) {
    const int xm = fpb_reflect_dup(x - 1, w), xp = fpb_reflect_dup(x + 1, w);
    const int ym = fpb_reflect_dup(y - 1, h), yp = fpb_reflect_dup(y + 1, h);
    float a[3][3];
    const int ys[3] = {ym, y, yp}, xs[3] = {xm, x, xp};
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) a[j][i] = ldv<T>(p, (size_t)ys[j] * W + xs[i]);
    Sobel2 s;
    // axis=1: derivative along x (float32 intermediate), then [1,2,1] along y
    const float d0 = (float)((double)a[0][2] - (double)a[0][0]);
    const float d1 = (float)((double)a[1][2] - (double)a[1][0]);
    const float d2 = (float)((double)a[2][2] - (double)a[2][0]);
    s.dx = (float)((double)d1 * 2.0 + ((double)d0 + (double)d2));
    // axis=0: derivative along y, then [1,2,1] along x
    const float e0 = (float)((double)a[2][0] - (double)a[0][0]);
    const float e1 = (float)((double)a[2][1] - (double)a[0][1]);
    const float e2 = (float)((double)a[2][2] - (double)a[0][2]);
    s.dy = (float)((double)e1 * 2.0 + ((double)e0 + (double)e2));
    return s;
}

__global__ void k_sm_init(const uint8_t* __restrict__ bin, int W, int H, const int4* __restrict__ roi,
                          float* __restrict__ ux, float* __restrict__ uy, float* __restrict__ acc) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x >= d.w || y >= d.h) return;
    const size_t base = (size_t)b * W * H, o = base + (size_t)y * W + x;
    const Sobel2 s = sobel_at<uint8_t>(bin + base, W, d.w, d.h, x, y);
    const float mag = sqrtf(s.dx * s.dx + s.dy * s.dy) + 1e-6f;
    ux[o] = s.dx / mag; uy[o] = s.dy / mag;
    acc[o] = (float)bin[o] / 255.0f;
}

__global__ void k_sm_step(const float* __restrict__ acc, int W, int H, const int4* __restrict__ roi,
                          const float* __restrict__ ux, const float* __restrict__ uy, float* __restrict__ out) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x >= d.w || y >= d.h) return;
    const size_t base = (size_t)b * W * H, o = base + (size_t)y * W + x;
    const Sobel2 s = sobel_at<float>(acc + base, W, d.w, d.h, x, y);
    const float proj = s.dx * uy[o] - s.dy * ux[o];
    out[o] = acc[o] + 1.4f * proj;
}

__global__ void k_sm_final(const float* __restrict__ sm, int W, int H, const int4* __restrict__ roi,
                           uint8_t* __restrict__ dst) {
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const FpbDims d = fpb_dims(roi, b, W, H);
    if (x >= d.w || y >= d.h) return;
    const size_t o = (size_t)b * W * H + (size_t)y * W + x;
    float v = sm[o] * 1.25f;
    v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    dst[o] = (v > 0.35f) ? 255 : 0;
}

void fpb_smooth_core(FpbLaunch L, const uint8_t* binary, int n, int W, int H, const int4* roi,
                     float* ux, float* uy, float* acc, float* acc2, float* tmp, uint8_t* dst) {
    const dim3 blk(32, 8), grid((W + 31) / 32, (H + 7) / 8, n);
    k_sm_init<<<grid, blk, 0, L.st>>>(binary, W, H, roi, ux, uy, acc);          LAUNCH_COUNT(L);
    k_sm_step<<<grid, blk, 0, L.st>>>(acc, W, H, roi, ux, uy, acc2);            LAUNCH_COUNT(L);
    k_sm_step<<<grid, blk, 0, L.st>>>(acc2, W, H, roi, ux, uy, acc);            LAUNCH_COUNT(L);
    k_sm_step<<<grid, blk, 0, L.st>>>(acc, W, H, roi, ux, uy, acc2);            LAUNCH_COUNT(L);
    fpb_gaussian_f32(L, acc2, n, W, H, roi, 0.6, tmp, acc);
    k_sm_final<<<grid, blk, 0, L.st>>>(acc, W, H, roi, dst);                    LAUNCH_COUNT(L);
}
