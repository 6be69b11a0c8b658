// Synthetic ridge-pattern prints generated ON THE DEVICE from (seed, image index) with a counter-based generator
// (BASELINE.json configs[3]: "1M synthetic 240x320 images" - host I/O must not be what is timed).
// Same formula as synth.ridge_image: phase = 2 pi (r + 6 sin 2 phi) / period about a jittered core,
// I = 60 + 150 (0.5 + 0.5 cos phase) inside an ellipse (semi-axes 0.42 w, 0.46 h), background 235, + N(0, sigma), clipped.
// Randomness: Philox4x32-10 keyed by the seed; counter = (image index, pixel pair / parameter slot): any image of any
// batch can be regenerated alone, in any order, on any rank.  synth.ridge_image_counter is the NumPy twin (same
// integers, float32 formula; the transcendental functions differ in the last ulp, so a pixel may differ by one level -
// parity tests therefore run the oracle on the very images the device produced, fetched with fpb_fetch_input).
#include "fpb_kernels.h"

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                                       uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0, 1)

__global__ void __launch_bounds__(256)
k_synth_ridge(uint8_t* __restrict__ dst, int W, int H, unsigned long long seed, unsigned long long first_index, float period_fixed,
              float noise_sigma, float jitter) {
    const int b = blockIdx.z;
    const unsigned long long index = first_index + (unsigned long long)b;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32), i0 = (uint32_t)index, i1 = (uint32_t)(index >> 32);
    // per-image parameters: counter slot 0xFFFFFFFF
    uint32_t pr[4];
    philox4x32_10(i0, i1, 0xFFFFFFFFu, 0u, k0, k1, pr);
    const float period = period_fixed > 0.0f ? period_fixed : 7.0f + 4.0f * u01(pr[0]);
    const float cx = 0.5f * (float)W + jitter * (2.0f * u01(pr[1]) - 1.0f), cy = 0.5f * (float)H + jitter * (2.0f * u01(pr[2]) - 1.0f);
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 2, y = blockIdx.y * blockDim.y + threadIdx.y;   // two pixels per thread
    if (x >= W || y >= H) return;
    uint32_t rn[4];
    philox4x32_10(i0, i1, (uint32_t)(y * ((W + 1) / 2) + (x >> 1)), 1u, k0, k1, rn);
    // Box-Muller: two normals from two uniforms
    const float rad = sqrtf(-2.0f * __logf(u01(rn[0]))), ang = 6.28318530718f * u01(rn[1]);
    const float nz[2] = {rad * __cosf(ang), rad * __sinf(ang)};
    uint8_t* row = dst + (size_t)b * W * H + (size_t)y * W;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int xx = x + u;
        if (xx >= W) break;
        const float dx = (float)xx - cx, dy = (float)y - cy;
        const float r = sqrtf(dx * dx + dy * dy), phi = atan2f(dy, dx);
        float v = 60.0f + 150.0f * (0.5f + 0.5f * cosf(6.28318530718f * (r + 6.0f * sinf(2.0f * phi)) / period));
        const float ex = ((float)xx - 0.5f * (float)W) / (0.42f * (float)W), ey = ((float)y - 0.5f * (float)H) / (0.46f * (float)H);
        if (ex * ex + ey * ey > 1.0f) v = 235.0f;
        v += noise_sigma * nz[u];
        row[xx] = (uint8_t)fminf(fmaxf(v, 0.0f), 255.0f);            // clip, truncate (np.clip(...).astype(uint8))
    }
}

void fpb_synth_ridge_launch(FpbLaunch L, uint8_t* dst, int n, int W, int H, unsigned long long seed, unsigned long long first_index,
                            float period, float noise_sigma, float jitter) {
    dim3 blk(32, 8), grid(((W + 1) / 2 + 31) / 32, (H + 7) / 8, n);
    k_synth_ridge<<<grid, blk, 0, L.st>>>(dst, W, H, seed, first_index, period, noise_sigma, jitter);
    LAUNCH_COUNT(L);
}
