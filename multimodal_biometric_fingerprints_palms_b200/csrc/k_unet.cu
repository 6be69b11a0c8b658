// U-Net++ segmenter inference (include/fpb200_unet.h): the reference's NestedUNet
// (/root/reference/src/preprocessing/segmentation/model.py:26-83) in eval mode, as inference.py:87-133 runs it.
//
// Every 3x3 convolution (+ folded BatchNorm + ReLU) is ONE implicit GEMM on the tensor cores:
//     M = 128 output pixels per CTA, N = 64 / 128 output channels, K = 9 taps x input channels (16 per k-block)
//   * activations are zero-bordered NHWC planes, so the K slice of a (pixel, tap) is 64 contiguous bytes at a shifted
//     address - no boundary tests, no im2col buffer; a block's input may be the concatenation of up to four tensors
//     (torch.cat along channels, model.py:74-81): the k loop simply walks the sources
//   * operands go global -> shared with 16-byte cp.async into the canonical no-swizzle K-major UMMA layout (8-row x 16-byte
//     core matrices: the geometry k_nlm_mma uses, checked against a CPU GEMM by tools/ubench), three stages
//   * tcgen05.mma.cta_group::1.kind::tf32, accumulators in tensor memory; 3xTF32: every tensor is stored as hi (upper 19
//     bits) + lo (exact remainder) and D += A_lo B_hi + A_hi B_lo + A_hi B_hi, which restores fp32 accuracy (the dropped
//     term is 2^-22 relative), so the logits agree with the fp32 PyTorch module to round-off
//   * epilogue: tcgen05.ld -> scale / shift (conv bias + eval BatchNorm) -> ReLU -> hi / lo planes of the next tensor
// Max-pooling, bilinear x2 up-sampling (align_corners=True) and the final 1x1 convolution are CUDA-core kernels.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <new>
#include <vector>

#include "../../include/fpb200.h"
#include "../../include/fpb200_unet.h"
#include "fpb_kernels.h"                 // FPB_OPT_IN_SMEM (per-device opt-in to large dynamic shared memory)

#define UN_BM 128
#define UN_BK 16                      // floats per k-block = 64 bytes per operand row
#define UN_STAGES 3
#define UN_MAX_SRC 4
#define UN_THREADS 128

struct UnSrc { const float* hi; const float* lo; int C; };          // [N][H+2][W+2][C], zero border
struct UnConvArgs {
    UnSrc src[UN_MAX_SRC]; int nsrc;
    int N, H, W;
    const float* w_hi; const float* w_lo;     // [OC][Ktot], k order = (source, tap, channel)
    const float* scale; const float* shift;   // [OC]
    float* out_hi; float* out_lo; int OC;     // [N][H+2][W+2][OC]
    int Ktot;
};

__device__ __forceinline__ uint32_t un_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// K-major, no swizzle: start >> 4 | LBO (128 B between the two core matrices of a k-step) | SBO (512 B between 8-row groups)
__device__ __forceinline__ uint64_t un_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46);
}
// kind::tf32: D = f32 (1 << 4), A = B = tf32 (2), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
#define UN_IDESC(N) ((1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)(N) >> 3) << 17) | ((128u >> 4) << 24))

__device__ __forceinline__ void un_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void un_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void un_bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (unsigned spin = 0; ; ++spin) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if (spin > (1u << 24)) __trap();         // a lost arrival must fail loudly, never hang the GPU
    }
}
__device__ __forceinline__ void un_cp16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
#define UN_LD16(taddr, v) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
                 : "r"(taddr) : "memory")

__device__ __forceinline__ void un_split(float v, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);       // the 19 bits a tf32 operand keeps
    lo = v - hi;                                                  // exact
}

template <int BN>
__global__ void __launch_bounds__(UN_THREADS)
k_unet_conv3x3(const UnConvArgs a) {
    constexpr int A_BYTES = UN_BM * UN_BK * 4, B_BYTES = BN * UN_BK * 4, STAGE = 2 * A_BYTES + 2 * B_BYTES;
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + UN_STAGES * STAGE);          // empty[UN_STAGES], done
    uint32_t* tmem_sh = reinterpret_cast<uint32_t*>(sm + UN_STAGES * STAGE + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar_empty = un_smem(&bars[0]), bar_done = un_smem(&bars[UN_STAGES]);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(un_smem(tmem_sh)), "r"(BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        if (tid == 0) {
            for (int i = 0; i <= UN_STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_empty + 8 * i));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_sh;

    // ---- this thread's operand rows: A row = output pixel tid of the tile, B row = output channel tid of the tile
    const long long P = (long long)a.N * a.H * a.W;
    const long long p = (long long)blockIdx.x * UN_BM + tid;
    const bool pvalid = p < P;
    int n = 0, y = 0, x = 0;
    if (pvalid) { n = (int)(p / ((long long)a.H * a.W)); const int rem = (int)(p - (long long)n * a.H * a.W); y = rem / a.W; x = rem - y * a.W; }
    const size_t pix00 = ((size_t)n * (a.H + 2) + y) * (a.W + 2) + x;             // padded index of tap (0,0) = input pixel (y-1, x-1)
    const int oc_row = blockIdx.y * BN + tid;
    const bool bvalid = tid < BN;
    const uint32_t a_dst = (uint32_t)((tid >> 3) * 512 + (tid & 7) * 16);          // + chunk * 128
    const int KB = a.Ktot / UN_BK;

    // k-block iterator (identical in every thread): source s, tap t, channel block cb
    int ld_s = 0, ld_t = 0, ld_cb = 0;
    auto load_stage = [&](int kb, int st) {
        uint8_t* base = sm + st * STAGE;
        const UnSrc S = a.src[ld_s];
        const int ky = ld_t / 3, kx = ld_t - 3 * ky;
        const size_t off = (pix00 + (size_t)ky * (a.W + 2) + kx) * S.C + (size_t)ld_cb * UN_BK;
        const uint32_t nbytes = pvalid ? 16u : 0u;                                  // rows past the last pixel: zero fill
        const float* gh = pvalid ? S.hi + off : S.hi;
        const float* gl = pvalid ? S.lo + off : S.lo;
        const uint32_t da = un_smem(base) + a_dst;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            un_cp16(da + c * 128, gh + 4 * c, nbytes);
            un_cp16(da + A_BYTES + c * 128, gl + 4 * c, nbytes);
        }
        if (bvalid) {
            const size_t woff = (size_t)oc_row * a.Ktot + (size_t)kb * UN_BK;
            const uint32_t db = un_smem(base) + 2 * A_BYTES + a_dst;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                un_cp16(db + c * 128, a.w_hi + woff + 4 * c, 16u);
                un_cp16(db + B_BYTES + c * 128, a.w_lo + woff + 4 * c, 16u);
            }
        }
        if (++ld_cb == S.C / UN_BK) { ld_cb = 0; if (++ld_t == 9) { ld_t = 0; ++ld_s; } }
    };

    for (int i = 0; i < UN_STAGES - 1; ++i) {
        if (i < KB) load_stage(i, i);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int kb = 0; kb < KB; ++kb) {
        const int st = kb % UN_STAGES;
        asm volatile("cp.async.wait_group %0;" :: "n"(UN_STAGES - 2) : "memory");   // this thread's copies of stage kb have landed
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                // ... and are visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sb = un_smem(sm + st * STAGE);
            const uint64_t a_hi = un_desc(sb), a_lo = un_desc(sb + A_BYTES), b_hi = un_desc(sb + 2 * A_BYTES),
                           b_lo = un_desc(sb + 2 * A_BYTES + B_BYTES);
#pragma unroll
            for (int j = 0; j < UN_BK / 8; ++j) {                                   // one k-step = 8 floats = two core matrices = +256 bytes
                un_mma(tmem, a_lo + 16 * j, b_hi + 16 * j, UN_IDESC(BN), (kb | j) ? 1u : 0u);
                un_mma(tmem, a_hi + 16 * j, b_lo + 16 * j, UN_IDESC(BN), 1u);
                un_mma(tmem, a_hi + 16 * j, b_hi + 16 * j, UN_IDESC(BN), 1u);
            }
            un_commit(bar_empty + 8 * st);                                          // stage st may be refilled when these complete
            if (kb == KB - 1) un_commit(bar_done);
        }
        // refill the stage k-block kb-1 used (its MMAs were committed one iteration ago; the ones just issued keep the pipe busy)
        const int nxt = kb + UN_STAGES - 1;
        if (nxt < KB) {
            if (kb >= 1) un_bar_wait(bar_empty + 8 * ((kb - 1) % UN_STAGES), (uint32_t)(((kb - 1) / UN_STAGES) & 1));
            load_stage(nxt, nxt % UN_STAGES);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    un_bar_wait(bar_done, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: lane = pixel row, 16 output channels per tcgen05.ld
    const size_t out_pix = (((size_t)n * (a.H + 2) + y + 1) * (a.W + 2) + x + 1) * a.OC + (size_t)blockIdx.y * BN;
#pragma unroll 1
    for (int c16 = 0; c16 < BN / 16; ++c16) {
        uint32_t v[16];
        UN_LD16(tmem + ((uint32_t)(warp * 32) << 16) + c16 * 16, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (pvalid) {
            float hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int oc = blockIdx.y * BN + c16 * 16 + i;
                float r = __uint_as_float(v[i]) * __ldg(a.scale + oc) + __ldg(a.shift + oc);
                r = fmaxf(r, 0.0f);
                un_split(r, hi[i], lo[i]);
            }
            float4* oh = reinterpret_cast<float4*>(a.out_hi + out_pix + c16 * 16);
            float4* ol = reinterpret_cast<float4*>(a.out_lo + out_pix + c16 * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                oh[i] = make_float4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                ol[i] = make_float4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(BN));
}

// ---- CUDA-core kernels on the zero-bordered NHWC hi / lo planes (value = hi + lo, exact) ------------------------------
// input [n][IC][H][W] float (NCHW, as the torch module takes it) -> 16-channel NHWC planes (channels >= IC stay zero)
__global__ void k_unet_input(const float* __restrict__ in, int N, int IC, int H, int W, float* __restrict__ hi, float* __restrict__ lo) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, P = (long long)N * H * W;
    if (i >= P) return;
    const int n = (int)(i / ((long long)H * W)), rem = (int)(i - (long long)n * H * W), y = rem / W, x = rem - y * W;
    const size_t o = (((size_t)n * (H + 2) + y + 1) * (W + 2) + x + 1) * 16;
    for (int c = 0; c < IC; ++c) {
        float h_, l_;
        un_split(in[(((size_t)n * IC + c) * H + y) * W + x], h_, l_);
        hi[o + c] = h_; lo[o + c] = l_;
    }
}

// nn.MaxPool2d(2): [N][H][W][C] -> [N][H/2][W/2][C]; four channels per thread
__global__ void k_unet_pool(const float* __restrict__ ih, const float* __restrict__ il, int N, int H, int W, int C,
                            float* __restrict__ oh, float* __restrict__ ol) {
    const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, T = (long long)N * Ho * Wo * C4;
    if (i >= T) return;
    const int c4 = (int)(i % C4); long long r = i / C4;
    const int x = (int)(r % Wo); r /= Wo; const int y = (int)(r % Ho), n = (int)(r / Ho);
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const size_t o = (((size_t)n * (H + 2) + 2 * y + dy + 1) * (W + 2) + 2 * x + dx + 1) * C + 4 * c4;
            const float4 h_ = *reinterpret_cast<const float4*>(ih + o), l_ = *reinterpret_cast<const float4*>(il + o);
            m[0] = fmaxf(m[0], h_.x + l_.x); m[1] = fmaxf(m[1], h_.y + l_.y); m[2] = fmaxf(m[2], h_.z + l_.z); m[3] = fmaxf(m[3], h_.w + l_.w);
        }
    float h4[4], l4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) un_split(m[k], h4[k], l4[k]);
    const size_t o = (((size_t)n * (Ho + 2) + y + 1) * (Wo + 2) + x + 1) * C + 4 * c4;
    *reinterpret_cast<float4*>(oh + o) = make_float4(h4[0], h4[1], h4[2], h4[3]);
    *reinterpret_cast<float4*>(ol + o) = make_float4(l4[0], l4[1], l4[2], l4[3]);
}

// nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True): [N][H][W][C] -> [N][2H][2W][C]
// PyTorch: src = dst * (in - 1) / (out - 1) in float, i0 = int(src), l1 = src - i0, l0 = 1 - l1,
//          out = l0y * (l0x * a + l1x * b) + l1y * (l0x * c + l1x * d)
__global__ void k_unet_up(const float* __restrict__ ih, const float* __restrict__ il, int N, int H, int W, int C,
                          float* __restrict__ oh, float* __restrict__ ol) {
    const int Ho = 2 * H, Wo = 2 * W, C4 = C / 4;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, T = (long long)N * Ho * Wo * C4;
    if (i >= T) return;
    const int c4 = (int)(i % C4); long long r = i / C4;
    const int x = (int)(r % Wo); r /= Wo; const int y = (int)(r % Ho), n = (int)(r / Ho);
    const float sy = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.0f, sx = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.0f;
    const float fy = sy * (float)y, fx = sx * (float)x;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly1 = fy - (float)y0, ly0 = 1.0f - ly1, lx1 = fx - (float)x0, lx0 = 1.0f - lx1;
    auto at = [&](int yy, int xx) {
        const size_t o = (((size_t)n * (H + 2) + yy + 1) * (W + 2) + xx + 1) * C + 4 * c4;
        const float4 h_ = *reinterpret_cast<const float4*>(ih + o), l_ = *reinterpret_cast<const float4*>(il + o);
        return make_float4(h_.x + l_.x, h_.y + l_.y, h_.z + l_.z, h_.w + l_.w);
    };
    const float4 A = at(y0, x0), B = at(y0, x1), Cc = at(y1, x0), D = at(y1, x1);
    const float v[4] = {ly0 * (lx0 * A.x + lx1 * B.x) + ly1 * (lx0 * Cc.x + lx1 * D.x), ly0 * (lx0 * A.y + lx1 * B.y) + ly1 * (lx0 * Cc.y + lx1 * D.y),
                        ly0 * (lx0 * A.z + lx1 * B.z) + ly1 * (lx0 * Cc.z + lx1 * D.z), ly0 * (lx0 * A.w + lx1 * B.w) + ly1 * (lx0 * Cc.w + lx1 * D.w)};
    float h4[4], l4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) un_split(v[k], h4[k], l4[k]);
    const size_t o = (((size_t)n * (Ho + 2) + y + 1) * (Wo + 2) + x + 1) * C + 4 * c4;
    *reinterpret_cast<float4*>(oh + o) = make_float4(h4[0], h4[1], h4[2], h4[3]);
    *reinterpret_cast<float4*>(ol + o) = make_float4(l4[0], l4[1], l4[2], l4[3]);
}

// final nn.Conv2d(64, 1, kernel_size=1): logits[n][y][x] = bias + sum_c x[c] * w[c]
__global__ void k_unet_final(const float* __restrict__ ih, const float* __restrict__ il, int N, int H, int W, int C,
                             const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, P = (long long)N * H * W;
    if (i >= P) return;
    const int n = (int)(i / ((long long)H * W)), rem = (int)(i - (long long)n * H * W), y = rem / W, x = rem - y * W;
    const size_t o = (((size_t)n * (H + 2) + y + 1) * (W + 2) + x + 1) * C;
    float acc = 0.0f;
    for (int c = 0; c < C; c += 4) {
        const float4 h_ = *reinterpret_cast<const float4*>(ih + o + c), l_ = *reinterpret_cast<const float4*>(il + o + c);
        const float4 w_ = *reinterpret_cast<const float4*>(w + c);
        acc += (h_.x + l_.x) * w_.x; acc += (h_.y + l_.y) * w_.y; acc += (h_.z + l_.z) * w_.z; acc += (h_.w + l_.w) * w_.w;
    }
    out[i] = acc + b[0];
}

// ------------------------------------------------------------------------------------------------ host side
static char g_unet_create_error[512] = "";

struct UnTensor { float* hi = nullptr; float* lo = nullptr; int C = 0, level = 0; };
struct UnConv { int nsrc = 0, srcC[UN_MAX_SRC] = {0, 0, 0, 0}, IC = 0, ICreal = 0, OC = 0, Ktot = 0;
                float *w_hi = nullptr, *w_lo = nullptr, *scale = nullptr, *shift = nullptr; bool set = false; };

struct fpb_unet {
    int device, maxB, H, W, IC;
    cudaStream_t st;
    char err[512];
    UnConv conv[FPB_UNET_NUM_BLOCKS][2];
    float *final_w, *final_b; bool final_set;
    // tensors (hi / lo planes for max_batch images)
    UnTensor in16, mid[4], x00, x01, x02, x03, x10, x11, x12, x20, x21, x30, p0, p1, p2, u0, u1, u2;
    float *d_input, *d_logits;
    std::vector<void*> allocs;
    int launches, tc_launches;
};

static int ufail(fpb_unet* u, int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(u ? u->err : g_unet_create_error, 512, fmt, ap);
    va_end(ap);
    return code;
}
#define UCU(u, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return ufail(u, FPB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

extern "C" const char* fpb_unet_last_error(const fpb_unet* u) { return u ? u->err : g_unet_create_error; }

extern "C" void fpb_unet_destroy(fpb_unet* u) {
    if (!u) return;
    cudaSetDevice(u->device);
    if (u->st) { cudaStreamSynchronize(u->st); cudaStreamDestroy(u->st); }
    for (void* p : u->allocs) cudaFree(p);
    delete u;
}

static const int kUnFilters[5] = {64, 128, 256, 512, 1024};

static void un_block_sources(int block, int ic16, int* nsrc, int* srcC, int* oc, int* level) {
    const int* f = kUnFilters;
    auto set = [&](int lv, int o, std::initializer_list<int> cs) { *level = lv; *oc = o; *nsrc = 0; for (int c : cs) srcC[(*nsrc)++] = c; };
    switch (block) {
        case FPB_UNET_CONV0_0: set(0, f[0], {ic16}); break;
        case FPB_UNET_CONV1_0: set(1, f[1], {f[0]}); break;
        case FPB_UNET_CONV2_0: set(2, f[2], {f[1]}); break;
        case FPB_UNET_CONV3_0: set(3, f[3], {f[2]}); break;
        case FPB_UNET_CONV4_0: set(4, f[4], {f[3]}); break;
        case FPB_UNET_UP1_0: set(0, f[0], {f[0], f[1]}); break;
        case FPB_UNET_UP2_0: set(1, f[1], {f[1], f[2]}); break;
        case FPB_UNET_UP3_0: set(2, f[2], {f[2], f[3]}); break;
        case FPB_UNET_UP1_1: set(0, f[0], {f[0], f[0], f[1]}); break;
        case FPB_UNET_UP2_1: set(1, f[1], {f[1], f[1], f[2]}); break;
        default: set(0, f[0], {f[0], f[0], f[0], f[1]}); break;      // FPB_UNET_UP1_2
    }
}

extern "C" int fpb_unet_create(fpb_unet** out, int device, int max_batch, int height, int width, int input_channels) {
    if (!out) return ufail(nullptr, FPB_E_ARG, "fpb_unet_create: out is NULL");
    *out = nullptr;
    if (max_batch < 1 || height < 16 || width < 16 || height > 1024 || width > 1024 || (height % 16) || (width % 16))
        return ufail(nullptr, FPB_E_SHAPE, "fpb_unet_create: height and width must be multiples of 16 in [16, 1024] (four 2x2 poolings)");
    if (input_channels < 1 || input_channels > 16) return ufail(nullptr, FPB_E_ARG, "fpb_unet_create: input_channels must be 1..16");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return ufail(nullptr, FPB_E_CUDA, "fpb_unet_create: no CUDA device (%s) - this library has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return ufail(nullptr, FPB_E_ARG, "fpb_unet_create: device %d out of range", device);
    fpb_unet* u = new (std::nothrow) fpb_unet();
    if (!u) return ufail(nullptr, FPB_E_NOMEM, "out of host memory");
    u->device = device; u->maxB = max_batch; u->H = height; u->W = width; u->IC = input_channels; u->st = nullptr; u->err[0] = 0;
    u->final_w = u->final_b = nullptr; u->final_set = false; u->launches = u->tc_launches = 0;
#define UCC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
        ufail(nullptr, FPB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); fpb_unet_destroy(u); return FPB_E_CUDA; } } while (0)
    UCC(cudaSetDevice(device));
    UCC(cudaStreamCreateWithFlags(&u->st, cudaStreamNonBlocking));
    auto dalloc = [&](void** p, size_t bytes) -> cudaError_t {
        cudaError_t r = cudaMalloc(p, bytes);
        if (r == cudaSuccess) { u->allocs.push_back(*p); r = cudaMemsetAsync(*p, 0, bytes, u->st); }
        return r;
    };
    auto talloc = [&](UnTensor& t, int level, int C) -> cudaError_t {
        t.C = C; t.level = level;
        const size_t el = (size_t)max_batch * ((height >> level) + 2) * ((width >> level) + 2) * C;
        cudaError_t r = dalloc((void**)&t.hi, el * sizeof(float));
        if (r == cudaSuccess) r = dalloc((void**)&t.lo, el * sizeof(float));
        return r;
    };
    const int* f = kUnFilters;
    UCC(talloc(u->in16, 0, 16));
    for (int l = 0; l < 4; ++l) UCC(talloc(u->mid[l], l, f[l]));
    UCC(talloc(u->x00, 0, f[0])); UCC(talloc(u->x01, 0, f[0])); UCC(talloc(u->x02, 0, f[0])); UCC(talloc(u->x03, 0, f[0]));
    UCC(talloc(u->x10, 1, f[1])); UCC(talloc(u->x11, 1, f[1])); UCC(talloc(u->x12, 1, f[1]));
    UCC(talloc(u->x20, 2, f[2])); UCC(talloc(u->x21, 2, f[2])); UCC(talloc(u->x30, 3, f[3]));
    UCC(talloc(u->p0, 1, f[0])); UCC(talloc(u->p1, 2, f[1])); UCC(talloc(u->p2, 3, f[2]));
    UCC(talloc(u->u0, 0, f[1])); UCC(talloc(u->u1, 1, f[2])); UCC(talloc(u->u2, 2, f[3]));
    UCC(dalloc((void**)&u->d_input, (size_t)max_batch * input_channels * height * width * sizeof(float)));
    UCC(dalloc((void**)&u->d_logits, (size_t)max_batch * height * width * sizeof(float)));
    UCC(dalloc((void**)&u->final_w, 64 * sizeof(float)));
    UCC(dalloc((void**)&u->final_b, sizeof(float)));
    for (int b = 0; b < FPB_UNET_NUM_BLOCKS; ++b) {
        int nsrc, srcC[UN_MAX_SRC], oc, level;
        un_block_sources(b, 16, &nsrc, srcC, &oc, &level);
        UnConv& c0 = u->conv[b][0]; UnConv& c1 = u->conv[b][1];
        c0.nsrc = nsrc; c0.IC = 0;
        for (int s = 0; s < nsrc; ++s) { c0.srcC[s] = srcC[s]; c0.IC += srcC[s]; }
        c0.ICreal = b == FPB_UNET_CONV0_0 ? input_channels : c0.IC;
        c0.OC = oc; c0.Ktot = 9 * c0.IC;
        c1.nsrc = 1; c1.srcC[0] = oc; c1.IC = c1.ICreal = oc; c1.OC = oc; c1.Ktot = 9 * oc;
        if (b == FPB_UNET_CONV4_0) continue;                 // loaded for state_dict compatibility, never evaluated (model.py:69)
        for (UnConv* c : {&c0, &c1}) {
            UCC(dalloc((void**)&c->w_hi, (size_t)c->OC * c->Ktot * sizeof(float)));
            UCC(dalloc((void**)&c->w_lo, (size_t)c->OC * c->Ktot * sizeof(float)));
            UCC(dalloc((void**)&c->scale, c->OC * sizeof(float)));
            UCC(dalloc((void**)&c->shift, c->OC * sizeof(float)));
        }
    }
    UCC(cudaStreamSynchronize(u->st));
#undef UCC
    *out = u;
    return FPB_OK;
}

extern "C" int fpb_unet_conv_shape(const fpb_unet* u, int block, int conv, int* in_channels, int* out_channels) {
    if (!u || block < 0 || block >= FPB_UNET_NUM_BLOCKS || conv < 0 || conv > 1) return FPB_E_ARG;
    if (in_channels) *in_channels = u->conv[block][conv].ICreal;
    if (out_channels) *out_channels = u->conv[block][conv].OC;
    return FPB_OK;
}

extern "C" int fpb_unet_set_conv(fpb_unet* u, int block, int conv, const float* weight, const float* bias, const float* bn_weight,
                                 const float* bn_bias, const float* bn_mean, const float* bn_var, double bn_eps) {
    if (!u || block < 0 || block >= FPB_UNET_NUM_BLOCKS || conv < 0 || conv > 1) return FPB_E_ARG;
    if (!weight || !bias || !bn_weight || !bn_bias || !bn_mean || !bn_var) return ufail(u, FPB_E_ARG, "null parameter array");
    UnConv& c = u->conv[block][conv];
    c.set = true;
    if (block == FPB_UNET_CONV4_0) return FPB_OK;
    UCU(u, cudaSetDevice(u->device));
    // [OC][ICreal][3][3] -> [OC][Ktot] with k = 9 * (channel offset of source s) + tap * C_s + c, split into hi / lo
    std::vector<float> wh((size_t)c.OC * c.Ktot, 0.0f), wl((size_t)c.OC * c.Ktot, 0.0f), sc(c.OC), sh(c.OC);
    for (int oc = 0; oc < c.OC; ++oc) {
        int coff = 0;
        for (int s = 0; s < c.nsrc; ++s) {
            for (int t = 0; t < 9; ++t)
                for (int ch = 0; ch < c.srcC[s]; ++ch) {
                    const int ic = coff + ch;
                    if (ic >= c.ICreal) continue;                                    // zero-padded input channels of the first layer
                    const float v = weight[((size_t)oc * c.ICreal + ic) * 9 + t];
                    uint32_t bits; memcpy(&bits, &v, 4); bits &= 0xFFFFE000u;
                    float hi; memcpy(&hi, &bits, 4);
                    const size_t k = (size_t)oc * c.Ktot + 9 * (size_t)coff + (size_t)t * c.srcC[s] + ch;
                    wh[k] = hi; wl[k] = v - hi;
                }
            coff += c.srcC[s];
        }
        const float s_ = bn_weight[oc] / sqrtf(bn_var[oc] + (float)bn_eps);
        sc[oc] = s_; sh[oc] = (bias[oc] - bn_mean[oc]) * s_ + bn_bias[oc];
    }
    UCU(u, cudaMemcpyAsync(c.w_hi, wh.data(), wh.size() * sizeof(float), cudaMemcpyHostToDevice, u->st));
    UCU(u, cudaMemcpyAsync(c.w_lo, wl.data(), wl.size() * sizeof(float), cudaMemcpyHostToDevice, u->st));
    UCU(u, cudaMemcpyAsync(c.scale, sc.data(), sc.size() * sizeof(float), cudaMemcpyHostToDevice, u->st));
    UCU(u, cudaMemcpyAsync(c.shift, sh.data(), sh.size() * sizeof(float), cudaMemcpyHostToDevice, u->st));
    UCU(u, cudaStreamSynchronize(u->st));
    return FPB_OK;
}

extern "C" int fpb_unet_set_final(fpb_unet* u, const float* weight, const float* bias) {
    if (!u || !weight || !bias) return FPB_E_ARG;
    UCU(u, cudaSetDevice(u->device));
    UCU(u, cudaMemcpyAsync(u->final_w, weight, 64 * sizeof(float), cudaMemcpyHostToDevice, u->st));
    UCU(u, cudaMemcpyAsync(u->final_b, bias, sizeof(float), cudaMemcpyHostToDevice, u->st));
    UCU(u, cudaStreamSynchronize(u->st));
    u->final_set = true;
    return FPB_OK;
}

static void un_conv(fpb_unet* u, const UnConv& c, std::initializer_list<const UnTensor*> srcs, const UnTensor& out, int n) {
    UnConvArgs a; memset(&a, 0, sizeof(a));
    int s = 0;
    for (const UnTensor* t : srcs) { a.src[s].hi = t->hi; a.src[s].lo = t->lo; a.src[s].C = t->C; ++s; }
    a.nsrc = s; a.N = n; a.H = u->H >> out.level; a.W = u->W >> out.level;
    a.w_hi = c.w_hi; a.w_lo = c.w_lo; a.scale = c.scale; a.shift = c.shift;
    a.out_hi = out.hi; a.out_lo = out.lo; a.OC = c.OC; a.Ktot = c.Ktot;
    const long long P = (long long)n * a.H * a.W;
    const unsigned gx = (unsigned)((P + UN_BM - 1) / UN_BM);
    if (c.OC >= 128) {
        const size_t smem = UN_STAGES * (2 * UN_BM * UN_BK * 4 + 2 * 128 * UN_BK * 4) + 128;
        FPB_OPT_IN_SMEM(k_unet_conv3x3<128>, smem);
        k_unet_conv3x3<128><<<dim3(gx, c.OC / 128), UN_THREADS, smem, u->st>>>(a);
    } else {
        const size_t smem = UN_STAGES * (2 * UN_BM * UN_BK * 4 + 2 * 64 * UN_BK * 4) + 128;
        FPB_OPT_IN_SMEM(k_unet_conv3x3<64>, smem);
        k_unet_conv3x3<64><<<dim3(gx, c.OC / 64), UN_THREADS, smem, u->st>>>(a);
    }
    ++u->launches; ++u->tc_launches;
}

static void un_block(fpb_unet* u, int block, std::initializer_list<const UnTensor*> srcs, const UnTensor& out, int n) {
    const UnTensor& mid = u->mid[out.level];
    un_conv(u, u->conv[block][0], srcs, mid, n);
    un_conv(u, u->conv[block][1], {&mid}, out, n);
}

static void un_pool(fpb_unet* u, const UnTensor& in, const UnTensor& out, int n) {
    const int H = u->H >> in.level, W = u->W >> in.level;
    const long long T = (long long)n * (H / 2) * (W / 2) * (in.C / 4);
    k_unet_pool<<<(unsigned)((T + 255) / 256), 256, 0, u->st>>>(in.hi, in.lo, n, H, W, in.C, out.hi, out.lo);
    ++u->launches;
}

static void un_up(fpb_unet* u, const UnTensor& in, const UnTensor& out, int n) {
    const int H = u->H >> in.level, W = u->W >> in.level;
    const long long T = (long long)n * (2 * H) * (2 * W) * (in.C / 4);
    k_unet_up<<<(unsigned)((T + 255) / 256), 256, 0, u->st>>>(in.hi, in.lo, n, H, W, in.C, out.hi, out.lo);
    ++u->launches;
}

extern "C" int fpb_unet_forward(fpb_unet* u, const float* input, int n, float* logits) {
    if (!u || !input || !logits) return FPB_E_ARG;
    if (n < 1 || n > u->maxB) return ufail(u, FPB_E_ARG, "batch %d outside [1, %d]", n, u->maxB);
    for (int b = 0; b < FPB_UNET_NUM_BLOCKS; ++b)
        for (int c = 0; c < 2; ++c)
            if (!u->conv[b][c].set && b != FPB_UNET_CONV4_0) return ufail(u, FPB_E_STATE, "parameters of block %d conv %d were never set", b, c);
    if (!u->final_set) return ufail(u, FPB_E_STATE, "parameters of the final convolution were never set");
    UCU(u, cudaSetDevice(u->device));
    u->launches = u->tc_launches = 0;
    const size_t in_el = (size_t)n * u->IC * u->H * u->W;
    UCU(u, cudaMemcpyAsync(u->d_input, input, in_el * sizeof(float), cudaMemcpyHostToDevice, u->st));
    const long long P = (long long)n * u->H * u->W;
    k_unet_input<<<(unsigned)((P + 255) / 256), 256, 0, u->st>>>(u->d_input, n, u->IC, u->H, u->W, u->in16.hi, u->in16.lo);
    ++u->launches;
    // encoder (model.py:65-69; x4_0 is computed by the reference but never used)
    un_block(u, FPB_UNET_CONV0_0, {&u->in16}, u->x00, n);
    un_pool(u, u->x00, u->p0, n); un_block(u, FPB_UNET_CONV1_0, {&u->p0}, u->x10, n);
    un_pool(u, u->x10, u->p1, n); un_block(u, FPB_UNET_CONV2_0, {&u->p1}, u->x20, n);
    un_pool(u, u->x20, u->p2, n); un_block(u, FPB_UNET_CONV3_0, {&u->p2}, u->x30, n);
    // nested decoder (model.py:72-81)
    un_up(u, u->x10, u->u0, n); un_block(u, FPB_UNET_UP1_0, {&u->x00, &u->u0}, u->x01, n);
    un_up(u, u->x20, u->u1, n); un_block(u, FPB_UNET_UP2_0, {&u->x10, &u->u1}, u->x11, n);
    un_up(u, u->x30, u->u2, n); un_block(u, FPB_UNET_UP3_0, {&u->x20, &u->u2}, u->x21, n);
    un_up(u, u->x11, u->u0, n); un_block(u, FPB_UNET_UP1_1, {&u->x00, &u->x01, &u->u0}, u->x02, n);
    un_up(u, u->x21, u->u1, n); un_block(u, FPB_UNET_UP2_1, {&u->x10, &u->x11, &u->u1}, u->x12, n);
    un_up(u, u->x12, u->u0, n); un_block(u, FPB_UNET_UP1_2, {&u->x00, &u->x01, &u->x02, &u->u0}, u->x03, n);
    k_unet_final<<<(unsigned)((P + 255) / 256), 256, 0, u->st>>>(u->x03.hi, u->x03.lo, n, u->H, u->W, 64, u->final_w, u->final_b, u->d_logits);
    ++u->launches;
    UCU(u, cudaGetLastError());
    UCU(u, cudaMemcpyAsync(logits, u->d_logits, (size_t)P * sizeof(float), cudaMemcpyDeviceToHost, u->st));
    UCU(u, cudaStreamSynchronize(u->st));
    UCU(u, cudaGetLastError());
    return FPB_OK;
}

extern "C" int fpb_unet_launches(const fpb_unet* u, int* total, int* tensor_core) {
    if (!u) return FPB_E_ARG;
    if (total) *total = u->launches;
    if (tensor_core) *tensor_core = u->tc_launches;
    return FPB_OK;
}
