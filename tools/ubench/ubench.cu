// Micro-benchmarks and probes for the B200 (sm_100a) figures DESIGN.md / bench.py quote.
//
//   ubench alu        thread-instruction issue rates of the integer pipes k_nlm lives on (IADD3, LOP3, IMAD, IDP.4A,
//                     VABSDIFF4, ISETP) and of LDS.64 - the MEASURED denominator of the "alu" roofline in bench.py
//   ubench umma       correctness of tcgen05.mma kind::i8 (u8 x u8 and s8 x s8 -> s32) on the canonical no-swizzle
//                     K-major shared-memory layout k_nlm_mma builds by hand, + MMA and tcgen05.ld (TMEM read) rates
//
// Prints one JSON object per test on stdout.  Build: make -C tools/ubench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

// ------------------------------------------------------------------------------------------------ ALU issue rates
template <int OP>
__global__ void __launch_bounds__(256) k_alu(unsigned* out, unsigned a0, unsigned b0, int iters) {
    unsigned x0 = a0 + threadIdx.x, x1 = b0 ^ threadIdx.x, x2 = a0 * 3 + 1, x3 = b0 + 7, x4 = a0 ^ 0x55, x5 = b0 + 11, x6 = a0 + 13, x7 = b0 * 5;
    const unsigned c = b0 | 1u, zero = b0 >> 31;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#define R8(STMT) STMT(x0, x1) STMT(x1, x2) STMT(x2, x3) STMT(x3, x4) STMT(x4, x5) STMT(x5, x6) STMT(x6, x7) STMT(x7, x0)
#define S_IADD(a, b) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
#define S_LOP3(a, b) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
#define S_IMAD(a, b) asm volatile("mad.lo.u32 %0, %0, %2, %1;" : "+r"(a) : "r"(b), "r"(c));
#define S_DP4A(a, b) asm volatile("dp4a.u32.u32 %0, %0, %2, %1;" : "+r"(a) : "r"(b), "r"(c));
#define S_VABS(a, b) asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(zero));
#define S_MIN(a, b)  asm volatile("min.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
#define S_SETP(a, b) asm volatile("{ .reg .pred p; setp.ge.s32 p, %0, %1; @p add.u32 %0, %0, 1; }" : "+r"(a) : "r"(b));
            if (OP == 0) { R8(S_IADD) }
            if (OP == 1) { R8(S_LOP3) }
            if (OP == 2) { R8(S_IMAD) }
            if (OP == 3) { R8(S_DP4A) }
            if (OP == 4) { R8(S_VABS) }
            if (OP == 5) { R8(S_MIN) }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}

__global__ void __launch_bounds__(256) k_lds64(unsigned* out, int iters) {
    __shared__ uint2 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) buf[i] = make_uint2(i, i * 3);
    __syncthreads();
    unsigned acc = 0; int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) { const uint2 v = buf[(idx + u * 37) & 1023]; acc += v.x ^ v.y; }
        idx += 5;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

static float time_launch(void (*launch)(void*), void* ctx, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(ctx); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(ctx); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

struct AluCtx { int op; unsigned* out; int grid, iters; };
static void launch_alu(void* p) {
    AluCtx* c = (AluCtx*)p;
    switch (c->op) {
    case 0: k_alu<0><<<c->grid, 256>>>(c->out, 3, 5, c->iters); break;
    case 1: k_alu<1><<<c->grid, 256>>>(c->out, 3, 5, c->iters); break;
    case 2: k_alu<2><<<c->grid, 256>>>(c->out, 3, 5, c->iters); break;
    case 3: k_alu<3><<<c->grid, 256>>>(c->out, 3, 5, c->iters); break;
    case 4: k_alu<4><<<c->grid, 256>>>(c->out, 3, 5, c->iters); break;
    case 5: k_alu<5><<<c->grid, 256>>>(c->out, 3, 5, c->iters); break;
    default: k_lds64<<<c->grid, 256>>>(c->out, c->iters); break;
    }
}

static void run_alu() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int grid = pr.multiProcessorCount * 8, iters = 2000;
    unsigned* out; CK(cudaMalloc(&out, (size_t)grid * 256 * 4));
    const char* names[7] = {"iadd3", "lop3", "imad", "idp4a", "vabsdiff4", "imnmx", "lds64"};
    const int per_iter[7] = {128, 128, 128, 128, 128, 128, 16};      // thread-instructions per loop iteration (SASS checked)
    printf("{\"test\": \"alu\", \"sms\": %d, \"clock_mhz\": %d", pr.multiProcessorCount, pr.clockRate / 1000);
    for (int op = 0; op < 7; ++op) {
        AluCtx c{op, out, grid, iters};
        const float ms = time_launch(launch_alu, &c);
        const double tinstr = (double)grid * 256 * iters * per_iter[op];
        printf(", \"%s_tinstr_per_s\": %.4g", names[op], tinstr / (ms * 1e-3));
    }
    printf("}\n");
    CK(cudaFree(out));
}

// ------------------------------------------------------------------------------------------------ tcgen05 probe
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: element (row, k) of an operand lives at
//   (row % 8) * 16 + (row / 8) * SBO + (k / 16) * LBO + (k % 16)        [bytes]
__host__ __device__ inline uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version 1 (sm_100)
    return d;                                     // layout type 0 = no swizzle, base offset 0
}
// kind::i8 instruction descriptor: c_format S32 (2) at [4,6), a/b format at [7,10)/[10,13) (0 = u8, 1 = s8), K-major both,
// n_dim = N >> 3 at [17,23), m_dim = M >> 4 at [24,29)
__host__ __device__ inline uint32_t umma_idesc_i8(int M, int N, int a_signed, int b_signed) {
    return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(mbar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
    uint32_t done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
        if (spin > (1u << 24)) __trap();
    }
}
#define TMEM_LD32(taddr, v) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), \
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), \
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), \
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr) : "memory")
#define TMEM_WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

// One CTA: D[128][N] (s32) = A[128][K] * B[N][K]^T, operands given row-major in global memory, re-laid into the
// canonical layout by the threads (exactly what k_nlm_mma does with image patches).
__global__ void __launch_bounds__(128) k_umma_probe(const uint8_t* __restrict__ A, const uint8_t* __restrict__ B, int N, int K,
                                                    int a_signed, int b_signed, int* __restrict__ D) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int kc = K / 16;                                   // 16-byte K chunks per row
    const uint32_t LBO = 128, SBO = 128 * kc;
    uint8_t* sA = sm; uint8_t* sB = sm + 128 * K;
    for (int i = threadIdx.x; i < 128 * kc; i += 128) {
        const int row = i / kc, c = i - row * kc;
        *reinterpret_cast<uint4*>(sA + (row & 7) * 16 + (row >> 3) * SBO + c * LBO) = *reinterpret_cast<const uint4*>(A + (size_t)row * K + c * 16);
    }
    for (int i = threadIdx.x; i < N * kc; i += 128) {
        const int row = i / kc, c = i - row * kc;
        *reinterpret_cast<uint4*>(sB + (row & 7) * 16 + (row >> 3) * SBO + c * LBO) = *reinterpret_cast<const uint4*>(B + (size_t)row * K + c * 16);
    }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) { mbar_init(&mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy operand stores -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_i8(128, N, a_signed, b_signed);
        for (int ks = 0; ks < K / 32; ++ks)
            umma_i8(tm, umma_desc(smem_u32(sA) + ks * 2 * LBO, LBO, SBO), umma_desc(smem_u32(sB) + ks * 2 * LBO, LBO, SBO), idesc, ks > 0);
        umma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        TMEM_LD32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
        TMEM_WAIT_LD();
#pragma unroll
        for (int j = 0; j < 32; ++j) if (c0 + j < N) D[(size_t)threadIdx.x * N + c0 + j] = (int)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(256));
}

// Rates: every CTA issues `iters` x (128 x 256 x 64) i8 MMAs from resident operands; separately 4 / 8 warps read TMEM back.
__global__ void __launch_bounds__(256) k_umma_rate(int iters, int mode, unsigned* sink) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < (128 + 256) * 64 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i * 2654435761u;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) { mbar_init(&mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    unsigned acc = 0;
    if (mode == 0) {                       // MMA issue rate
        if (threadIdx.x == 0) {
            const uint32_t idesc = umma_idesc_i8(128, 256, 0, 0);
            const uint64_t ad = umma_desc(smem_u32(sm), 128, 512), bd = umma_desc(smem_u32(sm) + 128 * 64, 128, 512);
            for (int i = 0; i < iters; ++i) { umma_i8(tm, ad, bd, idesc, 1); umma_i8(tm, ad + 16, bd + 16, idesc, 1); }   // +16 = +256 B = k-step 1
            umma_commit(&mbar);
        }
        mbar_wait(&mbar, 0);
    } else {                               // TMEM read rate: every warp streams its lane quadrant's 256 columns
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int c0 = 0; c0 < 256; c0 += 32) {
                uint32_t v[32];
                TMEM_LD32(tm + ((uint32_t)((warp & 3) * 32) << 16) + c0, v);
                TMEM_WAIT_LD();
                acc += v[0] ^ v[13] ^ v[31];
            }
        }
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(256));
}

static int probe_case(int N, int K, int sa, int sb) {
    std::vector<uint8_t> A(128 * K), B((size_t)N * K);
    srand(1234 + N + K + sa * 7 + sb * 13);
    for (auto& v : A) v = (uint8_t)(rand() & 255);
    for (auto& v : B) v = (uint8_t)(rand() & 255);
    uint8_t *dA, *dB; int* dD;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dD, (size_t)128 * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, (size_t)128 * N * 4));
    const int smem = (128 + N) * K;
    CK(cudaFuncSetAttribute(k_umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_umma_probe<<<1, 128, smem>>>(dA, dB, N, K, sa, sb, dD);
    CK(cudaDeviceSynchronize());
    std::vector<int> D((size_t)128 * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            long long s = 0;
            for (int k = 0; k < K; ++k) {
                const int a = sa ? (int)(int8_t)A[m * K + k] : (int)A[m * K + k];
                const int b = sb ? (int)(int8_t)B[(size_t)n * K + k] : (int)B[(size_t)n * K + k];
                s += a * b;
            }
            if ((int)s != D[(size_t)m * N + n]) ++bad;
        }
    printf("{\"test\": \"umma_i8\", \"M\": 128, \"N\": %d, \"K\": %d, \"a_signed\": %d, \"b_signed\": %d, \"mismatches\": %lld}\n", N, K, sa, sb, bad);
    CK(cudaFree(dA)); CK(cudaFree(dB)); CK(cudaFree(dD));
    return bad != 0;
}

struct RateCtx { int iters, mode, threads, grid; unsigned* sink; };
static void launch_rate(void* p) {
    RateCtx* c = (RateCtx*)p;
    k_umma_rate<<<c->grid, c->threads, (128 + 256) * 64>>>(c->iters, c->mode, c->sink);
}

static int run_umma() {
    int fail = 0;
    fail |= probe_case(256, 64, 0, 0);
    fail |= probe_case(256, 64, 1, 1);
    fail |= probe_case(128, 96, 1, 1);
    fail |= probe_case(64, 32, 0, 1);
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    unsigned* sink; CK(cudaMalloc(&sink, (size_t)pr.multiProcessorCount * 2 * 256 * 4));
    CK(cudaFuncSetAttribute(k_umma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 + 256) * 64));
    {
        RateCtx c{2000, 0, 128, pr.multiProcessorCount, sink};
        const float ms = time_launch(launch_rate, &c);
        const double macs = (double)c.grid * c.iters * 2.0 * 128 * 256 * 32;
        printf("{\"test\": \"umma_i8_rate\", \"ctas\": %d, \"tops\": %.4g, \"macs_per_clk_per_sm\": %.1f, \"ms\": %.3f}\n", c.grid, 2 * macs / (ms * 1e-3) / 1e12,
               macs / c.grid / (ms * 1e-3 * pr.clockRate * 1e3), ms);
    }
    for (int threads = 128; threads <= 256; threads += 128) {
        RateCtx c{2000, 1, threads, pr.multiProcessorCount, sink};
        const float ms = time_launch(launch_rate, &c);
        const double bytes = (double)c.grid * c.iters * (threads / 32) * 8.0 * 32 * 32 * 4;
        printf("{\"test\": \"tmem_ld_rate\", \"ctas\": %d, \"warps\": %d, \"bytes_per_clk_per_sm\": %.1f, \"ms\": %.3f}\n", c.grid, threads / 32,
               bytes / c.grid / (ms * 1e-3 * pr.clockRate * 1e3), ms);
    }
    CK(cudaFree(sink));
    return fail;
}

int main(int argc, char** argv) {
    const char* what = argc > 1 ? argv[1] : "all";
    int fail = 0;
    if (!strcmp(what, "alu") || !strcmp(what, "all")) run_alu();
    if (!strcmp(what, "umma") || !strcmp(what, "all")) fail |= run_umma();
    return fail;
}
