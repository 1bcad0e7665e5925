// Single-CTA tcgen05 self-test: D[128][N] = A . B^T for K-major or MN-major shared-memory operands (fp16, fp32
// accumulate).  Validates the descriptor encodings the decoder kernels rely on (run by tests/test_gpu_kernels.py).
//   a_major / b_major = 0: source is [rows = M or N][K] row-major, staged as K-major SWIZZLE_128B slabs
//                         (slab s = K columns 64s..64s+63, rows = M/N index)
//                     = 1: source is [K][M or N] row-major, staged as MN-major slabs
//                         (slab s = MN columns 64s..64s+63, rows = K index; LBO = slab stride, SBO = 1024 B)
#include "hps_common.cuh"
#include <cuda_fp16.h>

namespace hps {
namespace tcst {

constexpr int SLAB = 16384;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t slab_off(int row, int chunk) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
    return pred != 0;
}
// generic descriptor: start, LBO, SBO in bytes, SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(128, 1)
selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K,
                int a_major, int b_major) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // A region: 4 slabs (64 KB), B region: 4 slabs (64 KB)
    uint8_t* sA = smem;
    uint8_t* sB = smem + 4 * SLAB;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    constexpr int M = 128;
    for (int i = threadIdx.x; i < 8 * SLAB / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    // stage A
    for (int idx = threadIdx.x; idx < M * K; idx += blockDim.x) {
        int m, k;
        if (a_major == 0) { m = idx / K; k = idx % K; } else { k = idx / M; m = idx % M; }
        const __half h = __float2half_rn(A[idx]);
        uint32_t off;
        if (a_major == 0) off = (k / 64) * SLAB + slab_off(m, (k % 64) >> 3) + (k & 7) * 2;
        else              off = (m / 64) * SLAB + slab_off(k, (m % 64) >> 3) + (m & 7) * 2;
        *reinterpret_cast<__half*>(sA + off) = h;
    }
    for (int idx = threadIdx.x; idx < N * K; idx += blockDim.x) {
        int n, k;
        if (b_major == 0) { n = idx / K; k = idx % K; } else { k = idx / N; n = idx % N; }
        const __half h = __float2half_rn(B[idx]);
        uint32_t off;
        if (b_major == 0) off = (k / 64) * SLAB + slab_off(n, (k % 64) >> 3) + (k & 7) * 2;
        else              off = (n / 64) * SLAB + slab_off(k, (n % 64) >> 3) + (n & 7) * 2;
        *reinterpret_cast<__half*>(sB + off) = h;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16) |
                                   ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
            for (int ks = 0; ks < K / 16; ++ks) {
                uint64_t ad, bd;
                if (a_major == 0) ad = make_desc(smem_u32(sA) + (ks / 4) * SLAB + (ks % 4) * 32, 16, 1024);
                else              ad = make_desc(smem_u32(sA) + ks * 2048, SLAB, 1024);
                if (b_major == 0) bd = make_desc(smem_u32(sB) + (ks / 4) * SLAB + (ks % 4) * 32, 16, 1024);
                else              bd = make_desc(smem_u32(sB) + ks * 2048, SLAB, 1024);
                const uint32_t acc = ks > 0 ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd),
                    "r"(idesc), "r"(acc)
                    : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
    }
    {
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        } while (!done);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + (threadIdx.x & 31);
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
}

}  // namespace tcst
}  // namespace hps

extern "C" int hps_tc_selftest(const float* A, const float* B, float* D, int N, int K, int a_major, int b_major,
                               void* stream) {
    using namespace hps;
    HPS_REQUIRE(A && B && D, "null buffer");
    HPS_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 256, "need N, K multiples of 16 <= 256");
    const int smem = 8 * tcst::SLAB + 1024;
    static bool configured = false;
    if (!configured) {
        HPS_CUDA_OK(cudaFuncSetAttribute(tcst::selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    tcst::selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K, a_major, b_major);
    HPS_LAUNCH_OK();
    return 0;
}
