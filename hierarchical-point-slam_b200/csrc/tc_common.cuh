// Shared definitions of the tensor-core decoder kernels (mlp_tc.cu forward, mlp_tc_bwd.cu backward): geometry of
// the shared-memory slabs, the per-tile global workspace exchanged between forward, dgrad and wgrad, and the
// PTX wrappers (mbarrier, cp.async.bulk, tcgen05).
#pragma once
#include "hps_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace hps {
namespace tc {

constexpr int TILE = 128;            // sample points per tile = UMMA M = TMEM lanes
constexpr int SLAB = 16384;          // bytes: 128 rows x 128 B
constexpr int NST = 6;               // weight ring stages
constexpr int EPI_WARP0 = 2;         // warp 0 = producer, warp 1 = MMA issuer, then the epilogue warps
constexpr int EPI_WARPS = 16;        // 4 TMEM lane quadrants x CG column groups
constexpr int CG = EPI_WARPS / 4;    // column groups: an epilogue thread owns one point (row) x 1/CG of the columns
constexpr int NEPI = EPI_WARPS * 32;
constexpr int DUMP_WARP = EPI_WARP0 + EPI_WARPS;   // last warp: bulk-stores the operand slabs to the backward workspace
constexpr int NTHR = (DUMP_WARP + 1) * 32;
constexpr float kTwoPi = 6.283185307179586f;

template <int KIND> struct Cfg;
template <> struct Cfg<HPS_MLP_GEO> {
    static constexpr int H = 32, NF = 93, E = 93, NOUT = 1;
    static constexpr bool CONCAT = false, RELU = true;
};
template <> struct Cfg<HPS_MLP_COL> {
    static constexpr int H = 128, NF = 20, E = 40, NOUT = 3;
    static constexpr bool CONCAT = true, RELU = false;
};
template <int KIND> struct Geo {
    using C = Cfg<KIND>;
    static constexpr int HS = (C::H + 63) / 64;                 // hidden slabs
    static constexpr int ES = (C::E + 63) / 64;                 // embedding slabs
    static constexpr int AS = HS + ES;                          // A slabs per precision
    static constexpr int HCOLS = C::H < 64 ? C::H : 64;         // columns used per hidden slab
    __host__ __device__ static constexpr int e_cols(int s) { return (C::E - 64 * s) < 64 ? (C::E - 64 * s) : 64; }
    __host__ __device__ static constexpr int ksteps(int cols) { return (cols + 15) / 16; }
    // slabs consumed per tile: per layer 1 (Wc) + 2 per k-slab, + 1 output slab
    static constexpr int SLABS_PER_TILE = 5 + 2 * (ES + HS + HS + (HS + ES) + HS) + 1;
    // TMEM columns: a [0,H), u [H,2H), out [2H, 2H+16)
    static constexpr int TMEM_COLS = (2 * C::H + 16) <= 128 ? 128 : ((2 * C::H + 16) <= 256 ? 256 : 512);
    // slabs + biases [5][2][H] + Fourier matrix [3][NF] + output bias / exposure affine + barriers
    static constexpr int SMEM_BYTES = 1024 /*align slack*/ + (2 * AS + 1 + NST) * SLAB + 2 * 5 * C::H * 4 +
                                      (3 * C::NF + 4 + 12 + 2) * 4 + 256;
};

__host__ __device__ constexpr int iW(int i) { return 1 + 4 * i; }
__host__ __device__ constexpr int iB(int i) { return 2 + 4 * i; }
__host__ __device__ constexpr int iWc(int i) { return 3 + 4 * i; }
__host__ __device__ constexpr int iBc(int i) { return 4 + 4 * i; }
constexpr int iWo = 21, iBo = 22;

struct Ptrs {
    const float* v[HPS_MLP_NPARAM];
};

// ---------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// The same with an L2 eviction policy: evict_first for data that is streamed once (the weight-gradient kernels read
// 236 MB of operand slabs per iteration - more than the L2 - and must not push out what kernels running beside them
// keep resident: the neighbour search's hash table and cell-sorted points, the feature tables' hot rows).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
// shared -> global bulk copy (TMA store path) and its group bookkeeping
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// One lane of a converged warp (cute::elect_one_sync).  ptxas recognises elect.sync and issues the following
// tcgen05 / bulk-copy instructions from a single thread with uniform operands; issuing them under a plain
// `if (lane == 0)` makes it emit per-thread ELECT/R2UR replay loops (and, with conditional MMAs, miscompiled
// predicated descriptor moves - observed on CUDA 12.9).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tc_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100 version 1):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused when swizzled) | SBO>>4 [32,46) (=1024 B between 8-row groups)
// | version [46,48) = 1 | layout [61,64) = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A/B 16-bit formats (0 = f16, 1 = bf16), K-major
__host__ __device__ constexpr uint32_t instr_desc(int afmt, int bfmt, int M, int N) {
    return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
constexpr int FMT_F16 = 0, FMT_BF16 = 1;

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a slab
__device__ __forceinline__ uint32_t slab_off(int row, int chunk) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

// x = hi + lo, both fp16 (22 significand bits together).  tcgen05 kind::f16 requires A and B to have the same
// 16-bit format (a bf16 x fp16 product traps as an illegal instruction on sm_100a), hence fp16 for both halves.
__device__ __forceinline__ void split16(float x, uint16_t& hi, uint16_t& lo) {
    const __half h = __float2half_rn(x);
    const __half l = __float2half_rn(x - __half2float(h));
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
}

// MUFU wrappers without the denormal pre/post-scaling of __expf/__logf (inputs here are never denormal where it
// matters: lg2 sees 1 + e in [1,2], a flushed ex2 result is a correct 0)
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// sin / cos of Fourier arguments (|x| up to ~1e4 rad): exact-product reduction by 2*pi in two FMAs (error of the
// reduced argument ~1e-7), then the MUFU approximations on |r| <= pi (absolute error 2^-20.9).  The fp32 argument
// itself carries ~6e-5 rad of rounding at |x| ~ 1e3, in the reference as well, so nothing is lost against sinf/cosf,
// at a quarter of the instructions.
__device__ __forceinline__ float reduce_2pi(float x) {
    const float k = rintf(x * 0.15915494309189535f);
    return fmaf(-k, -1.7484555314695172e-07f, fmaf(-k, 6.2831854820251465f, x));
}
__device__ __forceinline__ float sin_2pi(float x) { return __sinf(reduce_2pi(x)); }
__device__ __forceinline__ float cos_2pi(float x) { return __cosf(reduce_2pi(x)); }

// Eight values at once: two packed-conversion instructions per pair instead of four scalar conversions; the
// halves are bit-identical to split16's.  vh / vl are one 16-byte slab chunk each.
__device__ __forceinline__ void split8(const float (&x)[8], uint4& vh, uint4& vl) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __half2 hh = __floats2half2_rn(x[2 * q], x[2 * q + 1]);
        const float2 back = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(x[2 * q] - back.x, x[2 * q + 1] - back.y);
        h[q] = *reinterpret_cast<const uint32_t*>(&hh);
        l[q] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    vh = make_uint4(h[0], h[1], h[2], h[3]);
    vl = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---------------------------------------------------------------------------------------------- workspace
// Per 128-point tile, written by the forward when a tensor-core backward will follow:
//   a   : float [5][H][128]           pre-activations a_i, feature-major (coalesced for one-thread-per-point access)
//   raw : float [128][4]              decoder output before the exposure tail
//   X   : NX 16-KB slab images (fp16, K-major SWIZZLE_128B, rows = points):
//           e hi [ES], e lo [ES], c (cols 0..31 hi | 32..63 lo), then for i = 0..4: y_i hi [HS], y_i lo [HS]
// written by the dgrad kernel, read by the wgrad kernel:
//   G   : NG slab images: for i = 0..4: da_i hi [HS], da_i lo [HS]; then d raw hi, d raw lo
template <int KIND> struct Ws {
    using G_ = Geo<KIND>;
    static constexpr int HS = G_::HS, ES = G_::ES, H = Cfg<KIND>::H;
    static constexpr int NX = 2 * ES + 1 + 10 * HS;
    static constexpr int NG = 10 * HS + 2;      // da_i hi/lo per layer + d raw (hi, lo): columns 0..NOUT-1 of two slabs
    static constexpr size_t A_BYTES = (size_t)5 * H * TILE * 4;
    static constexpr size_t RAW_BYTES = (size_t)TILE * 4 * 4;
    static constexpr size_t X_BYTES = (size_t)NX * SLAB;
    static constexpr size_t G_BYTES = (size_t)NG * SLAB;
    static constexpr size_t TILE_BYTES = A_BYTES + RAW_BYTES + X_BYTES + G_BYTES;
    __host__ __device__ static constexpr int x_e(int lo, int s) { return lo * ES + s; }
    __host__ __device__ static constexpr int x_c() { return 2 * ES; }
    __host__ __device__ static constexpr int x_y(int i, int lo, int s) { return 2 * ES + 1 + i * 2 * HS + lo * HS + s; }
    __host__ __device__ static constexpr int g_da(int i, int lo, int s) { return i * 2 * HS + lo * HS + s; }
    __host__ __device__ static constexpr int g_dr(int lo) { return 10 * HS + lo; }
    __device__ static float* a_ptr(uint8_t* ws, int64_t tile) { return reinterpret_cast<float*>(ws + tile * TILE_BYTES); }
    __device__ static float* raw_ptr(uint8_t* ws, int64_t tile) { return reinterpret_cast<float*>(ws + tile * TILE_BYTES + A_BYTES); }
    __device__ static uint8_t* x_ptr(uint8_t* ws, int64_t tile) { return ws + tile * TILE_BYTES + A_BYTES + RAW_BYTES; }
    __device__ static uint8_t* g_ptr(uint8_t* ws, int64_t tile) { return ws + tile * TILE_BYTES + A_BYTES + RAW_BYTES + X_BYTES; }
};

// Sum over the 32 lanes of each of 8 per-lane values (9 shuffles).  Every lane returns the total of column
// colsum8_col(lane).
__device__ __forceinline__ int colsum8_col(int lane) { return ((lane >> 2) & 1) * 4 + ((lane >> 1) & 1) * 2 + (lane & 1); }
__device__ __forceinline__ float warp_colsum8(const float (&v)[8], int lane) {
    float t4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool up = lane & 4;
        const float keep = up ? v[i + 4] : v[i], send = up ? v[i] : v[i + 4];
        t4[i] = keep + __shfl_xor_sync(kFull, send, 4);
    }
    float t2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const bool up = lane & 2;
        const float keep = up ? t4[i + 2] : t4[i], send = up ? t4[i] : t4[i + 2];
        t2[i] = keep + __shfl_xor_sync(kFull, send, 2);
    }
    const bool up = lane & 1;
    float r = (up ? t2[1] : t2[0]) + __shfl_xor_sync(kFull, up ? t2[0] : t2[1], 1);
    r += __shfl_xor_sync(kFull, r, 8);
    r += __shfl_xor_sync(kFull, r, 16);
    return r;
}

// lane j returns sum over the 32 lanes of v[j]  (31 shuffles: recursive halving)
__device__ __forceinline__ float warp_colsum32(const float (&v)[32], int lane) {
    float t16[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const bool up = lane & 16;
        const float keep = up ? v[i + 16] : v[i], send = up ? v[i] : v[i + 16];
        t16[i] = keep + __shfl_xor_sync(kFull, send, 16);
    }
    float t8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool up = lane & 8;
        const float keep = up ? t16[i + 8] : t16[i], send = up ? t16[i] : t16[i + 8];
        t8[i] = keep + __shfl_xor_sync(kFull, send, 8);
    }
    float t4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool up = lane & 4;
        const float keep = up ? t8[i + 4] : t8[i], send = up ? t8[i] : t8[i + 4];
        t4[i] = keep + __shfl_xor_sync(kFull, send, 4);
    }
    float t2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const bool up = lane & 2;
        const float keep = up ? t4[i + 2] : t4[i], send = up ? t4[i] : t4[i + 2];
        t2[i] = keep + __shfl_xor_sync(kFull, send, 2);
    }
    const bool up = lane & 1;
    const float keep = up ? t2[1] : t2[0], send = up ? t2[0] : t2[1];
    return keep + __shfl_xor_sync(kFull, send, 1);
}

// Diagnostic timeline (hps_tc_trace): CTA 0 stamps %globaltimer (ns) at phase boundaries into a caller buffer.
__device__ __forceinline__ void trace_stamp(unsigned long long* tr, int slot) {
    if (tr != nullptr) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        tr[slot] = t;
    }
}
int set_trace_fwd(void* buf);
int set_trace_bwd(void* buf);

}  // namespace tc
}  // namespace hps
