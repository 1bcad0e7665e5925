// Decoder MLP forward on the 5th-generation tensor cores (tcgen05 + TMEM), second edition of kernel 3.
//
// Same function as mlp_fwd_kernel in mlp.cu (reference src/conv_onet/models/decoder.py:316-341, :581-618),
// same inputs / outputs / activation workspace layout, so the two editions are interchangeable and the
// fp32 SIMT edition stays the parity anchor.
//
// Arithmetic: every fp32 operand x is split as x = hi + lo with hi = fp16(x), lo = fp16(x - hi) (22 significand
// bits) and each product is evaluated as three tensor-core MMAs  hi*hi + hi*lo + lo*hi  accumulated in fp32 in
// TMEM: relative error ~2^-22 per product (emulated end to end: 4e-7 on the decoder output, 2e-6 on gradients; a
// bf16 three-way product gives 1.4e-5 / 8e-5, too close to the 1e-4 parity bar) - see DESIGN.md section 4.
//
// Structure (one persistent CTA per SM, 128 sample points per tile, 320 threads):
//   warp 0      producer: streams pre-swizzled 16 KB weight slabs global -> shared with cp.async.bulk
//               (1-D TMA) through an mbarrier ring;
//   warp 1      one elected thread issues tcgen05.mma (M=128, N=H, K=16) for the whole CTA; accumulators
//               a = W x and u = Wc c live in TMEM; tcgen05.commit signals the epilogue;
//   warps 2..9  epilogue: tcgen05.ld the accumulators (one TMEM lane = one sample point), bias + activation
//               + feature injection in registers, re-split into bf16/fp16 and write the next layer's A operand
//               straight back into the swizzled shared-memory slabs (activations never leave the SM); the
//               Fourier features and the c tile are produced the same way at the start of a tile.
// Shared-memory operand layout = the canonical K-major SWIZZLE_128B UMMA layout: slabs of [rows][64 x 16 bit],
// 8-row / 1024-byte swizzle atoms, 16-byte chunks XORed with (row & 7).
#include "hps_common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace hps {
namespace tc {

constexpr int TILE = 128;            // sample points per tile = UMMA M = TMEM lanes
constexpr int SLAB = 16384;          // bytes: 128 rows x 128 B
constexpr int NST = 6;               // weight ring stages
constexpr int NTHR = 320;
constexpr int EPI_WARP0 = 2;
constexpr int NEPI = 256;
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
    static constexpr int SMEM_BYTES = 1024 /*align slack*/ + (2 * AS + 1 + NST) * SLAB + 2 * 5 * C::H * 4 + 256;
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

// ---------------------------------------------------------------------------------------------- pack
// Global image of the weight slabs in consumption order (see Geo::SLABS_PER_TILE):
//   for layer i: [Wc_i: cols 0..31 fp16 hi, 32..63 fp16 lo]  then for every k-slab s: [W_i hi], [W_i lo]
//   k-slab order of layer 3 = hidden slabs first, embedding slabs last (matches the A slabs in shared memory)
//   last slab: output layer, four 16-row mini-slabs [hi s0 | hi s1 | lo s0 | lo s1] (2 KB each; HS of them used)
template <int KIND>
__device__ __forceinline__ int layer_first_slab(int i) {
    using G = Geo<KIND>;
    int n = 0;
    for (int j = 0; j < i; ++j) {
        const int ks = j == 0 ? G::ES : (j == 3 ? G::HS + G::ES : G::HS);
        n += 1 + 2 * ks;
    }
    return n;
}

template <int KIND>
__global__ void pack_tc_kernel(Ptrs prm, uint16_t* __restrict__ img) {
    using C = Cfg<KIND>;
    using G = Geo<KIND>;
    constexpr int H = C::H, E = C::E;
    const int total = G::SLABS_PER_TILE * (SLAB / 2);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int slab = idx / (SLAB / 2);
        const int byte = (idx % (SLAB / 2)) * 2;
        float val = 0.f;
        bool want_lo = false;
        if (slab == G::SLABS_PER_TILE - 1) {
            // output layer mini-slabs: 2048 B each = 16 rows x 128 B
            const int mini = byte / 2048, b = byte % 2048;
            const int row = ((b >> 10) << 3) | ((b >> 7) & 7);
            const int chunk = ((b >> 4) & 7) ^ (row & 7);
            const int col = chunk * 8 + ((b >> 1) & 7);
            const int s = mini % 2;
            want_lo = mini >= 2;
            const int k = s * 64 + col;
            if (mini < 4 && s < G::HS && row < C::NOUT && k < H) val = prm.v[iWo][row * H + k];
        } else {
            int layer = 0;
            while (layer < 4 && slab >= layer_first_slab<KIND>(layer + 1)) ++layer;
            const int rel = slab - layer_first_slab<KIND>(layer);
            const int row = ((byte >> 10) << 3) | ((byte >> 7) & 7);       // output feature n
            const int chunk = ((byte >> 4) & 7) ^ (row & 7);
            const int col = chunk * 8 + ((byte >> 1) & 7);
            if (row < H) {
                if (rel == 0) {   // Wc_i
                    want_lo = col >= 32;
                    val = prm.v[iWc(layer)][row * kC + (col & 31)];
                } else {
                    const int s = (rel - 1) / 2;
                    want_lo = ((rel - 1) & 1) != 0;
                    const int K = layer == 0 ? E : (layer == 3 ? E + H : H);
                    // source column: layers 1,2,4: hidden slab s; layer 0: embedding slab s;
                    // layer 3: slabs [0,HS) hidden (reference columns E..E+H), then embedding (columns 0..E)
                    int k = -1;
                    if (layer == 0) {
                        if (s * 64 + col < E) k = s * 64 + col;
                    } else if (layer == 3) {
                        if (s < G::HS) {
                            if (s * 64 + col < H) k = E + s * 64 + col;
                        } else if ((s - G::HS) * 64 + col < E) {
                            k = (s - G::HS) * 64 + col;
                        }
                    } else if (s * 64 + col < H) {
                        k = s * 64 + col;
                    }
                    if (k >= 0) val = prm.v[iW(layer)][row * K + k];
                }
            }
        }
        uint16_t hi, lo;
        split16(val, hi, lo);
        img[idx] = want_lo ? lo : hi;
    }
}

// ---------------------------------------------------------------------------------------------- forward
template <bool RELU> __device__ __forceinline__ float act_fast(float a) {
    if (RELU) return fmaxf(a, 0.f);
    // softplus(beta=100): max(a,0) + log(1 + exp(-|100 a|)) / 100   (absolute error < 1e-9)
    return fmaxf(a, 0.f) + __logf(1.0f + __expf(-fabsf(100.0f * a))) * 0.01f;
}

template <int KIND>
__global__ void __launch_bounds__(NTHR, 1)
mlp_fwd_tc_kernel(Ptrs prm, const uint8_t* __restrict__ wimg, const float* __restrict__ p,
                  const float* __restrict__ c, const float* __restrict__ affine, int64_t P, int num_tiles,
                  float* __restrict__ out, float* __restrict__ act) {
    using C = Cfg<KIND>;
    using G = Geo<KIND>;
    constexpr int H = C::H, E = C::E, NF = C::NF, NOUT = C::NOUT, HS = G::HS, ES = G::ES, AS = G::AS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sAhi = smem;                          // AS slabs: [0,HS) hidden, [HS,AS) embedding
    uint8_t* sAlo = sAhi + AS * SLAB;
    uint8_t* sC = sAlo + AS * SLAB;                // cols 0..31 c hi, 32..63 c lo (fp16 both)
    uint8_t* sW = sC + SLAB;                       // NST ring stages
    float* sBias = reinterpret_cast<float*>(sW + NST * SLAB);      // [5][2][H]: b_i, bc_i
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 2 * 5 * H);
    uint64_t* full = bars;                         // [NST]
    uint64_t* empty = bars + NST;                  // [NST]
    uint64_t* a_ready = bars + 2 * NST;            // epilogue -> MMA: A operand in place, accumulators drained
    uint64_t* acc_ready = bars + 2 * NST + 1;      // MMA -> epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup ----
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(a_ready, NEPI);
        mbar_init(acc_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)G::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = threadIdx.x; i < 5 * H; i += NTHR) {
        const int l = i / H, n = i % H;
        sBias[(l * 2 + 0) * H + n] = prm.v[iB(l)][n];
        sBias[(l * 2 + 1) * H + n] = prm.v[iBc(l)][n];
    }
    // zero the A slabs and the c slab once (padding columns stay zero for the whole kernel)
    for (int i = threadIdx.x; i < (2 * AS + 1) * SLAB / 16; i += NTHR)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ================= producer (whole warp converged, one elected lane issues) =================
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int j = 0; j < G::SLABS_PER_TILE; ++j, ++it) {
                const int st = it % NST;
                mbar_wait(&empty[st], ((it / NST) & 1) ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&full[st], SLAB);
                    bulk_g2s(sW + st * SLAB, wimg + (size_t)j * SLAB, SLAB, &full[st]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (whole warp converged, one elected lane issues) =================
        constexpr uint32_t ID = instr_desc(FMT_F16, FMT_F16, TILE, H);
        constexpr uint32_t OD = instr_desc(FMT_F16, FMT_F16, TILE, 16);
        const uint32_t t_a = tmem, t_u = tmem + H, t_o = tmem + 2 * H;
        const uint32_t aHi = smem_u32(sAhi), aLo = smem_u32(sAlo), aC = smem_u32(sC), aW = smem_u32(sW);
        uint32_t it = 0, nready = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
#pragma unroll 1
            for (int layer = 0; layer < 6; ++layer) {
                mbar_wait(a_ready, nready & 1);
                ++nready;
                tc_fence_after();
                if (layer < 5) {
                    // u = c . Wc^T   (K = 32: two k-steps; c slab / Wc slab hold hi in columns 0..31, lo in 32..63)
                    {
                        const int st = it % NST;
                        mbar_wait(&full[st], (it / NST) & 1);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t w = aW + st * SLAB;
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                const uint64_t ah = smem_desc(aC + k * 32), al = smem_desc(aC + 64 + k * 32);
                                const uint64_t bh = smem_desc(w + k * 32), bl = smem_desc(w + 64 + k * 32);
                                tc_mma(t_u, ah, bh, ID, k > 0 ? 1u : 0u);
                                tc_mma(t_u, ah, bl, ID, 1u);
                                tc_mma(t_u, al, bh, ID, 1u);
                            }
                            tc_commit(&empty[st]);
                        }
                        __syncwarp();
                        ++it;
                    }
                    // a = x . W^T over the layer's k-slabs
                    const int nks = layer == 0 ? ES : (layer == 3 ? HS + ES : HS);
#pragma unroll 1
                    for (int s = 0; s < nks; ++s) {
                        int aslab, cols;
                        if (layer == 0) { aslab = HS + s; cols = G::e_cols(s); }
                        else if (layer == 3 && s >= HS) { aslab = s; cols = G::e_cols(s - HS); }
                        else { aslab = s; cols = G::HCOLS; }
                        const int st_hi = it % NST;
                        mbar_wait(&full[st_hi], (it / NST) & 1);
                        ++it;
                        const int st_lo = it % NST;
                        mbar_wait(&full[st_lo], (it / NST) & 1);
                        ++it;
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t wh = aW + st_hi * SLAB, wl = aW + st_lo * SLAB;
                            const uint32_t xh = aHi + aslab * SLAB, xl = aLo + aslab * SLAB;
                            const int nk = G::ksteps(cols);
#pragma unroll 1
                            for (int k = 0; k < nk; ++k) {
                                const uint64_t ah = smem_desc(xh + k * 32), al = smem_desc(xl + k * 32);
                                const uint64_t bh = smem_desc(wh + k * 32), bl = smem_desc(wl + k * 32);
                                tc_mma(t_a, ah, bh, ID, (s > 0 || k > 0) ? 1u : 0u);
                                tc_mma(t_a, ah, bl, ID, 1u);
                                tc_mma(t_a, al, bh, ID, 1u);
                            }
                            tc_commit(&empty[st_hi]);
                            tc_commit(&empty[st_lo]);
                        }
                        __syncwarp();
                    }
                } else {
                    // output layer: out[128 x 16] = y_4 . Wo^T  (mini-slabs [hi s0 | hi s1 | lo s0 | lo s1], 2 KB each)
                    const int st = it % NST;
                    mbar_wait(&full[st], (it / NST) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t w = aW + st * SLAB;
#pragma unroll 1
                        for (int s = 0; s < HS; ++s) {
                            const int nk = G::ksteps(G::HCOLS);
#pragma unroll 1
                            for (int k = 0; k < nk; ++k) {
                                const uint64_t ah = smem_desc(aHi + s * SLAB + k * 32), al = smem_desc(aLo + s * SLAB + k * 32);
                                const uint64_t bh = smem_desc(w + s * 2048 + k * 32), bl = smem_desc(w + (2 + s) * 2048 + k * 32);
                                tc_mma(t_o, ah, bh, OD, (s > 0 || k > 0) ? 1u : 0u);
                                tc_mma(t_o, ah, bl, OD, 1u);
                                tc_mma(t_o, al, bh, OD, 1u);
                            }
                        }
                        tc_commit(&empty[st]);
                    }
                    __syncwarp();
                    ++it;
                }
                if (elect_one()) tc_commit(acc_ready);
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue warps =================
        const int ew = warp - EPI_WARP0;          // 0..7
        const int quad = warp & 3;                // TMEM lane quadrant this warp may access
        const int half = ew >> 2;                 // column half handled by this warp
        const int row = quad * 32 + lane;         // tile row = TMEM lane
        constexpr int CPT = (H >= 64) ? H / 2 : H;   // hidden columns per thread: colour 64, geometry 32 (half 0 only)
        const bool col_owner = (H >= 64) || (half == 0);
        const int col0 = (H >= 64) ? half * CPT : 0;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        uint32_t nacc = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int64_t pt = (int64_t)tile * TILE + row;
            const bool live = pt < P;
            // ---- tile prologue: Fourier features -> embedding slabs, c -> c slab ----
            {
                float x0 = 0.f, x1 = 0.f, x2 = 0.f;
                if (live) {
                    x0 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 0));
                    x1 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 1));
                    x2 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 2));
                }
                const float* B = prm.v[0];
                for (int j = half; j < NF; j += 2) {
                    const float arg = fmaf(x2, __ldg(B + 2 * NF + j), fmaf(x1, __ldg(B + NF + j), x0 * __ldg(B + j)));
                    float s, co;
                    if (C::CONCAT) sincosf(arg, &s, &co); else { s = sinf(arg); co = 0.f; }
                    uint16_t hi, lo;
                    {
                        split16(s, hi, lo);
                        const int slab = HS + j / 64, cc = j % 64;
                        const uint32_t off = slab * SLAB + slab_off(row, cc >> 3) + (cc & 7) * 2;
                        *reinterpret_cast<uint16_t*>(sAhi + off) = hi;
                        *reinterpret_cast<uint16_t*>(sAlo + off) = lo;
                    }
                    if (C::CONCAT) {
                        split16(co, hi, lo);
                        const int e = NF + j;
                        const int slab = HS + e / 64, cc = e % 64;
                        const uint32_t off = slab * SLAB + slab_off(row, cc >> 3) + (cc & 7) * 2;
                        *reinterpret_cast<uint16_t*>(sAhi + off) = hi;
                        *reinterpret_cast<uint16_t*>(sAlo + off) = lo;
                    }
                }
                // c: this thread converts 16 of the 32 channels
                float4 cv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    cv[q] = live ? ldg4(c + pt * kC + half * 16 + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float* cf = reinterpret_cast<const float*>(cv);
                uint16_t hi[16], lo[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) split16(cf[q], hi[q], lo[q]);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const uint4 vh = make_uint4(hi[q * 8 + 0] | (hi[q * 8 + 1] << 16), hi[q * 8 + 2] | (hi[q * 8 + 3] << 16),
                                                hi[q * 8 + 4] | (hi[q * 8 + 5] << 16), hi[q * 8 + 6] | (hi[q * 8 + 7] << 16));
                    const uint4 vl = make_uint4(lo[q * 8 + 0] | (lo[q * 8 + 1] << 16), lo[q * 8 + 2] | (lo[q * 8 + 3] << 16),
                                                lo[q * 8 + 4] | (lo[q * 8 + 5] << 16), lo[q * 8 + 6] | (lo[q * 8 + 7] << 16));
                    *reinterpret_cast<uint4*>(sC + slab_off(row, half * 2 + q)) = vh;
                    *reinterpret_cast<uint4*>(sC + slab_off(row, 4 + half * 2 + q)) = vl;
                }
            }
            fence_proxy_async();
            mbar_arrive(a_ready);

            // activation workspace: same layout as the SIMT edition ([64-point tile][slot][feature][point]); rows
            // beyond P but inside the last 64-point tile are written too (finite values: the backward multiplies
            // them by zero gradients)
            float* act_tile0 = (act != nullptr && pt < ((P + 63) / 64) * 64)
                                   ? act + (pt / 64) * (int64_t)(10 * H * 64) + (pt % 64) : nullptr;
            for (int layer = 0; layer < 5; ++layer) {
                mbar_wait(acc_ready, nacc & 1);
                ++nacc;
                tc_fence_after();
                if (col_owner) {
#pragma unroll 1
                    for (int cb = 0; cb < CPT; cb += 32) {
                        float a[32], u[32];
                        tc_ld32(tmem + lane_addr + (col0 + cb), a);
                        tc_ld32(tmem + lane_addr + H + (col0 + cb), u);
                        const float* bi = sBias + (layer * 2 + 0) * H + col0 + cb;
                        const float* bc = sBias + (layer * 2 + 1) * H + col0 + cb;
                        uint16_t hi[32], lo[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float av = a[j] + bi[j];
                            const float yv = act_fast<C::RELU>(av) + (u[j] + bc[j]);
                            if (act_tile0 != nullptr) {
                                const int n = col0 + cb + j;
                                act_tile0[(int64_t)(2 * layer) * H * 64 + n * 64] = av;
                                act_tile0[(int64_t)(2 * layer + 1) * H * 64 + n * 64] = yv;
                            }
                            split16(yv, hi[j], lo[j]);
                        }
                        const int colabs = col0 + cb;              // first hidden column of this block
                        const int slab = colabs / 64, cc = colabs % 64;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 vh = make_uint4(hi[q * 8 + 0] | (hi[q * 8 + 1] << 16), hi[q * 8 + 2] | (hi[q * 8 + 3] << 16),
                                                        hi[q * 8 + 4] | (hi[q * 8 + 5] << 16), hi[q * 8 + 6] | (hi[q * 8 + 7] << 16));
                            const uint4 vl = make_uint4(lo[q * 8 + 0] | (lo[q * 8 + 1] << 16), lo[q * 8 + 2] | (lo[q * 8 + 3] << 16),
                                                        lo[q * 8 + 4] | (lo[q * 8 + 5] << 16), lo[q * 8 + 6] | (lo[q * 8 + 7] << 16));
                            const uint32_t off = slab * SLAB + slab_off(row, (cc >> 3) + q);
                            *reinterpret_cast<uint4*>(sAhi + off) = vh;
                            *reinterpret_cast<uint4*>(sAlo + off) = vl;
                        }
                    }
                }
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(a_ready);
            }
            // ---- output layer result ----
            mbar_wait(acc_ready, nacc & 1);
            ++nacc;
            tc_fence_after();
            if (half == 0) {
                float o[32];
                tc_ld32(tmem + lane_addr + 2 * H, o);
                if (live) {
                    if (NOUT == 1) {
                        out[pt] = o[0] + __ldg(prm.v[iBo]);
                    } else {
                        float r[3] = {o[0] + __ldg(prm.v[iBo] + 0), o[1] + __ldg(prm.v[iBo] + 1), o[2] + __ldg(prm.v[iBo] + 2)};
                        if (affine != nullptr) {
                            float q[3];
#pragma unroll
                            for (int cc = 0; cc < 3; ++cc) {
                                const float lin = r[0] * __ldg(affine + cc) + r[1] * __ldg(affine + 3 + cc) +
                                                  r[2] * __ldg(affine + 6 + cc) + __ldg(affine + 9 + cc);
                                q[cc] = 1.0f / (1.0f + expf(-lin));
                            }
                            r[0] = q[0]; r[1] = q[1]; r[2] = q[2];
                        }
                        out[pt * 3 + 0] = r[0];
                        out[pt * 3 + 1] = r[1];
                        out[pt * 3 + 2] = r[2];
                    }
                }
            }
            tc_fence_before();
        }
    }
    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)G::TMEM_COLS));
    }
}

template <int KIND> int launch_pack(const float* const* params, void* img, cudaStream_t s) {
    Ptrs prm;
    for (int i = 0; i < HPS_MLP_NPARAM; ++i) prm.v[i] = params[i];
    pack_tc_kernel<KIND><<<148, 256, 0, s>>>(prm, reinterpret_cast<uint16_t*>(img));
    return 0;
}

template <int KIND>
int launch_fwd(const float* const* params, const void* img, const float* p, const float* c, const float* affine,
               int64_t P, float* out, float* act, cudaStream_t s) {
    Ptrs prm;
    for (int i = 0; i < HPS_MLP_NPARAM; ++i) prm.v[i] = params[i];
    const int smem = Geo<KIND>::SMEM_BYTES;
    static bool configured = false;
    if (!configured) {
        HPS_CUDA_OK(cudaFuncSetAttribute(mlp_fwd_tc_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const int num_tiles = (int)ceil_div(P, TILE);
    int grid = num_tiles < kNumSMs ? num_tiles : kNumSMs;
    if (const char* e = getenv("HPS_TC_GRID")) grid = atoi(e) < grid ? atoi(e) : grid;   // debug aid
    mlp_fwd_tc_kernel<KIND><<<grid, NTHR, smem, s>>>(prm, reinterpret_cast<const uint8_t*>(img), p, c, affine, P,
                                                     num_tiles, out, act);
    return 0;
}

}  // namespace tc
}  // namespace hps

extern "C" int64_t hps_mlp_tc_packed_bytes(int kind) {
    using namespace hps::tc;
    return kind == HPS_MLP_GEO ? (int64_t)Geo<HPS_MLP_GEO>::SLABS_PER_TILE * SLAB
                               : (int64_t)Geo<HPS_MLP_COL>::SLABS_PER_TILE * SLAB;
}

extern "C" int hps_mlp_tc_pack(int kind, const float* const* params, void* packed, void* stream) {
    using namespace hps;
    HPS_REQUIRE(kind == HPS_MLP_GEO || kind == HPS_MLP_COL, "bad decoder kind");
    HPS_REQUIRE(params && packed, "null buffer");
    HPS_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "packed buffer must be 16-byte aligned");
    for (int i = 0; i < HPS_MLP_NPARAM; ++i) HPS_REQUIRE(params[i] != nullptr, "null parameter pointer");
    int rc = kind == HPS_MLP_GEO ? tc::launch_pack<HPS_MLP_GEO>(params, packed, (cudaStream_t)stream)
                                 : tc::launch_pack<HPS_MLP_COL>(params, packed, (cudaStream_t)stream);
    if (rc) return rc;
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_mlp_tc_fwd(int kind, const float* const* params, const void* packed, const float* p,
                              const float* c, const float* affine, int64_t P, float* out, float* act,
                              void* stream) {
    using namespace hps;
    HPS_REQUIRE(kind == HPS_MLP_GEO || kind == HPS_MLP_COL, "bad decoder kind");
    HPS_REQUIRE(P >= 0 && P < (1ll << 31) * 64, "point count out of range");
    if (P == 0) return 0;
    HPS_REQUIRE(params && packed && p && c && out, "null buffer");
    HPS_REQUIRE(kind == HPS_MLP_COL || affine == nullptr, "affine is colour-only");
    for (int i = 0; i < HPS_MLP_NPARAM; ++i) HPS_REQUIRE(params[i] != nullptr, "null parameter pointer");
    int rc = kind == HPS_MLP_GEO
                 ? tc::launch_fwd<HPS_MLP_GEO>(params, packed, p, c, affine, P, out, act, (cudaStream_t)stream)
                 : tc::launch_fwd<HPS_MLP_COL>(params, packed, p, c, affine, P, out, act, (cudaStream_t)stream);
    if (rc) return rc;
    HPS_LAUNCH_OK();
    return 0;
}
