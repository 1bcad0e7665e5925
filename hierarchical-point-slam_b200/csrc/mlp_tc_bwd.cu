// Decoder MLP backward on tcgen05 tensor cores: a dgrad chain kernel, a streaming wgrad kernel and a small
// finalisation kernel.
//
// Backward of the 5-block decoders (reference: autograd through src/conv_onet/models/decoder.py:316-341 and
// :583-605).  With y_i = act(a_i) + Wc_i c + bc_i, a_i = W_i x_i + b_i, raw = Wo y_4 + bo and dy_i the gradient
// w.r.t. y_i:
//     da_i = dy_i * act'(a_i),   dy_{i-1} = da_i W_i[:, hidden],   dE += da_i W_i[:, embedding],   dy_4 = d raw Wo
//     dW_i = da_i^T x_i,  db_i = colsum(da_i),  dWo = d raw^T y_4,  dbo = colsum(d raw)
//     dc   = sum_i dy_i Wc_i,   dWc_i = dy_i^T c,   dbc_i = colsum(dy_i)
// The last line is evaluated WITHOUT ever materialising dy_i as an operand: dy_i = da_{i+1} W_{i+1} (i < 4), so
//     dc    = sum_{j=1..4} da_j Q_{j-1} + d raw Q_4          Q_{j-1} = W_j[:, hidden] Wc_{j-1},  Q_4 = Wo Wc_4
//     dWc_i = W_{i+1}[:, hidden]^T (da_{i+1}^T c),           dWc_4 = Wo^T (d raw^T c)
//     dbc_i = db_{i+1} W_{i+1}[:, hidden],                   dbc_4 = dbo Wo
// Q is folded into the packed weight image (the dc term rides along the dx MMAs as 32 extra B rows), the small
// products on the second and third line are done by mlp_wfinal_kernel from (H x 32) sums the wgrad kernel leaves in
// a scratch area.  Only da_i is kept in shared memory and dumped: half the operand traffic and half the MMAs of
// the first edition, one conversion instead of two in the epilogue.
//
// mlp_dgrad_tc_kernel walks the chain 4..0 per 128-point tile with the same warp-specialised structure as the
// forward (producer / one MMA lane / 16 epilogue warps / dump warp).  mlp_wgrad_tc_kernel streams the slab images
// (da_i from dgrad; x_i, c from the forward) with cp.async.bulk and contracts over the points with MN-major
// operands, accumulating in TMEM across all tiles of a CTA.  Same split-fp16 arithmetic as the forward.
#include "tc_common.cuh"

namespace hps {
namespace tc {

static __device__ unsigned long long* g_trace_bwd = nullptr;
int set_trace_bwd(void* buf) {
    HPS_CUDA_OK(cudaMemcpyToSymbol(g_trace_bwd, &buf, sizeof(buf)));
    return 0;
}


struct GradPtrs {
    float* v[HPS_MLP_NPARAM];
};

// CTAs of the geometry decoder's dgrad kernel per SM (HPS_GEO_DGRAD_CTAS=1 restores one; read once)
static const int GEO_DGRAD_CTAS_PER_SM = [] {
    const char* e = getenv("HPS_GEO_DGRAD_CTAS");
    return (e != nullptr && e[0] == '1') ? 1 : 2;
}();

template <int KIND> struct GeoB {
    using C = Cfg<KIND>;
    using G = Geo<KIND>;
    static constexpr int HS = G::HS, ES = G::ES, H = C::H, NOUT = C::NOUT;
    static constexpr int NE = 64 * ES;                           // padded embedding width
    static constexpr int NB = H + 32;                            // B rows of a hidden stage: W_j^T (dx) + Q_{j-1}^T (dc)
    // one ring stage holds either a hidden entry (NB rows x 64 h_out columns) or an embedding entry (HS mini-slabs of
    // NE rows x 64 columns)
    static constexpr int STAGE_BYTES = (NB * 128 > NE * 128 * HS) ? NB * 128 : NE * 128 * HS;
    static_assert(STAGE_BYTES % 1024 == 0, "stage alignment");
    static constexpr int ENTRIES = 8 * HS + 4;                   // layers 4..1: 2 HS each; layers 3 and 0: + 2
    static constexpr int64_t IMG_BYTES = (int64_t)ENTRIES * STAGE_BYTES;
    __host__ __device__ static constexpr int first_entry(int layer) {
        return layer == 4 ? 0 : layer == 3 ? 2 * HS : layer == 2 ? 4 * HS + 2 : layer == 1 ? 6 * HS + 2 : 8 * HS + 2;
    }
    // tail of the image: [max|dout| word, 16 B] then fp32 scratch
    //   Q4 [NOUT][32] | T [4][H][32] (da_{j}^T c, j = 1..4) | V [NOUT][32] (d raw^T c) | Sdb [4][H] | Sdbo [4]
    static constexpr int OFF_Q4 = 0, OFF_T = OFF_Q4 + NOUT * 32, OFF_V = OFF_T + 4 * H * 32, OFF_SDB = OFF_V + NOUT * 32,
                         OFF_SDBO = OFF_SDB + 4 * H, SCRATCH_FLOATS = OFF_SDBO + 4;
    static constexpr int64_t TOTAL_BYTES = IMG_BYTES + 16 + (int64_t)SCRATCH_FLOATS * 4;
    // TMEM columns of the dgrad kernel: dx hidden, this layer's dc part, dx embedding (layers 3 and 0 accumulate)
    static constexpr int T_H = 0, T_DCP = H, T_E = H + 32;
    static constexpr int TMEM_COLS = 256;
    static_assert(H + 32 + NE <= 256, "TMEM budget");
    // small fp32 tables in shared memory: Wo [NOUT][H], exposure affine [12], Fourier matrix [3][NF], Q4 [NOUT][32]
    static constexpr int TAB_FLOATS = NOUT * H + 12 + 3 * C::NF + NOUT * 32 + 2;
    // ring stages: the geometry decoder keeps two CTAs per SM (see the kernel's launch bounds), with half the ring each
    static constexpr int NSTB = KIND == HPS_MLP_GEO ? 3 : 6;
    static constexpr int SMEM_BYTES = 1024 + 2 * HS * SLAB + NSTB * STAGE_BYTES + TAB_FLOATS * 4 + 256;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
    static constexpr int ECPT = NE / CG;                         // embedding columns per epilogue thread
    __host__ __device__ static const uint32_t* maxbits(const void* img) {
        return reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(img) + IMG_BYTES);
    }
    __host__ __device__ static float* scratch(void* img) {
        return reinterpret_cast<float*>(static_cast<uint8_t*>(img) + IMG_BYTES + 16);
    }
};

// ---------------------------------------------------------------------------------------------- pack
// Image consumed by the dgrad kernel, per tile, as ENTRIES stage-sized entries in processing order:
//   layer j = 4..1: for every 64-wide h_out slab s: [hi], [lo] with rows 0..H-1   = W_j[h_out][hidden r]  (B rows = r)
//                                                           rows H..H+31  = Q_{j-1}[h_out][k]      (B rows = k)
//                   columns = h_out (the contraction index), K-major SWIZZLE_128B
//   layers 3 and 0: [embedding hi], [embedding lo]: per h_out slab s a mini-slab of NE rows (e index) x 64 columns
// Q_{j-1} = W_j[:, hidden] Wc_{j-1} (j = 1..4), fp32, staged in the T area of the scratch (same shape; the backward
// zeroes that area before it uses it).  One thread per element, 32 consecutive threads share a W row.
template <int KIND>
__global__ void pack_q_kernel(Ptrs prm, float* __restrict__ q) {
    using C = Cfg<KIND>;
    constexpr int H = C::H, E = C::E;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 4 * H * 32) return;
    const int j = idx / (H * 32) + 1, ho = (idx / 32) % H, k = idx % 32;
    const int K = j == 3 ? E + H : H, hoff = j == 3 ? E : 0;
    const float* w = prm.v[iW(j)] + (int64_t)ho * K + hoff;
    const float* wc = prm.v[iWc(j - 1)] + k;
    float acc = 0.f;
    for (int r = 0; r < H; ++r) acc = fmaf(__ldg(w + r), __ldg(wc + r * kC), acc);
    q[idx] = acc;
}

template <int KIND>
__global__ void pack_bwd_kernel(Ptrs prm, uint16_t* __restrict__ img) {
    using C = Cfg<KIND>;
    using B = GeoB<KIND>;
    constexpr int H = C::H, E = C::E, HS = B::HS, NE = B::NE, NOUT = C::NOUT, STG = B::STAGE_BYTES;
    const int total = B::ENTRIES * (STG / 2);
    const float* q = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(img) + B::IMG_BYTES + 16) + B::OFF_T;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int entry = idx / (STG / 2);
        const int byte = (idx % (STG / 2)) * 2;
        int layer = 0;
        for (int l = 4; l >= 1; --l)
            if (entry < B::first_entry(l - 1)) { layer = l; break; }
        const int rel = entry - B::first_entry(layer);
        const int K = layer == 0 ? E : (layer == 3 ? E + H : H);
        const int hoff = layer == 3 ? E : 0;                 // first hidden input column of W_layer
        float val = 0.f;
        bool want_lo = false;
        if (layer >= 1 && rel < 2 * HS) {
            const int s = rel >> 1;
            want_lo = (rel & 1) != 0;
            const int row = ((byte >> 10) << 3) | ((byte >> 7) & 7);
            const int col = ((((byte >> 4) & 7) ^ (row & 7)) << 3) + ((byte >> 1) & 7);
            const int ho = s * 64 + col;
            if (ho < H) {
                const float* w = prm.v[iW(layer)] + (int64_t)ho * K + hoff;
                if (row < H) {
                    val = w[row];
                } else if (row < H + 32) {
                    val = q[((layer - 1) * H + ho) * 32 + (row - H)];      // Q_{layer-1}[ho][k] from pack_q_kernel
                }
            }
        } else {
            want_lo = (rel - (layer >= 1 ? 2 * HS : 0)) == 1;
            const int msz = NE * 128;                                         // bytes per mini-slab
            const int s = byte / msz, b = byte % msz;
            const int row = ((b >> 10) << 3) | ((b >> 7) & 7);                // embedding index e
            const int col = ((((b >> 4) & 7) ^ (row & 7)) << 3) + ((b >> 1) & 7);
            const int h = s * 64 + col;
            if (s < HS && row < E && h < H) val = prm.v[iW(layer)][h * K + row];
        }
        uint16_t hi, lo;
        split16(val, hi, lo);
        img[idx] = want_lo ? lo : hi;
    }
    // tail: max|dout| word = 0, Q4 = Wo Wc_4, remaining scratch = 0 (the backward re-zeroes it before every use)
    uint8_t* tail = reinterpret_cast<uint8_t*>(img) + B::IMG_BYTES;
    float* scr = reinterpret_cast<float*>(tail + 16);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4 + B::SCRATCH_FLOATS; i += gridDim.x * blockDim.x) {
        if (i < 4) {
            reinterpret_cast<uint32_t*>(tail)[i] = 0u;
        } else {
            const int j = i - 4;
            if (j < NOUT * 32) {
                const int o = j / 32, k = j % 32;
                float v = 0.f;
                for (int h = 0; h < H; ++h) v = fmaf(prm.v[iWo][o * H + h], prm.v[iWc(4)][h * kC + k], v);
                scr[j] = v;
            } else if (j >= B::OFF_V) {
                scr[j] = 0.f;          // [OFF_T, OFF_V) holds the staged Q (read above by other threads)
            }
        }
    }
}

// Gradient scaling.  Upstream gradients can be tiny (1e-7 .. 1e-3); fp16 hi/lo operands would lose their low half
// to subnormals.  The whole backward therefore runs on S * gradient with S a power of two chosen so that
// max|dout| lands in [8, 16) (exact), and un-scales on output.  Headroom: the gradients may grow 4,000-fold through the
// five layers before an fp16 high half overflows (round 1 used [128, 256): weights scaled by 4 - a 1,000-fold growth -
// overflowed to NaN, tests/test_gpu_scale.py); at the other end an element 2^-17 of the largest one still has a normal
// high half, and the low halves' subnormal floor (2^-25) is 4e-9 of the largest element.
// The max is reduced on the device into the 4 bytes behind the packed weight image (zeroed by the pack kernel).
// `up`: |upstream gradient| when `dout` holds d loss / d raw for an upstream gradient of 1 (the fused compositing + loss
// kernel writes it in the forward pass) and the true gradient is that times a device scalar.
__device__ __forceinline__ float grad_scale(const uint32_t* maxbits, float up = 1.0f) {
    const float m = __uint_as_float(*maxbits) * up;
    if (!(m > 0.f) || m > 3.0e38f) return 1.0f;
    int e;
    frexpf(m, &e);
    e = 4 - e;
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    return ldexpf(1.0f, e);
}

__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, uint32_t* __restrict__ maxbits) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = fabsf(x[i]);
        if (v <= 3.0e38f) m = fmaxf(m, v);     // ignores NaN / inf
    }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(maxbits, __float_as_uint(m));
}

template <bool RELU> __device__ __forceinline__ float act_grad(float a) {
    if (RELU) return a > 0.f ? 1.f : 0.f;
    // sigmoid(100 a) = 1 / (1 + exp(-100 a)); nn.Softplus's threshold (100 a > 20 -> slope 1) differs by < 2.1e-9
    const float z = ex2_ftz(fminf(a * -144.26950408889634f, 126.0f));         // exp(-100 a), clamped below overflow
    return rcp_ftz(1.0f + z);
}

// MN-major SWIZZLE_128B descriptor for an operand stored rows = contraction index (points), 64 columns per slab:
// LBO = byte distance between consecutive 64-column slabs, SBO = 1024 (8-row groups along the contraction).
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t saddr, uint32_t lbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | (64ull << 32) | (1ull << 46) |
           (2ull << 61);
}
__host__ __device__ constexpr uint32_t instr_desc_mn(int M, int N) {   // both operands MN-major, fp16, f32 accumulate
    return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}


// ---------------------------------------------------------------------------------------------- dgrad
__device__ __forceinline__ float ldg_pinned(const float* p) {   // a load the compiler may neither sink nor re-issue
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int KIND>
__global__ void __launch_bounds__(NTHR, KIND == HPS_MLP_GEO ? 2 : 1)
mlp_dgrad_tc_kernel(Ptrs prm, const uint8_t* __restrict__ wimg, uint8_t* __restrict__ ws,
                    const float* __restrict__ p, const float* __restrict__ affine,
                    const float* __restrict__ dout, int64_t P, int num_tiles, float* __restrict__ dc,
                    float* __restrict__ dp, float* __restrict__ daffine, float* __restrict__ dB, int want_w,
                    const float* __restrict__ gscale) {
    using C = Cfg<KIND>;
    using G = Geo<KIND>;
    using B = GeoB<KIND>;
    const float gu = gscale != nullptr ? __ldg(gscale) : 1.0f;      // upstream gradient of the scalar loss (see grad_scale)
    using W = Ws<KIND>;
    constexpr int H = C::H, E = C::E, NF = C::NF, NOUT = C::NOUT, HS = G::HS, NE = B::NE, STG = B::STAGE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sDAhi = smem;
    uint8_t* sDAlo = sDAhi + HS * SLAB;
    uint8_t* sW = sDAlo + HS * SLAB;
    float* sWo = reinterpret_cast<float*>(sW + B::NSTB * STG);         // [NOUT][H]
    float* sAff = sWo + NOUT * H;                                   // [12]
    float* sB = sAff + 12;                                          // [3][NF]
    float* sQ4 = sB + 3 * NF;                                       // [NOUT][32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sQ4 + NOUT * 32 + ((NOUT * H + 3 * NF + NOUT * 32) & 1));
    uint64_t* full = bars;
    uint64_t* empty = bars + B::NSTB;
    uint64_t* a_ready = bars + 2 * B::NSTB;
    uint64_t* acc_ready = bars + 2 * B::NSTB + 1;
    uint64_t* dump_done = bars + 2 * B::NSTB + 2;     // dump warp -> epilogue: operand slabs copied out, reusable
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * B::NSTB + 3);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float gS = grad_scale(B::maxbits(wimg), fabsf(gu));
    const float gInv = 1.0f / gS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < B::NSTB; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(a_ready, NEPI);
        mbar_init(acc_ready, 1);
        mbar_init(dump_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)B::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    {
        const float* q4 = reinterpret_cast<const float*>(wimg + B::IMG_BYTES + 16) + B::OFF_Q4;
        for (int i = threadIdx.x; i < NOUT * H; i += NTHR) sWo[i] = prm.v[iWo][i];
        for (int i = threadIdx.x; i < 3 * NF; i += NTHR) sB[i] = prm.v[0][i];
        for (int i = threadIdx.x; i < NOUT * 32; i += NTHR) sQ4[i] = q4[i];
        if (threadIdx.x < 12 && affine != nullptr) sAff[threadIdx.x] = affine[threadIdx.x];
    }
    for (int i = threadIdx.x; i < 2 * HS * SLAB / 16; i += NTHR) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int j = 0; j < B::ENTRIES; ++j, ++it) {
                const int st = it % B::NSTB;
                mbar_wait(&empty[st], ((it / B::NSTB) & 1) ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&full[st], STG);
                    bulk_g2s(sW + st * STG, wimg + (size_t)j * STG, STG, &full[st]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t ID_H = instr_desc(FMT_F16, FMT_F16, TILE, B::NB);      // dx (H columns) + dc part (32)
        constexpr uint32_t ID_E = instr_desc(FMT_F16, FMT_F16, TILE, NE);
        const uint32_t t_h = tmem + B::T_H, t_e = tmem + B::T_E;
        const uint32_t aDAh = smem_u32(sDAhi), aDAl = smem_u32(sDAlo);
        const uint32_t aW = smem_u32(sW);
        constexpr int NK = G::ksteps(G::HCOLS);
        uint32_t it = 0, nready = 0;
        unsigned long long* tr = (blockIdx.x == 0 && lane == 0) ? g_trace_bwd : nullptr;
        int ts = 192;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
#pragma unroll 1
            for (int layer = 4; layer >= 0; --layer) {
                mbar_wait(a_ready, nready & 1);
                ++nready;
                tc_fence_after();
                trace_stamp(tr, ts++);
                if (layer >= 1) {   // [dx | dc part] = da_j . [W_j[:, hidden] | Q_{j-1}]
#pragma unroll 1
                    for (int s = 0; s < HS; ++s) {
                        const int st_hi = it % B::NSTB;
                        mbar_wait(&full[st_hi], (it / B::NSTB) & 1);
                        ++it;
                        const int st_lo = it % B::NSTB;
                        mbar_wait(&full[st_lo], (it / B::NSTB) & 1);
                        ++it;
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t wh = aW + st_hi * STG, wl = aW + st_lo * STG;
#pragma unroll 1
                            for (int k = 0; k < NK; ++k) {
                                const uint64_t ah = smem_desc(aDAh + s * SLAB + k * 32), al = smem_desc(aDAl + s * SLAB + k * 32);
                                const uint64_t bh = smem_desc(wh + k * 32), bl = smem_desc(wl + k * 32);
                                tc_mma(t_h, ah, bh, ID_H, (s > 0 || k > 0) ? 1u : 0u);
                                tc_mma(t_h, ah, bl, ID_H, 1u);
                                tc_mma(t_h, al, bh, ID_H, 1u);
                            }
                            tc_commit(&empty[st_hi]);
                            tc_commit(&empty[st_lo]);
                        }
                        __syncwarp();
                    }
                }
                if (layer == 0 || layer == 3) {   // dx (embedding columns) += da_i . W_i[:, embedding]
                    const int st_hi = it % B::NSTB;
                    mbar_wait(&full[st_hi], (it / B::NSTB) & 1);
                    ++it;
                    const int st_lo = it % B::NSTB;
                    mbar_wait(&full[st_lo], (it / B::NSTB) & 1);
                    ++it;
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t wh = aW + st_hi * STG, wl = aW + st_lo * STG;
#pragma unroll 1
                        for (int s = 0; s < HS; ++s) {
#pragma unroll 1
                            for (int k = 0; k < NK; ++k) {
                                const uint64_t ah = smem_desc(aDAh + s * SLAB + k * 32), al = smem_desc(aDAl + s * SLAB + k * 32);
                                const uint64_t bh = smem_desc(wh + s * (NE * 128) + k * 32), bl = smem_desc(wl + s * (NE * 128) + k * 32);
                                // layer 3 starts the embedding gradient, layer 0 accumulates onto it in TMEM
                                tc_mma(t_e, ah, bh, ID_E, (layer == 0 || s > 0 || k > 0) ? 1u : 0u);
                                tc_mma(t_e, ah, bl, ID_E, 1u);
                                tc_mma(t_e, al, bh, ID_E, 1u);
                            }
                        }
                        tc_commit(&empty[st_hi]);
                        tc_commit(&empty[st_lo]);
                    }
                    __syncwarp();
                }
                // all MMAs that read this layer's operands are tracked: the epilogue may overwrite them afterwards
                if (elect_one()) tc_commit(acc_ready);
                __syncwarp();
                trace_stamp(tr, ts++);
            }
        }
    } else if (warp == DUMP_WARP) {
        // ================= dump warp: da_i operand slabs -> workspace (for the weight-gradient kernel) =========
        uint32_t nready = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            uint8_t* gg = W::g_ptr(ws, tile);
#pragma unroll 1
            for (int layer = 4; layer >= 0; --layer) {
                mbar_wait(a_ready, nready & 1);       // the epilogue has written (and proxy-fenced) layer `layer`'s operands
                ++nready;
                if (want_w) {
                    if (elect_one()) {
                        // shared order da_hi[HS] da_lo[HS] == workspace order g_da(layer, 0, .) g_da(layer, 1, .)
                        for (int j = 0; j < 2 * HS; ++j)
                            bulk_s2g(gg + (size_t)(W::g_da(layer, 0, 0) + j) * SLAB, sDAhi + j * SLAB, SLAB);
                        bulk_commit();
                        bulk_wait_read();
                    }
                    __syncwarp();
                }
                if (elect_one()) mbar_arrive(dump_done);
                __syncwarp();
            }
        }
        bulk_wait_all();
    } else {
        // ================= epilogue warps =================
        const int ew = warp - EPI_WARP0;
        const int quad = warp & 3;
        const int cgp = ew >> 2;
        const int row = quad * 32 + lane;
        uint32_t ndump = 0;
        constexpr int CPT = H / CG;               // hidden columns per thread: colour 32, geometry 8
        static_assert(CPT == 32 || CPT == 8, "epilogue column split");
        const int col0 = cgp * CPT;
        constexpr int ECPT = B::ECPT;             // embedding columns per thread: colour 16, geometry 32
        const int ecol0 = cgp * ECPT;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        uint32_t nacc = 0;
        unsigned long long* tr = (blockIdx.x == 0 && ew == 0 && lane == 0) ? g_trace_bwd : nullptr;
        int ts = 128;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            trace_stamp(tr, ts++);
            const int64_t pt = (int64_t)tile * TILE + row;
            const bool live = pt < P;
            uint8_t* gg = W::g_ptr(ws, tile);
            const float* ag = W::a_ptr(ws, tile);

            // pre-activations a_layer (this thread's row, its CPT hidden columns): issued before the wait on the
            // accumulator so that the HBM/L2 latency of the forward workspace hides behind the MMAs
            auto load_a = [&](int layer, float (&av)[CPT]) {
#pragma unroll
                for (int j = 0; j < CPT; ++j) av[j] = ldg_pinned(ag + (layer * H + col0 + j) * TILE + row);
            };
            // dy (this thread's hidden columns) -> da = dy * act'(a): the next A operand (and the wgrad dump)
            auto emit = [&](const float (&av)[CPT], const float (&dy)[CPT]) {
                mbar_wait(dump_done, (ndump & 1) ^ 1);      // the previous operands have been copied out
                ++ndump;
#pragma unroll
                for (int q = 0; q < CPT / 8; ++q) {
                    float a8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) a8[j] = dy[q * 8 + j] * act_grad<C::RELU>(av[q * 8 + j]);
                    uint4 v2, v3;
                    split8(a8, v2, v3);
                    const int colabs = col0 + q * 8;
                    const int slab = colabs >> 6;
                    const uint32_t so = slab_off(row, (colabs & 63) >> 3);
                    *reinterpret_cast<uint4*>(sDAhi + slab * SLAB + so) = v2;
                    *reinterpret_cast<uint4*>(sDAlo + slab * SLAB + so) = v3;
                }
            };

            // ---- output layer (SIMT): d raw, daffine, dy_4 ----
            float dr[3] = {0.f, 0.f, 0.f};
            {
                float a4[CPT];
                load_a(4, a4);
                if (live) {
#pragma unroll
                    for (int o = 0; o < NOUT; ++o) dr[o] = __ldg(dout + pt * NOUT + o) * gu;
                }
                if (NOUT == 3 && affine != nullptr) {
                    const float4 r4 = *reinterpret_cast<const float4*>(W::raw_ptr(ws, tile) + row * 4);
                    const float r[3] = {r4.x, r4.y, r4.z};
                    float dl[3];
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) {
                        const float lin = r[0] * sAff[cc] + r[1] * sAff[3 + cc] + r[2] * sAff[6 + cc] + sAff[9 + cc];
                        const float s = 1.0f / (1.0f + expf(-lin));
                        dl[cc] = dr[cc] * s * (1.0f - s);
                    }
                    if (daffine != nullptr && cgp == 0) {
                        float v[12];
#pragma unroll
                        for (int k = 0; k < 3; ++k)
#pragma unroll
                            for (int cc = 0; cc < 3; ++cc) v[k * 3 + cc] = r[k] * dl[cc];
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) v[9 + cc] = dl[cc];
#pragma unroll
                        for (int k = 0; k < 12; ++k) {
                            float s = v[k];
                            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
                            if (lane == 0) red_add(daffine + k, s);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        dr[k] = dl[0] * sAff[3 * k] + dl[1] * sAff[3 * k + 1] + dl[2] * sAff[3 * k + 2];
                }
                trace_stamp(tr, ts++);   // P1: d raw known
                if (want_w && cgp == 0) {
                    // d raw (scaled) as an operand of the weight-gradient kernel: dWo = y_4^T d raw, dbo = 1^T d raw, ...
                    const float d8[8] = {dr[0] * gS, dr[1] * gS, dr[2] * gS, 0.f, 0.f, 0.f, 0.f, 0.f};
                    uint4 vh, vl;
                    split8(d8, vh, vl);
                    *reinterpret_cast<uint4*>(gg + W::g_dr(0) * SLAB + slab_off(row, 0)) = vh;
                    *reinterpret_cast<uint4*>(gg + W::g_dr(1) * SLAB + slab_off(row, 0)) = vl;
                }
                float dy[CPT];
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    float s = 0.f;
#pragma unroll
                    for (int o = 0; o < NOUT; ++o) s = fmaf(dr[o], sWo[o * H + col0 + j], s);
                    dy[j] = s * gS;      // the chain below runs on scaled gradients
                }
                trace_stamp(tr, ts++);   // P2: dy_4
                emit(a4, dy);
            }
            fence_proxy_async();
            mbar_arrive(a_ready);
            trace_stamp(tr, ts++);

            // dc: the output-layer term d raw Q_4 (unscaled) + the per-layer parts read from TMEM (scaled)
            float dcacc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) dcacc[j] = 0.f;
#pragma unroll 1
            for (int layer = 4; layer >= 1; --layer) {
                float an[CPT];
                load_a(layer - 1, an);
                mbar_wait(acc_ready, nacc & 1);
                ++nacc;
                tc_fence_after();
                trace_stamp(tr, ts++);
                float dy[CPT];
                if constexpr (CPT == 32) tc_ld32(tmem + lane_addr + B::T_H + col0, dy);
                else tc_ld8(tmem + lane_addr + B::T_H + col0, dy);
                {
                    float v[8];
                    tc_ld8(tmem + lane_addr + B::T_DCP + cgp * 8, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dcacc[j] += v[j];
                }
                emit(an, dy);
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(a_ready);
                trace_stamp(tr, ts++);
            }
            {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float s = dcacc[j] * gInv;
#pragma unroll
                    for (int o = 0; o < NOUT; ++o) s = fmaf(dr[o], sQ4[o * 32 + cgp * 8 + j], s);
                    v[j] = s;
                }
                if (live) {
                    *reinterpret_cast<float4*>(dc + pt * kC + cgp * 8) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(dc + pt * kC + cgp * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
            // ---- end of the chain: the embedding gradient ----
            float x0 = 0.f, x1 = 0.f, x2 = 0.f;
            const bool want_b = (KIND == HPS_MLP_GEO) && dB != nullptr;
            if ((dp != nullptr || want_b) && live) {
                x0 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 0));
                x1 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 1));
                x2 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 2));
            }
            mbar_wait(acc_ready, nacc & 1);
            ++nacc;
            tc_fence_after();
            trace_stamp(tr, ts++);
            // ---- dE -> dp, dB ----
            if (dp != nullptr || want_b) {
                float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll 1
                for (int eb = 0; eb < ECPT; eb += 8) {
                    const int e0 = ecol0 + eb;               // warp-uniform
                    if (e0 >= E) break;
                    float de[8];       // d loss / d embedding (layers 3 and 0 summed in TMEM)
                    tc_ld8(tmem + lane_addr + B::T_E + e0, de);
                    float gb0[8], gb1[8], gb2[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int e = e0 + j;
                        float g = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
                        if (e < E) {
                            const bool is_cos = C::CONCAT && e >= NF;
                            const int f = is_cos ? e - NF : e;
                            b0 = sB[f];
                            b1 = sB[NF + f];
                            b2 = sB[2 * NF + f];
                            const float arg = fmaf(x2, b2, fmaf(x1, b1, x0 * b0));
                            // d sin = cos, d cos = -sin
                            g = live ? (is_cos ? -de[j] * sin_2pi(arg) : de[j] * cos_2pi(arg)) * gInv : 0.f;
                        }
                        g0 = fmaf(g * kTwoPi, b0, g0);
                        g1 = fmaf(g * kTwoPi, b1, g1);
                        g2 = fmaf(g * kTwoPi, b2, g2);
                        gb0[j] = g * x0;
                        gb1[j] = g * x1;
                        gb2[j] = g * x2;
                    }
                    if (want_b) {   // dB[d][f] += sum_pt g * x'_d   (geometry decoder: e == f)
                        const float s0 = warp_colsum8(gb0, lane), s1 = warp_colsum8(gb1, lane), s2 = warp_colsum8(gb2, lane);
                        const int e = e0 + colsum8_col(lane);
                        if (lane < 8 && e < E) {
                            red_add(dB + e, s0);
                            red_add(dB + NF + e, s1);
                            red_add(dB + 2 * NF + e, s2);
                        }
                    }
                }
                if (dp != nullptr && live) {   // dp is zero-initialised by the caller; the column groups add
                    red_add(dp + pt * 3 + 0, g0);
                    red_add(dp + pt * 3 + 1, g1);
                    red_add(dp + pt * 3 + 2, g2);
                }
            }
            tc_fence_before();     // the TMEM reads above precede the next tile's MMAs
            trace_stamp(tr, ts++);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)B::TMEM_COLS));
    }
}

// ---------------------------------------------------------------------------------------------- wgrad
template <int KIND> struct GeoW {
    using C = Cfg<KIND>;
    using G = Geo<KIND>;
    static constexpr int HS = G::HS, ES = G::ES, H = C::H, NE = 64 * ES, NH = 64 * HS;
    // One pipeline stage = a quarter tile (32 points = two K=16 steps) of every operand of one role, as 4 KB
    // chunks (32 rows x 128 B of a swizzled slab; the swizzle period is 8 rows, so a chunk is itself a valid
    // SWIZZLE_128B MN-major operand with atom stride CH).  Roles 0..4 (layer i):
    //   A chunks: da_hi[HS] da_lo[HS]                             (same order as the G region of the workspace)
    //   B chunks: y_hi[HS] e_hi[ES] y_lo[HS] e_lo[ES] c[1]        (so [y | e] is one N = NH+NE operand)
    // Role 5 (output layer): chunks y4_hi[HS] y4_lo[HS] c draw_hi draw_lo.
    static constexpr int CH = 4096;
    static constexpr int A_CHUNKS = 2 * HS, B_CHUNKS = 2 * (HS + ES) + 1;
    static constexpr int STAGE_BYTES = (A_CHUNKS + B_CHUNKS) * CH;
    static constexpr int NSTW = 4;
    static constexpr int NROLE = 6;
    // + CH: M=128 may read one atom past the last chunk; + CH: the chunk of ones (column sums)
    static constexpr int SMEM_BYTES = 1024 + NSTW * STAGE_BYTES + 2 * CH + 256;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
    // TMEM columns: dW hidden [0,NH), dW embedding [NH,NH+NE), T = da^T [c_hi | c_lo] [NH+NE, +64), db (16 equal
    // columns); role 5: dWo^T [0,16), V^T = c^T d raw [16,32), dbo [32,48)
    static constexpr int T_WH = 0, T_WE = NH, T_T = NH + NE, T_B = NH + NE + 64;
    static constexpr int TMEM_COLS = 512;
    static_assert(NH + NE + 80 <= 512, "TMEM budget");
    static constexpr int NTHRW = 320;      // producer warp, MMA warp, 8 flush warps
    // 4 KB chunks a role streams per quarter tile: the roles are not equally heavy (layer 3 also reads the embedding,
    // layer 0 reads no hidden slab and no c), so the CTAs are dealt out in proportion to this
    __host__ __device__ static constexpr int role_chunks(int role) {
        return role == 0 ? 2 * HS + 2 * ES
             : role == 3 ? 2 * HS + 2 * (HS + ES) + 1
             : role == 5 ? 2 * HS + 1 + 2
                         : 2 * HS + 2 * HS + 1;
    }
};

// first CTA of every role (begin[NROLE] = grid size): CTA b has the role r with begin[r] <= b < begin[r + 1]
struct RoleSplit {
    int begin[7];
};

// Deal `grid` (>= NROLE) CTAs out to the roles so that the heaviest per-CTA share is as small as possible: one
// each, then one at a time to the role with the most chunks per CTA.
template <int KIND> RoleSplit split_roles(int grid) {
    using GW = GeoW<KIND>;
    int n[GW::NROLE];
    for (int r = 0; r < GW::NROLE; ++r) n[r] = 1;
    for (int used = GW::NROLE; used < grid; ++used) {
        int best = 0;
        for (int r = 1; r < GW::NROLE; ++r)
            if ((int64_t)GW::role_chunks(r) * n[best] > (int64_t)GW::role_chunks(best) * n[r]) best = r;
        ++n[best];
    }
    RoleSplit s;
    s.begin[0] = 0;
    for (int r = 0; r < GW::NROLE; ++r) s.begin[r + 1] = s.begin[r] + n[r];
    return s;
}

// Everything that is a sum over the points.  The CTAs of a role (RoleSplit) share its tiles round robin: each streams
// the dgrad dumps (da_i, d raw) and the forward's operand dumps (y_i, embedding, c) through a 4-stage
// quarter-tile ring and accumulates over all its tiles in TMEM (contraction over the points: both operands
// MN-major)
//     roles 0..4:  dW_i = da_i^T [y_{i-1} | e],  db_i = da_i^T 1,  T_i = da_i^T c  (i >= 1, for dWc_{i-1})
//     role 5:      dWo^T = y_4^T d raw,  V^T = c^T d raw (for dWc_4),  dbo = d raw^T 1
// then adds its partial sums to the gradient buffers / the scratch area (red.global.add, 16-byte vectors where the
// pitch allows).
template <int KIND>
__global__ void __launch_bounds__(GeoW<KIND>::NTHRW, 1)
mlp_wgrad_tc_kernel(GradPtrs grd, const uint8_t* __restrict__ ws, int num_tiles, const uint32_t* __restrict__ maxbits,
                    float* __restrict__ scratch, const RoleSplit split, const float* __restrict__ gscale) {
    using C = Cfg<KIND>;
    using GW = GeoW<KIND>;
    using GB = GeoB<KIND>;
    using W = Ws<KIND>;
    constexpr int H = C::H, E = C::E, NOUT = C::NOUT, HS = GW::HS, ES = GW::ES, NH = GW::NH, NE = GW::NE, CH = GW::CH,
                  NSTW = GW::NSTW;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sOnes = smem + NSTW * GW::STAGE_BYTES + CH;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + CH);
    uint64_t* full = bars;
    uint64_t* empty = bars + NSTW;
    uint64_t* done = bars + 2 * NSTW;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTW + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int role = 0;
#pragma unroll
    for (int r = 1; r < GW::NROLE; ++r) role += (int)blockIdx.x >= split.begin[r] ? 1 : 0;
    const int layer = role;                                   // roles 0..4
    const int rank = (int)blockIdx.x - split.begin[role];
    const int gsize = split.begin[role + 1] - split.begin[role];               // CTAs with this role
    const bool has_h = role >= 1 && role < 5, has_e = (role == 0 || role == 3), has_c = role >= 1;
    const bool mine = rank < num_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTW; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)GW::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // operands that a role does not load (and the over-read atom) must still be finite: zero everything once
    for (int i = threadIdx.x; i < (NSTW * GW::STAGE_BYTES + CH) / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < CH / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);   // fp16 1.0
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ---- producer ----
        const uint32_t bytes = role == 5 ? (uint32_t)CH * (2 * HS + 3)
                                         : (uint32_t)CH * (2 * HS + (has_h ? 2 * HS : 0) + (has_e ? 2 * ES : 0) + (has_c ? 1 : 0));
        uint32_t it = 0;
        const uint64_t l2pol = l2_policy_evict_first();       // the slabs are read exactly once
        for (int tile = rank; tile < num_tiles; tile += gsize) {
            const uint8_t* xg = W::x_ptr(const_cast<uint8_t*>(ws), tile);
            const uint8_t* gt = W::g_ptr(const_cast<uint8_t*>(ws), tile);
            for (int q = 0; q < 4; ++q, ++it) {
                const int st = it % NSTW;
                if (it >= NSTW) mbar_wait(&empty[st], ((it / NSTW) - 1) & 1);
                if (elect_one()) {
                    uint8_t* sa = smem + st * GW::STAGE_BYTES;
                    uint8_t* sb = sa + GW::A_CHUNKS * CH;
                    mbar_expect_tx(&full[st], bytes);
                    if (role == 5) {
                        for (int j = 0; j < 2 * HS; ++j)      // y4 hi[HS], lo[HS] are consecutive X slabs
                            bulk_g2s_hint(sa + j * CH, xg + (size_t)(W::x_y(4, 0, 0) + j) * SLAB + q * CH, CH, &full[st], l2pol);
                        bulk_g2s_hint(sb, xg + (size_t)W::x_c() * SLAB + q * CH, CH, &full[st], l2pol);
                        bulk_g2s_hint(sb + CH, gt + (size_t)W::g_dr(0) * SLAB + q * CH, CH, &full[st], l2pol);
                        bulk_g2s_hint(sb + 2 * CH, gt + (size_t)W::g_dr(1) * SLAB + q * CH, CH, &full[st], l2pol);
                    } else {
                        const uint8_t* gg = gt + (size_t)W::g_da(layer, 0, 0) * SLAB;
                        for (int j = 0; j < 2 * HS; ++j) bulk_g2s_hint(sa + j * CH, gg + (size_t)j * SLAB + q * CH, CH, &full[st], l2pol);
                        if (has_h) {
                            for (int s2 = 0; s2 < HS; ++s2) {
                                bulk_g2s_hint(sb + s2 * CH, xg + (size_t)W::x_y(layer - 1, 0, s2) * SLAB + q * CH, CH, &full[st], l2pol);
                                bulk_g2s_hint(sb + (HS + ES + s2) * CH, xg + (size_t)W::x_y(layer - 1, 1, s2) * SLAB + q * CH, CH,
                                         &full[st], l2pol);
                            }
                        }
                        if (has_e) {
                            for (int s2 = 0; s2 < ES; ++s2) {
                                bulk_g2s_hint(sb + (HS + s2) * CH, xg + (size_t)W::x_e(0, s2) * SLAB + q * CH, CH, &full[st], l2pol);
                                bulk_g2s_hint(sb + (2 * HS + ES + s2) * CH, xg + (size_t)W::x_e(1, s2) * SLAB + q * CH, CH,
                                         &full[st], l2pol);
                            }
                        }
                        if (has_c) bulk_g2s_hint(sb + 2 * (HS + ES) * CH, xg + (size_t)W::x_c() * SLAB + q * CH, CH, &full[st], l2pol);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ---- MMA issue ----
        const int nW = (has_h ? NH : 0) + (has_e ? NE : 0);
        const uint32_t id_w = instr_desc_mn(TILE, nW > 0 ? nW : 16), id_c = instr_desc_mn(TILE, 64);
        const uint32_t id_16 = instr_desc_mn(TILE, 16);
        const uint32_t t_w = tmem + (has_h ? GW::T_WH : GW::T_WE), t_t = tmem + GW::T_T, t_b = tmem + GW::T_B;
        const uint32_t b_off = (uint32_t)(has_h ? 0 : HS) * CH;       // layer 0: B starts at the embedding chunks
        const uint64_t ones = smem_desc_mn(smem_u32(sOnes), CH);
        uint32_t it = 0;
        for (int tile = rank; tile < num_tiles; tile += gsize) {
            for (int q = 0; q < 4; ++q, ++it) {
                const int st = it % NSTW;
                mbar_wait(&full[st], (it / NSTW) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = smem_u32(smem + st * GW::STAGE_BYTES);
                    const uint32_t sb = sa + GW::A_CHUNKS * CH;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
                        if (role == 5) {
                            const uint64_t yh = smem_desc_mn(sa + k * 2048, CH), yl = smem_desc_mn(sa + HS * CH + k * 2048, CH);
                            const uint64_t cc = smem_desc_mn(sb + k * 2048, CH);
                            const uint64_t rh = smem_desc_mn(sb + CH + k * 2048, CH), rl = smem_desc_mn(sb + 2 * CH + k * 2048, CH);
                            tc_mma(tmem, yh, rh, id_16, acc);           // dWo^T
                            tc_mma(tmem, yh, rl, id_16, 1u);
                            tc_mma(tmem, yl, rh, id_16, 1u);
                            tc_mma(tmem + 16, cc, rh, id_16, acc);      // V^T: lanes = [c_hi | c_lo] columns
                            tc_mma(tmem + 16, cc, rl, id_16, 1u);
                            tc_mma(tmem + 32, rh, ones, id_16, acc);    // dbo: lanes = output channel
                            tc_mma(tmem + 32, rl, ones, id_16, 1u);
                        } else {
                            const uint64_t dah = smem_desc_mn(sa + k * 2048, CH), dal = smem_desc_mn(sa + HS * CH + k * 2048, CH);
                            const uint64_t bh = smem_desc_mn(sb + b_off + k * 2048, CH);
                            const uint64_t bl = smem_desc_mn(sb + (HS + ES) * CH + b_off + k * 2048, CH);
                            tc_mma(t_w, dah, bh, id_w, acc);
                            tc_mma(t_w, dah, bl, id_w, 1u);
                            tc_mma(t_w, dal, bh, id_w, 1u);
                            if (has_c) {
                                const uint64_t bc = smem_desc_mn(sb + 2 * (HS + ES) * CH + k * 2048, CH);
                                tc_mma(t_t, dah, bc, id_c, acc);      // c chunk = [c_hi | c_lo]: the halves are summed at flush
                                tc_mma(t_t, dal, bc, id_c, 1u);
                            }
                            tc_mma(t_b, dah, ones, id_16, acc);       // bias gradient: column sums over the points
                            tc_mma(t_b, dal, ones, id_16, 1u);
                        }
                    }
                    tc_commit(&empty[st]);
                }
                __syncwarp();
            }
        }
        if (mine) {
            if (elect_one()) tc_commit(done);
            __syncwarp();
        }
    } else if (mine) {
        // ---- flush: TMEM lane = output feature h, columns = input feature ----
        mbar_wait(done, 0);
        tc_fence_after();
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int h = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const float gInv = 1.0f / grad_scale(maxbits, gscale != nullptr ? fabsf(__ldg(gscale)) : 1.0f);
        if (role == 5) {
            if (half == 0) {
                float v[8], b[8];
                tc_ld8(tmem + lane_addr, v);
                tc_ld8(tmem + lane_addr + 32, b);
                if (h < H) {
#pragma unroll
                    for (int o = 0; o < NOUT; ++o) red_add(grd.v[iWo] + o * H + h, v[o] * gInv);
                }
                if (h < NOUT) {
                    red_add(grd.v[iBo] + h, b[0] * gInv);
                    red_add(scratch + GB::OFF_SDBO + h, b[0] * gInv);
                }
            } else {
                float v[8];
                tc_ld8(tmem + lane_addr + 16, v);
                if (h < 64) {       // lane = column of the c slab: hi part 0..31, lo part 32..63
#pragma unroll
                    for (int o = 0; o < NOUT; ++o) red_add(scratch + GB::OFF_V + o * 32 + (h & 31), v[o] * gInv);
                }
            }
        } else {
            const int K = layer == 0 ? E : (layer == 3 ? E + H : H);
            float* gw = grd.v[iW(layer)];
            const bool vec_w = ((reinterpret_cast<uintptr_t>(gw) & 15) == 0) && (K % 4 == 0) && (E % 4 == 0);
            auto flush16 = [&](float* dst, const float (&v)[16], int n_valid, bool vec) {     // dst: this row, 16 columns
                if (vec && n_valid == 16) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        red_add_v4(dst + j, make_float4(v[j] * gInv, v[j + 1] * gInv, v[j + 2] * gInv, v[j + 3] * gInv));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (j < n_valid) red_add(dst + j, v[j] * gInv);
                }
            };
            if (has_h) {
                for (int cb = half * (NH / 2); cb < (half + 1) * (NH / 2); cb += 16) {
                    float v[16];
                    tc_ld16(tmem + lane_addr + GW::T_WH + cb, v);
                    const int nv = H - cb < 16 ? H - cb : 16;
                    if (h < H && nv > 0) flush16(gw + (int64_t)h * K + (layer == 3 ? E : 0) + cb, v, nv, vec_w);
                }
            }
            if (has_e) {
                for (int cb = half * (NE / 2); cb < (half + 1) * (NE / 2); cb += 16) {
                    float v[16];
                    tc_ld16(tmem + lane_addr + GW::T_WE + cb, v);
                    const int nv = E - cb < 16 ? E - cb : 16;
                    if (h < H && nv > 0) flush16(gw + (int64_t)h * K + cb, v, nv, vec_w);
                }
            }
            if (has_c) {   // T_layer[h][k] = sum_pt da[pt][h] c[pt][k]  ->  scratch (dWc_{layer-1} = W_layer^T T_layer)
                float v[16], v2[16];
                tc_ld16(tmem + lane_addr + GW::T_T + half * 16, v);
                tc_ld16(tmem + lane_addr + GW::T_T + 32 + half * 16, v2);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += v2[j];
                if (h < H) flush16(scratch + GB::OFF_T + ((layer - 1) * H + h) * 32 + half * 16, v, 16, true);
            }
            if (half == 0) {
                float v[8];
                tc_ld8(tmem + lane_addr + GW::T_B, v);
                if (h < H) {
                    red_add(grd.v[iB(layer)] + h, v[0] * gInv);
                    if (layer >= 1) red_add(scratch + GB::OFF_SDB + (layer - 1) * H + h, v[0] * gInv);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)GW::TMEM_COLS));
    }
}

// dWc_i = W_{i+1}[:, hidden]^T T_{i+1},  dbc_i = db_{i+1} W_{i+1}[:, hidden]  (i = 0..3);  dWc_4 = Wo^T V,  dbc_4 = dbo Wo.
// WFINAL_PARTS threads per output element, each with a quarter of the H-term dot product over L2-resident weights
// (fully unrolled: all its loads are in flight at once - the kernel is pure latency), combined by the red.add that
// writes the gradient anyway.
constexpr int WFINAL_PARTS = 4;
template <int KIND>
__global__ void __launch_bounds__(128) mlp_wfinal_kernel(Ptrs prm, GradPtrs grd, const float* __restrict__ scratch) {
    using C = Cfg<KIND>;
    using GB = GeoB<KIND>;
    constexpr int H = C::H, E = C::E, NOUT = C::NOUT, HQ = H / WFINAL_PARTS;
    static_assert(H % WFINAL_PARTS == 0, "hidden width must split evenly");
    const int per = H * 32 + H;                               // outputs per layer: dWc (H x 32) + dbc (H)
    for (int tid = blockIdx.x * blockDim.x + threadIdx.x; tid < WFINAL_PARTS * 5 * per; tid += gridDim.x * blockDim.x) {
        const int part = tid / (5 * per), idx = tid - part * (5 * per);
        const int i = idx / per, r = idx % per;
        const bool bias = r >= H * 32;
        const int hin = bias ? r - H * 32 : r / 32, k = r % 32;
        float acc = 0.f;
        if (i < 4) {
            const int Kn = (i + 1) == 3 ? E + H : H, off = (i + 1) == 3 ? E : 0;
            const float* w = prm.v[iW(i + 1)] + off + hin;                    // W_{i+1}[ho][off + hin], pitch Kn
            if (bias) {
                const float* sdb = scratch + GB::OFF_SDB + i * H;
                float a[HQ], b[HQ];
#pragma unroll
                for (int j = 0; j < HQ; ++j) {               // every load issued before the first use
                    const int ho = part * HQ + j;
                    a[j] = __ldg(sdb + ho);
                    b[j] = __ldg(w + (int64_t)ho * Kn);
                }
#pragma unroll
                for (int j = 0; j < HQ; ++j) acc = fmaf(a[j], b[j], acc);
            } else {
                const float* t = scratch + GB::OFF_T + i * H * 32 + k;
                float a[HQ], b[HQ];
#pragma unroll
                for (int j = 0; j < HQ; ++j) {
                    const int ho = part * HQ + j;
                    a[j] = __ldg(w + (int64_t)ho * Kn);
                    b[j] = __ldg(t + ho * 32);
                }
#pragma unroll
                for (int j = 0; j < HQ; ++j) acc = fmaf(a[j], b[j], acc);
            }
        } else if (part == 0) {
            const float* wo = prm.v[iWo] + hin;
            if (bias) {
                for (int o = 0; o < NOUT; ++o) acc = fmaf(scratch[GB::OFF_SDBO + o], wo[o * H], acc);
            } else {
                for (int o = 0; o < NOUT; ++o) acc = fmaf(wo[o * H], scratch[GB::OFF_V + o * 32 + k], acc);
            }
        }
        if (i == 4 && part != 0) continue;                    // the output-layer terms are three products: one thread
        if (bias) red_add(grd.v[iBc(i)] + hin, acc);
        else red_add(grd.v[iWc(i)] + hin * kC + k, acc);
    }
}

// ---------------------------------------------------------------------------------------------- launch
template <int KIND> int launch_pack_bwd(const float* const* params, void* img, cudaStream_t s) {
    Ptrs prm;
    for (int i = 0; i < HPS_MLP_NPARAM; ++i) prm.v[i] = params[i];
    float* q = GeoB<KIND>::scratch(img) + GeoB<KIND>::OFF_T;
    // 128-thread CTAs (4 K registers): they fit beside a resident decoder CTA (58 of the SM's 64 K registers), so the
    // images for the backward are packed WHILE the forward decoders run instead of queueing for a free SM behind them
    // (round 2 timeline: with 256-thread CTAs this branch ended 20 us after the forward and held up the compositing)
    pack_q_kernel<KIND><<<(4 * Cfg<KIND>::H * 32 + 127) / 128, 128, 0, s>>>(prm, q);
    pack_bwd_kernel<KIND><<<296, 128, 0, s>>>(prm, reinterpret_cast<uint16_t*>(img));
    return 0;
}

template <int KIND>
int launch_bwd(const float* const* params, const void* img, void* ws, const float* p, const float* affine,
               const float* dout, int64_t P, float* const* dparams, float* dc, float* dp, float* daffine,
               int absmax_done, int max_ctas, cudaStream_t wstream, cudaStream_t s, const float* gscale) {
    using GB = GeoB<KIND>;
    Ptrs prm;
    GradPtrs grd;
    for (int i = 0; i < HPS_MLP_NPARAM; ++i) {
        prm.v[i] = params[i];
        grd.v[i] = dparams ? dparams[i] : nullptr;
    }
    if (dparams)
        for (int i = 1; i < HPS_MLP_NPARAM; ++i) HPS_REQUIRE(grd.v[i] != nullptr, "dparams: all Linear gradients or none");
    static bool configured = false;
    if (!configured) {
        HPS_CUDA_OK(cudaFuncSetAttribute(mlp_dgrad_tc_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         GB::SMEM_BYTES));
        HPS_CUDA_OK(cudaFuncSetAttribute(mlp_wgrad_tc_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         GeoW<KIND>::SMEM_BYTES));
        // The weight-gradient CTAs ask for 153 / 185 KB.  With the default carve-out the SM is configured for the next
        // size up (164 / 196 KB) and nothing that needs shared memory fits beside them; with the largest carve-out
        // (228 KB) the neighbour search of the next batch (13 KB per CTA) runs beside this HBM-bound kernel.
        HPS_CUDA_OK(cudaFuncSetAttribute(mlp_wgrad_tc_kernel<KIND>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
        configured = true;
    }
    const int num_tiles = (int)ceil_div(P, TILE);
    // dgrad only (wgrad is HBM-bound).  The geometry decoder's dgrad CTAs are small enough (48 registers, 83 KB, 256 TMEM
    // columns) for two per SM: its tiles are latency-bound chains of tiny MMAs and 8-column epilogues, and a second
    // resident CTA fills the bubbles of the first.
    const int per_sm = KIND == HPS_MLP_GEO ? GEO_DGRAD_CTAS_PER_SM : 1;
    const int lim = kNumSMs * per_sm;
    const int cap = (max_ctas > 0 && max_ctas * per_sm < lim) ? max_ctas * per_sm : lim;
    int grid = num_tiles < cap ? num_tiles : cap;
    void* img_rw = const_cast<void*>(img);
    uint32_t* maxbits = const_cast<uint32_t*>(GB::maxbits(img));
    float* scratch = GB::scratch(img_rw);
    if (!absmax_done) {
        const int64_t n = P * Cfg<KIND>::NOUT;
        const int nb = (int)(ceil_div(n, 1024) < kNumSMs ? ceil_div(n, 1024) : kNumSMs);
        absmax_kernel<<<nb, 1024, 0, s>>>(dout, n, maxbits);
    }
    mlp_dgrad_tc_kernel<KIND><<<grid, NTHR, GB::SMEM_BYTES, s>>>(
        prm, reinterpret_cast<const uint8_t*>(img), reinterpret_cast<uint8_t*>(ws), p, affine, dout, P, num_tiles, dc, dp,
        daffine, grd.v[0], dparams ? 1 : 0, gscale);
    if (dparams) {
        // The weight-gradient kernels may run on their own stream (ordered after the dgrad kernel by an event), so
        // that whatever the caller enqueues next on `s` - the feature scatter needs only dc - overlaps with them.
        cudaStream_t w = s;
        if (wstream != nullptr && wstream != s) {
            static cudaEvent_t dgrad_done = nullptr;          // one per decoder kind (template instantiation)
            if (dgrad_done == nullptr) HPS_CUDA_OK(cudaEventCreateWithFlags(&dgrad_done, cudaEventDisableTiming));
            HPS_CUDA_OK(cudaEventRecord(dgrad_done, s));
            HPS_CUDA_OK(cudaStreamWaitEvent(wstream, dgrad_done, 0));
            w = wstream;
        }
        // the sums of this call only (T, V, Sdb, Sdbo; Q4 stays)
        HPS_CUDA_OK(cudaMemsetAsync(scratch + GB::OFF_T, 0, (size_t)(GB::SCRATCH_FLOATS - GB::OFF_T) * 4, w));
        const int wgrid = GeoW<KIND>::NROLE * num_tiles < kNumSMs ? GeoW<KIND>::NROLE * num_tiles : kNumSMs;   // every role has a CTA
        mlp_wgrad_tc_kernel<KIND><<<wgrid, GeoW<KIND>::NTHRW, GeoW<KIND>::SMEM_BYTES, w>>>(
            grd, reinterpret_cast<const uint8_t*>(ws), num_tiles, maxbits, scratch, split_roles<KIND>(wgrid), gscale);
        const int nout = 5 * (Cfg<KIND>::H * 32 + Cfg<KIND>::H);
        HPS_SHARE_SM(mlp_wfinal_kernel<KIND>);
        mlp_wfinal_kernel<KIND><<<(WFINAL_PARTS * nout + 127) / 128, 128, 0, w>>>(prm, grd, scratch);
    }
    return 0;
}

}  // namespace tc
}  // namespace hps

extern "C" int64_t hps_mlp_tc_bwd_packed_bytes(int kind) {
    using namespace hps::tc;
    return kind == HPS_MLP_GEO ? GeoB<HPS_MLP_GEO>::TOTAL_BYTES : GeoB<HPS_MLP_COL>::TOTAL_BYTES;
}

extern "C" int64_t hps_mlp_tc_bwd_absmax_offset(int kind) {
    using namespace hps::tc;
    return kind == HPS_MLP_GEO ? GeoB<HPS_MLP_GEO>::IMG_BYTES : GeoB<HPS_MLP_COL>::IMG_BYTES;
}

extern "C" int hps_mlp_tc_bwd_pack(int kind, const float* const* params, void* packed, void* stream) {
    using namespace hps;
    HPS_REQUIRE(kind == HPS_MLP_GEO || kind == HPS_MLP_COL, "bad decoder kind");
    HPS_REQUIRE(params && packed, "null buffer");
    for (int i = 0; i < HPS_MLP_NPARAM; ++i) HPS_REQUIRE(params[i] != nullptr, "null parameter pointer");
    int rc = kind == HPS_MLP_GEO ? tc::launch_pack_bwd<HPS_MLP_GEO>(params, packed, (cudaStream_t)stream)
                                 : tc::launch_pack_bwd<HPS_MLP_COL>(params, packed, (cudaStream_t)stream);
    if (rc) return rc;
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_mlp_tc_bwd(int kind, const float* const* params, const void* packed, void* ws, const float* p,
                              const float* affine, const float* dout, int64_t P, float* const* dparams, float* dc,
                              float* dp, float* daffine, int absmax_done, int max_ctas, void* wgrad_stream,
                              const float* gscale, void* stream) {
    using namespace hps;
    HPS_REQUIRE(kind == HPS_MLP_GEO || kind == HPS_MLP_COL, "bad decoder kind");
    HPS_REQUIRE(P >= 0, "negative point count");
    if (P == 0) return 0;
    HPS_REQUIRE(params && packed && ws && p && dout && dc, "null buffer");
    HPS_REQUIRE(gscale == nullptr || (affine == nullptr && absmax_done), "gscale: without the exposure tail, with max|dout| folded");
    for (int i = 0; i < HPS_MLP_NPARAM; ++i) HPS_REQUIRE(params[i] != nullptr, "null parameter pointer");
    int rc = kind == HPS_MLP_GEO ? tc::launch_bwd<HPS_MLP_GEO>(params, packed, ws, p, affine, dout, P, dparams, dc, dp,
                                                               daffine, absmax_done, max_ctas, (cudaStream_t)wgrad_stream,
                                                               (cudaStream_t)stream, gscale)
                                 : tc::launch_bwd<HPS_MLP_COL>(params, packed, ws, p, affine, dout, P, dparams, dc, dp,
                                                               daffine, absmax_done, max_ctas, (cudaStream_t)wgrad_stream,
                                                               (cudaStream_t)stream, gscale);
    if (rc) return rc;
    HPS_LAUNCH_OK();
    return 0;
}
