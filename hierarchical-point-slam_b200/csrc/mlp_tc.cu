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
#include "tc_common.cuh"

namespace hps {
namespace tc {

static __device__ unsigned long long* g_trace_fwd = nullptr;
int set_trace_fwd(void* buf) {
    HPS_CUDA_OK(cudaMemcpyToSymbol(g_trace_fwd, &buf, sizeof(buf)));
    return 0;
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
    // softplus(beta=100): max(a,0) + log(1 + exp(-|100 a|)) / 100   (absolute error < 1e-9): 6 instructions
    const float e = ex2_ftz(fabsf(a) * -144.26950408889634f);               // exp(-|100 a|), 100 * log2(e)
    return fmaf(lg2_ftz(1.0f + e), 0.0069314718055994531f, fmaxf(a, 0.f));    // ln(2) / 100
}

// DUMP: 0 = inference, 1 = SIMT-layout activations for mlp_bwd_kernel (act), 2 = workspace for the tensor-core backward (ws)
template <int KIND, int DUMP>
__global__ void __launch_bounds__(NTHR, 1)
mlp_fwd_tc_kernel(Ptrs prm, const uint8_t* __restrict__ wimg, const float* __restrict__ p,
                  const float* __restrict__ c, const float* __restrict__ affine, int64_t P, int num_tiles,
                  float* __restrict__ out, float* __restrict__ act, uint8_t* __restrict__ ws) {
    using C = Cfg<KIND>;
    using G = Geo<KIND>;
    using W = Ws<KIND>;
    constexpr int H = C::H, E = C::E, NF = C::NF, NOUT = C::NOUT, HS = G::HS, ES = G::ES, AS = G::AS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sAhi = smem;                          // AS slabs: [0,HS) hidden, [HS,AS) embedding
    uint8_t* sAlo = sAhi + AS * SLAB;
    uint8_t* sC = sAlo + AS * SLAB;                // cols 0..31 c hi, 32..63 c lo (fp16 both)
    uint8_t* sW = sC + SLAB;                       // NST ring stages
    float* sBias = reinterpret_cast<float*>(sW + NST * SLAB);      // [5][2][H]: b_i, bc_i
    float* sB = sBias + 2 * 5 * H;                                  // [3][NF] Fourier projection
    float* sBo = sB + 3 * NF;                                       // [4] output bias
    float* sAff = sBo + 4;                                          // [12] exposure affine (colour)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sAff + 12 + ((3 * NF) & 1));   // 8-byte aligned
    uint64_t* full = bars;                         // [NST]
    uint64_t* empty = bars + NST;                  // [NST]
    uint64_t* a_ready = bars + 2 * NST;            // epilogue -> MMA: A operand in place, accumulators drained
    uint64_t* acc_ready = bars + 2 * NST + 1;      // MMA -> epilogue
    uint64_t* dump_done = bars + 2 * NST + 2;      // dump warp -> epilogue: operand slabs copied out, reusable
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup ----
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
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
                     "r"((uint32_t)G::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = threadIdx.x; i < 5 * H; i += NTHR) {
        const int l = i / H, n = i % H;
        sBias[(l * 2 + 0) * H + n] = prm.v[iB(l)][n];
        sBias[(l * 2 + 1) * H + n] = prm.v[iBc(l)][n];
    }
    for (int i = threadIdx.x; i < 3 * NF; i += NTHR) sB[i] = prm.v[0][i];
    if (threadIdx.x < NOUT) sBo[threadIdx.x] = prm.v[iBo][threadIdx.x];
    if (threadIdx.x < 12 && affine != nullptr) sAff[threadIdx.x] = affine[threadIdx.x];
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
        unsigned long long* tr = (blockIdx.x == 0 && lane == 0) ? g_trace_fwd : nullptr;
        int ts = 64;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
#pragma unroll 1
            for (int layer = 0; layer < 6; ++layer) {
                mbar_wait(a_ready, nready & 1);
                ++nready;
                tc_fence_after();
                trace_stamp(tr, ts++);
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
                trace_stamp(tr, ts++);
            }
        }
    } else if (warp == DUMP_WARP) {
        // ================= dump warp (DUMP == 2): operand slabs -> workspace of the tensor-core backward ==============
        if constexpr (DUMP == 2) {
            uint32_t nready = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                uint8_t* xg = W::x_ptr(ws, tile);
#pragma unroll 1
                for (int step = 0; step < 6; ++step) {     // 0: embedding + c, 1..5: y_0 .. y_4
                    mbar_wait(a_ready, nready & 1);
                    ++nready;
                    if (elect_one()) {
                        if (step == 0) {
                            for (int s2 = 0; s2 < ES; ++s2) {
                                bulk_s2g(xg + (size_t)W::x_e(0, s2) * SLAB, sAhi + (HS + s2) * SLAB, SLAB);
                                bulk_s2g(xg + (size_t)W::x_e(1, s2) * SLAB, sAlo + (HS + s2) * SLAB, SLAB);
                            }
                            bulk_s2g(xg + (size_t)W::x_c() * SLAB, sC, SLAB);
                        } else {
                            for (int s2 = 0; s2 < HS; ++s2) {
                                bulk_s2g(xg + (size_t)W::x_y(step - 1, 0, s2) * SLAB, sAhi + s2 * SLAB, SLAB);
                                bulk_s2g(xg + (size_t)W::x_y(step - 1, 1, s2) * SLAB, sAlo + s2 * SLAB, SLAB);
                            }
                        }
                        bulk_commit();
                        bulk_wait_read();
                        mbar_arrive(dump_done);
                    }
                    __syncwarp();
                }
            }
            bulk_wait_all();
        }
    } else {
        // ================= epilogue warps =================
        const int ew = warp - EPI_WARP0;          // 0..EPI_WARPS-1
        uint32_t ndump = 0;
        const int quad = warp & 3;                // TMEM lane quadrant this warp may access
        const int cgp = ew >> 2;                  // column group handled by this warp
        const int row = quad * 32 + lane;         // tile row = TMEM lane
        constexpr int CPT = H / CG;               // hidden columns per thread: colour 32, geometry 8
        static_assert(CPT == 32 || CPT == 8, "epilogue column split");
        constexpr int NCH = (E + 7) / 8;          // 16-byte chunks (8 columns) of the embedding
        const int col0 = cgp * CPT;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        uint32_t nacc = 0;
        unsigned long long* tr = (blockIdx.x == 0 && ew == 0 && lane == 0) ? g_trace_fwd : nullptr;
        int ts = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            trace_stamp(tr, ts++);
            const int64_t pt = (int64_t)tile * TILE + row;
            const bool live = pt < P;
            float* ag = DUMP == 2 ? W::a_ptr(ws, tile) : nullptr;
            // ---- tile prologue: c -> c slab, Fourier features -> embedding slabs ----
            {
                // this thread converts kC / CG = 8 of the 32 channels; loaded first so the latency hides behind the sines
                static_assert(kC / CG == 8, "c split");
                float cf[8];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float4 v4 = live ? ldg4(c + pt * kC + cgp * 8 + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    cf[q * 4] = v4.x; cf[q * 4 + 1] = v4.y; cf[q * 4 + 2] = v4.z; cf[q * 4 + 3] = v4.w;
                }
                float x0 = 0.f, x1 = 0.f, x2 = 0.f;
                if (live) {
                    x0 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 0));
                    x1 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 1));
                    x2 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 2));
                }
                if constexpr (DUMP == 2) {      // the previous tile's y_4 dump also covers its embedding / c slabs
                    mbar_wait(dump_done, (ndump & 1) ^ 1);
                    ++ndump;
                }
#pragma unroll 1
                for (int ch = cgp; ch < NCH; ch += CG) {        // ch is warp-uniform
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int e = ch * 8 + j;
                        float val = 0.f;                         // columns >= E stay zero
                        if (e < E) {
                            const bool is_cos = C::CONCAT && e >= NF;
                            const int f = is_cos ? e - NF : e;
                            const float arg = fmaf(x2, sB[2 * NF + f], fmaf(x1, sB[NF + f], x0 * sB[f]));
                            val = is_cos ? cos_2pi(arg) : sin_2pi(arg);
                        }
                        v[j] = val;
                    }
                    uint4 vh, vl;
                    split8(v, vh, vl);
                    const int slab = ch >> 3;
                    const uint32_t so = slab_off(row, ch & 7);
                    *reinterpret_cast<uint4*>(sAhi + (HS + slab) * SLAB + so) = vh;
                    *reinterpret_cast<uint4*>(sAlo + (HS + slab) * SLAB + so) = vl;
                }
                trace_stamp(tr, ts++);
                uint4 vh, vl;
                split8(cf, vh, vl);
                *reinterpret_cast<uint4*>(sC + slab_off(row, cgp)) = vh;
                *reinterpret_cast<uint4*>(sC + slab_off(row, 4 + cgp)) = vl;
            }
            fence_proxy_async();
            mbar_arrive(a_ready);
            trace_stamp(tr, ts++);

            // SIMT-backward activation workspace (DUMP == 1): [64-point tile][slot][feature][point]; rows beyond P
            // but inside the last 64-point tile are written too (finite values: the backward multiplies them by
            // zero gradients)
            float* act_tile0 = (DUMP == 1 && pt < ((P + 63) / 64) * 64)
                                   ? act + (pt / 64) * (int64_t)(10 * H * 64) + (pt % 64) : nullptr;
            for (int layer = 0; layer < 5; ++layer) {
                mbar_wait(acc_ready, nacc & 1);
                ++nacc;
                tc_fence_after();
                trace_stamp(tr, ts++);
                float a[CPT], u[CPT];
                if constexpr (CPT == 32) {
                    tc_ld32(tmem + lane_addr + col0, a);
                    tc_ld32(tmem + lane_addr + H + col0, u);
                } else {
                    tc_ld8(tmem + lane_addr + col0, a);
                    tc_ld8(tmem + lane_addr + H + col0, u);
                }
                if constexpr (DUMP == 2) {
                    mbar_wait(dump_done, (ndump & 1) ^ 1);
                    ++ndump;
                }
                const float4* bi4 = reinterpret_cast<const float4*>(sBias + (layer * 2 + 0) * H + col0);
                const float4* bc4 = reinterpret_cast<const float4*>(sBias + (layer * 2 + 1) * H + col0);
#pragma unroll
                for (int q = 0; q < CPT / 8; ++q) {
                    const float4 b0 = bi4[2 * q], b1 = bi4[2 * q + 1], c0 = bc4[2 * q], c1 = bc4[2 * q + 1];
                    const float bi[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    const float bc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                    float yv[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float av = a[q * 8 + j] + bi[j];
                        yv[j] = act_fast<C::RELU>(av) + (u[q * 8 + j] + bc[j]);
                        const int n = col0 + q * 8 + j;
                        if (DUMP == 1 && act_tile0 != nullptr) {
                            act_tile0[(int64_t)(2 * layer) * H * 64 + n * 64] = av;
                            act_tile0[(int64_t)(2 * layer + 1) * H * 64 + n * 64] = yv[j];
                        }
                        if (DUMP == 2) ag[(layer * H + n) * TILE + row] = av;
                    }
                    uint4 vh, vl;
                    split8(yv, vh, vl);
                    const int colabs = col0 + q * 8;
                    const int slab = colabs >> 6;
                    const uint32_t so = slab_off(row, (colabs & 63) >> 3);
                    *reinterpret_cast<uint4*>(sAhi + slab * SLAB + so) = vh;
                    *reinterpret_cast<uint4*>(sAlo + slab * SLAB + so) = vl;
                }
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(a_ready);
                trace_stamp(tr, ts++);
            }
            // ---- output layer result ----
            mbar_wait(acc_ready, nacc & 1);
            ++nacc;
            tc_fence_after();
            trace_stamp(tr, ts++);
            if (cgp == 0) {
                float o[8];
                tc_ld8(tmem + lane_addr + 2 * H, o);
                if (DUMP == 2) {
                    float* rg = W::raw_ptr(ws, tile) + row * 4;
#pragma unroll
                    for (int q = 0; q < NOUT; ++q) rg[q] = o[q] + sBo[q];
                }
                if (live) {
                    if (NOUT == 1) {
                        out[pt] = o[0] + sBo[0];
                    } else {
                        float r[3] = {o[0] + sBo[0], o[1] + sBo[1], o[2] + sBo[2]};
                        if (affine != nullptr) {
                            float q[3];
#pragma unroll
                            for (int cc = 0; cc < 3; ++cc) {
                                const float lin = r[0] * sAff[cc] + r[1] * sAff[3 + cc] + r[2] * sAff[6 + cc] + sAff[9 + cc];
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
            trace_stamp(tr, ts++);
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
               int64_t P, float* out, float* act, void* ws, int max_ctas, cudaStream_t s) {
    Ptrs prm;
    for (int i = 0; i < HPS_MLP_NPARAM; ++i) prm.v[i] = params[i];
    const int smem = Geo<KIND>::SMEM_BYTES;
    static bool configured = false;
    if (!configured) {
        HPS_CUDA_OK(cudaFuncSetAttribute(mlp_fwd_tc_kernel<KIND, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        HPS_CUDA_OK(cudaFuncSetAttribute(mlp_fwd_tc_kernel<KIND, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        HPS_CUDA_OK(cudaFuncSetAttribute(mlp_fwd_tc_kernel<KIND, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const int num_tiles = (int)ceil_div(P, TILE);
    const int cap = (max_ctas > 0 && max_ctas < kNumSMs) ? max_ctas : kNumSMs;
    const int grid = num_tiles < cap ? num_tiles : cap;
    auto kern = ws ? mlp_fwd_tc_kernel<KIND, 2> : (act ? mlp_fwd_tc_kernel<KIND, 1> : mlp_fwd_tc_kernel<KIND, 0>);
    kern<<<grid, NTHR, smem, s>>>(prm, reinterpret_cast<const uint8_t*>(img), p, c, affine, P, num_tiles, out, act,
                                  reinterpret_cast<uint8_t*>(ws));
    return 0;
}

}  // namespace tc
}  // namespace hps

extern "C" int hps_tc_trace(void* buf) {
    using namespace hps;
    int rc = tc::set_trace_fwd(buf);
    if (rc) return rc;
    return tc::set_trace_bwd(buf);
}

extern "C" int64_t hps_mlp_tc_ws_bytes(int kind, int64_t P) {
    using namespace hps::tc;
    const int64_t tiles = (P + TILE - 1) / TILE;
    return tiles * (int64_t)(kind == HPS_MLP_GEO ? Ws<HPS_MLP_GEO>::TILE_BYTES : Ws<HPS_MLP_COL>::TILE_BYTES);
}

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
                              const float* c, const float* affine, int64_t P, float* out, float* act, void* ws,
                              int max_ctas, void* stream) {
    using namespace hps;
    HPS_REQUIRE(kind == HPS_MLP_GEO || kind == HPS_MLP_COL, "bad decoder kind");
    HPS_REQUIRE(P >= 0 && P < (1ll << 31) * 64, "point count out of range");
    if (P == 0) return 0;
    HPS_REQUIRE(params && packed && p && c && out, "null buffer");
    HPS_REQUIRE(kind == HPS_MLP_COL || affine == nullptr, "affine is colour-only");
    for (int i = 0; i < HPS_MLP_NPARAM; ++i) HPS_REQUIRE(params[i] != nullptr, "null parameter pointer");
    int rc = kind == HPS_MLP_GEO
                 ? tc::launch_fwd<HPS_MLP_GEO>(params, packed, p, c, affine, P, out, act, ws, max_ctas, (cudaStream_t)stream)
                 : tc::launch_fwd<HPS_MLP_COL>(params, packed, p, c, affine, P, out, act, ws, max_ctas, (cudaStream_t)stream);
    if (rc) return rc;
    HPS_LAUNCH_OK();
    return 0;
}
