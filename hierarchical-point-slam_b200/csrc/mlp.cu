// Decoder MLPs, forward and backward, fp32 SIMT tiles (kernel 3 of the north star, first edition).
//
// Restates MLP_geometry.forward (reference src/conv_onet/models/decoder.py:316-341) and
// MLP_color.forward (:581-618):
//     e   = sin(2 pi p B)            (geo, B (3,93) learnable)   |  [sin, cos](2 pi p B)  (colour, B (3,20))
//     x_0 = e;  a_i = W_i x_i + b_i;  y_i = act(a_i) + Wc_i c + bc_i;  x_{i+1} = y_i,  x_3 = [e, y_2]
//     out = Wo y_4 + bo              (geo: ReLU, hidden 32, 1 output; colour: Softplus(100), hidden 128, 3)
//     colour tail with exposure: sigmoid(out @ A[:9].view(3,3) + A[9:])      (:606-612)
//
// Tiling: one CTA = 64 sample points, 256 threads.  Activations live in shared memory contraction-major
// ([feature][point]) for the whole 5-block chain, so HBM sees p, c in and out (+ the saved a_i, y_i
// when a backward will follow).  Each layer is a register-tiled GEMM (4 points x RN outputs per thread)
// against k-major weight chunks staged in shared memory from a pre-transposed copy (hps_mlp_pack).
// The backward walks the chain in reverse inside one CTA: dgrad with the same tile routine against the
// weights in their native (out,in) layout, wgrad as a dot-product-style tile over the 64 points
// (rows padded by 4 floats -> conflict-free 128-bit reads), parameter gradients reduced across CTAs
// with red.global.add.  fp32 throughout: this edition is the parity anchor for the tcgen05 edition.
#include "hps_common.cuh"

namespace hps {

constexpr int TM = 64;      // points per CTA tile
constexpr int TMP = TM + 4; // padded row stride used by the backward
constexpr int NT = 256;     // threads per CTA
constexpr int KC = 32;      // contraction rows staged per chunk (forward)
constexpr int KCB = 16;     // contraction rows staged per chunk (backward dgrad)
constexpr int RM = 4;       // points per thread
constexpr int MG = TM / RM; // 16 point-groups
constexpr float kTwoPi = 6.283185307179586f;

template <int KIND> struct MlpCfg;
template <> struct MlpCfg<HPS_MLP_GEO> {
    static constexpr int H = 32, NF = 93, E = 93, NOUT = 1, RN = 2;
    static constexpr bool CONCAT = false, RELU = true;
};
template <> struct MlpCfg<HPS_MLP_COL> {
    static constexpr int H = 128, NF = 20, E = 40, NOUT = 3, RN = 8;
    static constexpr bool CONCAT = true, RELU = false;
};

template <int KIND> __host__ __device__ constexpr int layer_k(int i) {
    return i == 0 ? MlpCfg<KIND>::E : (i == 3 ? MlpCfg<KIND>::E + MlpCfg<KIND>::H : MlpCfg<KIND>::H);
}
// packed (transposed) weights: Wt_0..Wt_4 [K_i][H], Wct_0..Wct_4 [32][H]
template <int KIND> __host__ __device__ constexpr int packed_w_off(int i) {
    int o = 0;
    for (int j = 0; j < i; ++j) o += layer_k<KIND>(j) * MlpCfg<KIND>::H;
    return o;
}
template <int KIND> __host__ __device__ constexpr int packed_wc_off(int i) {
    return packed_w_off<KIND>(5) + i * kC * MlpCfg<KIND>::H;
}
template <int KIND> __host__ __device__ constexpr int packed_total() { return packed_wc_off<KIND>(5); }

struct MlpPtrs {
    const float* v[HPS_MLP_NPARAM];
};
struct MlpGradPtrs {
    float* v[HPS_MLP_NPARAM];
};
__host__ __device__ constexpr int iW(int i) { return 1 + 4 * i; }
__host__ __device__ constexpr int iB(int i) { return 2 + 4 * i; }
__host__ __device__ constexpr int iWc(int i) { return 3 + 4 * i; }
__host__ __device__ constexpr int iBc(int i) { return 4 + 4 * i; }
constexpr int iWo = 21, iBo = 22;

template <bool RELU> __device__ __forceinline__ float act_fwd(float a) {
    if (RELU) return fmaxf(a, 0.f);
    const float t = a * 100.0f;
    return t > 20.0f ? a : log1pf(expf(t)) / 100.0f;
}
template <bool RELU> __device__ __forceinline__ float act_bwd(float a) {
    if (RELU) return a > 0.f ? 1.f : 0.f;
    const float t = a * 100.0f;
    if (t > 20.0f) return 1.f;
    const float z = expf(t);
    return z / (z + 1.0f);
}

__device__ __forceinline__ float embed_arg(const float* __restrict__ B, int NF, int j, float x0, float x1, float x2) {
    return fmaf(x2, __ldg(B + 2 * NF + j), fmaf(x1, __ldg(B + NF + j), x0 * __ldg(B + j)));
}

// ------------------------------------------------------------------------------------------------
// pack: W (H,K) row-major -> Wt (K,H)
// ------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void mlp_pack_kernel(MlpPtrs prm, float* __restrict__ packed) {
    using C = MlpCfg<KIND>;
    const int total = packed_total<KIND>();
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int i = 0, off = idx;
        const float* src = nullptr;
        int K = 0;
        if (idx < packed_w_off<KIND>(5)) {
            while (i < 4 && idx >= packed_w_off<KIND>(i + 1)) ++i;
            off = idx - packed_w_off<KIND>(i);
            K = layer_k<KIND>(i);
            src = prm.v[iW(i)];
        } else {
            i = (idx - packed_w_off<KIND>(5)) / (kC * C::H);
            off = idx - packed_wc_off<KIND>(i);
            K = kC;
            src = prm.v[iWc(i)];
        }
        const int k = off / C::H, n = off % C::H;
        packed[idx] = src[n * K + k];
    }
}

// ------------------------------------------------------------------------------------------------
// register-tiled contraction:  acc[rm][rn] += sum_kc A[kc][m0+rm] * B[kc][n0+rn]
// ------------------------------------------------------------------------------------------------
template <int RN>
__device__ __forceinline__ void tile_fma(const float* __restrict__ sA, int lda, const float* __restrict__ sB,
                                         int ldb, int kc, int m0, int n0, float (&acc)[RM][RN]) {
#pragma unroll 4
    for (int k = 0; k < kc; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(sA + k * lda + m0);
        const float a[RM] = {a4.x, a4.y, a4.z, a4.w};
        float b[RN];
        if constexpr (RN == 8) {
            const float4 b0 = *reinterpret_cast<const float4*>(sB + k * ldb + n0);
            const float4 b1 = *reinterpret_cast<const float4*>(sB + k * ldb + n0 + 4);
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
        } else {
            static_assert(RN == 2 || RN == 8, "tile width");
            const float2 b0 = *reinterpret_cast<const float2*>(sB + k * ldb + n0);
            b[0] = b0.x; b[1] = b0.y;
        }
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
            for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
}

template <int KIND>
struct FwdSmem {
    using C = MlpCfg<KIND>;
    static constexpr int ROWS = C::E + C::H;
    static constexpr int kBuf = ROWS * TM;
    static constexpr int kC_ = kC * TM;
    static constexpr int kW = KC * C::H;
    static constexpr int kOut = TM * 4;
    static constexpr int floats = 2 * kBuf + kC_ + kW + kOut;
};

template <int KIND>
__global__ void __launch_bounds__(NT, 2)
mlp_fwd_kernel(MlpPtrs prm, const float* __restrict__ packed, const float* __restrict__ p,
               const float* __restrict__ c, const float* __restrict__ affine, int64_t P,
               float* __restrict__ out, float* __restrict__ act) {
    using C = MlpCfg<KIND>;
    constexpr int H = C::H, E = C::E, NF = C::NF, RN = C::RN, NOUT = C::NOUT;
    extern __shared__ float4 smem4[];
    float* buf0 = reinterpret_cast<float*>(smem4);
    float* buf1 = buf0 + FwdSmem<KIND>::kBuf;
    float* sC = buf1 + FwdSmem<KIND>::kBuf;
    float* sW = sC + FwdSmem<KIND>::kC_;
    float* sOut = sW + FwdSmem<KIND>::kW;
    const int tid = threadIdx.x;
    const int64_t m_base = (int64_t)blockIdx.x * TM;
    const int mg = tid % MG, ng = tid / MG;
    const int m0 = mg * RM, n0 = ng * RN;

    // c tile -> sC[k][m]
    for (int idx = tid; idx < TM * (kC / 4); idx += NT) {
        const int m = idx / (kC / 4), q = idx % (kC / 4);
        const int64_t pt = m_base + m;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pt < P) v = ldg4(c + pt * kC + q * 4);
        sC[(q * 4 + 0) * TM + m] = v.x;
        sC[(q * 4 + 1) * TM + m] = v.y;
        sC[(q * 4 + 2) * TM + m] = v.z;
        sC[(q * 4 + 3) * TM + m] = v.w;
    }
    // Fourier features -> rows [0,E) of both ping-pong buffers
    for (int idx = tid; idx < NF * TM; idx += NT) {
        const int j = idx / TM, m = idx % TM;
        const int64_t pt = m_base + m;
        float x0 = 0.f, x1 = 0.f, x2 = 0.f;
        if (pt < P) {
            x0 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 0));
            x1 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 1));
            x2 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 2));
        }
        const float arg = embed_arg(prm.v[0], NF, j, x0, x1, x2);
        if (C::CONCAT) {
            float s, co;
            sincosf(arg, &s, &co);
            buf0[j * TM + m] = s;
            buf1[j * TM + m] = s;
            buf0[(NF + j) * TM + m] = co;
            buf1[(NF + j) * TM + m] = co;
        } else {
            const float s = sinf(arg);
            buf0[j * TM + m] = s;
            buf1[j * TM + m] = s;
        }
    }
    __syncthreads();

    float* cur = buf0;
    float* nxt = buf1;
    float* act_tile = act ? act + (int64_t)blockIdx.x * (10 * H * TM) : nullptr;
#pragma unroll 1
    for (int i = 0; i < 5; ++i) {
        const int K = layer_k<KIND>(i);
        const float* xin = (i == 0 || i == 3) ? cur : cur + E * TM;
        float acc_u[RM][RN], acc_a[RM][RN];
#pragma unroll
        for (int a = 0; a < RM; ++a)
#pragma unroll
            for (int b = 0; b < RN; ++b) acc_u[a][b] = acc_a[a][b] = 0.f;
        // feature injection  u = Wc_i c
        {
            const float4* src = reinterpret_cast<const float4*>(packed + packed_wc_off<KIND>(i));
            for (int idx = tid; idx < kC * H / 4; idx += NT) reinterpret_cast<float4*>(sW)[idx] = __ldg(src + idx);
            __syncthreads();
            tile_fma<RN>(sC, TM, sW, H, kC, m0, n0, acc_u);
            __syncthreads();
        }
        // a = W_i x
        const float* wt = packed + packed_w_off<KIND>(i);
        for (int k0 = 0; k0 < K; k0 += KC) {
            const int kc = min(KC, K - k0);
            const float4* src = reinterpret_cast<const float4*>(wt + k0 * H);
            for (int idx = tid; idx < kc * H / 4; idx += NT) reinterpret_cast<float4*>(sW)[idx] = __ldg(src + idx);
            __syncthreads();
            tile_fma<RN>(xin + k0 * TM, TM, sW, H, kc, m0, n0, acc_a);
            __syncthreads();
        }
        // epilogue
#pragma unroll
        for (int j = 0; j < RN; ++j) {
            const int n = n0 + j;
            const float bi = __ldg(prm.v[iB(i)] + n), bci = __ldg(prm.v[iBc(i)] + n);
            float av[RM], yv[RM];
#pragma unroll
            for (int r = 0; r < RM; ++r) {
                av[r] = acc_a[r][j] + bi;
                yv[r] = act_fwd<C::RELU>(av[r]) + (acc_u[r][j] + bci);
            }
            *reinterpret_cast<float4*>(nxt + (E + n) * TM + m0) = make_float4(yv[0], yv[1], yv[2], yv[3]);
            if (act_tile) {
                *reinterpret_cast<float4*>(act_tile + (2 * i) * H * TM + n * TM + m0) =
                    make_float4(av[0], av[1], av[2], av[3]);
                *reinterpret_cast<float4*>(act_tile + (2 * i + 1) * H * TM + n * TM + m0) =
                    make_float4(yv[0], yv[1], yv[2], yv[3]);
            }
        }
        __syncthreads();
        float* t = cur; cur = nxt; nxt = t;
    }
    // output layer
    for (int idx = tid; idx < TM * NOUT; idx += NT) {
        const int m = idx % TM, o = idx / TM;
        const float* wo = prm.v[iWo] + o * H;
        float s = 0.f;
        for (int n = 0; n < H; ++n) s = fmaf(cur[(E + n) * TM + m], __ldg(wo + n), s);
        sOut[m * 4 + o] = s + __ldg(prm.v[iBo] + o);
    }
    __syncthreads();
    if (tid < TM) {
        const int64_t pt = m_base + tid;
        if (pt < P) {
            if (NOUT == 1) {
                out[pt] = sOut[tid * 4];
            } else {
                float r[3] = {sOut[tid * 4 + 0], sOut[tid * 4 + 1], sOut[tid * 4 + 2]};
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
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <int KIND>
struct BwdSmem {
    using C = MlpCfg<KIND>;
    static constexpr int NCOL = (NT / MG) * C::RN;          // columns produced per dgrad pass
    static constexpr int kG = C::H * TMP;
    static constexpr int kDE = C::E * TMP;
    static constexpr int kX = (C::E + C::H) * TMP;
    static constexpr int kC_ = kC * TMP;
    static constexpr int kW = KCB * NCOL > KCB * kC ? KCB * NCOL : KCB * kC;
    static constexpr int kMisc = 4 * TMP + TM * 4 + 16;              // sDraw[4][TMP], sRaw[TM][4], scratch
    static constexpr int floats = 2 * kG + kDE + kX + kC_ + kW + kMisc;
};

// dW[n][k] += sum_m sA[n][m] * sX[k][m]   (thread tile: rows {nq + i*NGR}, cols k0..k0+RK-1)
template <int RNW, int RKW>
__device__ __forceinline__ void wgrad_tile(const float* __restrict__ sA, int N, const float* __restrict__ sX,
                                           int K, float* __restrict__ gW, int tid) {
    const int ngr = N / RNW, kgr = (K + RKW - 1) / RKW;
    for (int t = tid; t < ngr * kgr; t += NT) {
        const int nq = t % ngr, k0 = (t / ngr) * RKW;
        float acc[RNW][RKW];
#pragma unroll
        for (int i = 0; i < RNW; ++i)
#pragma unroll
            for (int j = 0; j < RKW; ++j) acc[i][j] = 0.f;
#pragma unroll 2
        for (int m = 0; m < TM; m += 4) {
            float4 a[RNW], b[RKW];
#pragma unroll
            for (int i = 0; i < RNW; ++i) a[i] = *reinterpret_cast<const float4*>(sA + (nq + i * ngr) * TMP + m);
#pragma unroll
            for (int j = 0; j < RKW; ++j) {
                const int k = min(k0 + j, K - 1);
                b[j] = *reinterpret_cast<const float4*>(sX + k * TMP + m);
            }
#pragma unroll
            for (int i = 0; i < RNW; ++i)
#pragma unroll
                for (int j = 0; j < RKW; ++j)
                    acc[i][j] += a[i].x * b[j].x + a[i].y * b[j].y + a[i].z * b[j].z + a[i].w * b[j].w;
        }
#pragma unroll
        for (int i = 0; i < RNW; ++i) {
            float* row = gW + (int64_t)(nq + i * ngr) * K + k0;
            static_assert(RKW == 4, "wgrad tile width");
            if ((K % 4) == 0) {
                red_add_v4(row, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
            } else {
#pragma unroll
                for (int j = 0; j < RKW; ++j)
                    if (k0 + j < K) red_add(row + j, acc[i][j]);
            }
        }
    }
}

// gdst[n] += sum_m sG[n][m]   (one warp per row)
__device__ __forceinline__ void rowsum_atomic(const float* __restrict__ sG, int N, float* __restrict__ gdst, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    for (int n = warp; n < N; n += NT / 32) {
        float s = sG[n * TMP + lane] + sG[n * TMP + lane + 32];
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
        if (lane == 0) red_add(gdst + n, s);
    }
}

template <int KIND>
__global__ void __launch_bounds__(NT)
mlp_bwd_kernel(MlpPtrs prm, MlpGradPtrs grd, const float* __restrict__ p, const float* __restrict__ c,
               const float* __restrict__ affine, const float* __restrict__ act,
               const float* __restrict__ dout, int64_t P, float* __restrict__ dc, float* __restrict__ dp,
               float* __restrict__ daffine) {
    using C = MlpCfg<KIND>;
    using S = BwdSmem<KIND>;
    constexpr int H = C::H, E = C::E, NF = C::NF, RN = C::RN, NOUT = C::NOUT, NCOL = S::NCOL;
    extern __shared__ float4 smem4[];
    float* G0 = reinterpret_cast<float*>(smem4);
    float* G1 = G0 + S::kG;
    float* sDE = G1 + S::kG;
    float* sX = sDE + S::kDE;
    float* sC = sX + S::kX;
    float* sW = sC + S::kC_;
    float* sDraw = sW + S::kW;          // [4][TMP]
    float* sRaw = sDraw + 4 * TMP;      // [TM][4]
    float* sRed = sRaw + TM * 4;        // [16]
    const int tid = threadIdx.x;
    const int64_t m_base = (int64_t)blockIdx.x * TM;
    const int mg = tid % MG, ng = tid / MG;
    const int m0 = mg * RM;
    const float* act_tile = act + (int64_t)blockIdx.x * (10 * H * TM);
    const bool want_w = grd.v[iWo] != nullptr;

    // ---- stage c, e, y_4; zero dE ----
    for (int idx = tid; idx < TM * (kC / 4); idx += NT) {
        const int m = idx / (kC / 4), q = idx % (kC / 4);
        const int64_t pt = m_base + m;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pt < P) v = ldg4(c + pt * kC + q * 4);
        sC[(q * 4 + 0) * TMP + m] = v.x;
        sC[(q * 4 + 1) * TMP + m] = v.y;
        sC[(q * 4 + 2) * TMP + m] = v.z;
        sC[(q * 4 + 3) * TMP + m] = v.w;
    }
    for (int idx = tid; idx < NF * TM; idx += NT) {
        const int j = idx / TM, m = idx % TM;
        const int64_t pt = m_base + m;
        float x0 = 0.f, x1 = 0.f, x2 = 0.f;
        if (pt < P) {
            x0 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 0));
            x1 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 1));
            x2 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 2));
        }
        const float arg = embed_arg(prm.v[0], NF, j, x0, x1, x2);
        if (C::CONCAT) {
            float s, co;
            sincosf(arg, &s, &co);
            sX[j * TMP + m] = s;
            sX[(NF + j) * TMP + m] = co;
        } else {
            sX[j * TMP + m] = sinf(arg);
        }
    }
    for (int idx = tid; idx < E * TMP; idx += NT) sDE[idx] = 0.f;
    for (int idx = tid; idx < H * (TM / 4); idx += NT) {   // y_4 -> hidden rows of sX
        const int n = idx / (TM / 4), q = idx % (TM / 4);
        *reinterpret_cast<float4*>(sX + (E + n) * TMP + q * 4) = ldg4(act_tile + 9 * H * TM + n * TM + q * 4);
    }
    if (tid < 16) sRed[tid] = 0.f;
    __syncthreads();

    // ---- output layer: d raw ----
    if (NOUT == 3 && affine != nullptr) {
        for (int idx = tid; idx < TM * NOUT; idx += NT) {
            const int m = idx % TM, o = idx / TM;
            const float* wo = prm.v[iWo] + o * H;
            float s = 0.f;
            for (int n = 0; n < H; ++n) s = fmaf(sX[(E + n) * TMP + m], __ldg(wo + n), s);
            sRaw[m * 4 + o] = s + __ldg(prm.v[iBo] + o);
        }
        __syncthreads();
    }
    if (tid < TM) {
        const int64_t pt = m_base + tid;
        float g[3] = {0.f, 0.f, 0.f};
        if (pt < P) {
#pragma unroll
            for (int o = 0; o < NOUT; ++o) g[o] = __ldg(dout + pt * NOUT + o);
        }
        if (NOUT == 3 && affine != nullptr) {
            const float r[3] = {sRaw[tid * 4 + 0], sRaw[tid * 4 + 1], sRaw[tid * 4 + 2]};
            float dl[3];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                const float lin = r[0] * __ldg(affine + cc) + r[1] * __ldg(affine + 3 + cc) +
                                  r[2] * __ldg(affine + 6 + cc) + __ldg(affine + 9 + cc);
                const float s = 1.0f / (1.0f + expf(-lin));
                dl[cc] = g[cc] * s * (1.0f - s);
            }
            if (daffine != nullptr) {
#pragma unroll
                for (int k = 0; k < 3; ++k)
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) atomicAdd(&sRed[k * 3 + cc], r[k] * dl[cc]);
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) atomicAdd(&sRed[9 + cc], dl[cc]);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k)
                g[k] = dl[0] * __ldg(affine + 3 * k + 0) + dl[1] * __ldg(affine + 3 * k + 1) +
                       dl[2] * __ldg(affine + 3 * k + 2);
        }
#pragma unroll
        for (int o = 0; o < 3; ++o) sDraw[o * TMP + tid] = g[o];
    }
    __syncthreads();
    if (NOUT == 3 && affine != nullptr && daffine != nullptr && tid < 12) red_add(daffine + tid, sRed[tid]);

    // dWo, dbo, dy_4
    if (want_w) {
        for (int idx = tid; idx < NOUT * H; idx += NT) {
            const int o = idx / H, n = idx % H;
            float s = 0.f;
            for (int m = 0; m < TM; ++m) s = fmaf(sDraw[o * TMP + m], sX[(E + n) * TMP + m], s);
            red_add(grd.v[iWo] + o * H + n, s);
        }
        if (tid < NOUT) {
            float s = 0.f;
            for (int m = 0; m < TM; ++m) s += sDraw[tid * TMP + m];
            red_add(grd.v[iBo] + tid, s);
        }
    }
    for (int idx = tid; idx < H * TM; idx += NT) {
        const int n = idx / TM, m = idx % TM;
        float s = 0.f;
#pragma unroll
        for (int o = 0; o < NOUT; ++o) s = fmaf(sDraw[o * TMP + m], __ldg(prm.v[iWo] + o * H + n), s);
        G0[n * TMP + m] = s;
    }
    __syncthreads();

    float* cur = G0;
    float* nxt = G1;
    float dc_acc[RM][2];
#pragma unroll
    for (int a = 0; a < RM; ++a) dc_acc[a][0] = dc_acc[a][1] = 0.f;
    const int kc0 = ng * 2;   // this thread's two dc columns

#pragma unroll 1
    for (int i = 4; i >= 0; --i) {
        const int K = layer_k<KIND>(i);
        // (a) dbc_i, (c) dWc_i from dy_i
        if (want_w) {
            rowsum_atomic(cur, H, grd.v[iBc(i)], tid);
            wgrad_tile<(H >= 128 ? 8 : 4), 4>(cur, H, sC, kC, grd.v[iWc(i)], tid);
        }
        // (b) dc += dy_i . Wc_i        (contraction over n, B = Wc_i rows in native layout)
        for (int nb0 = 0; nb0 < H; nb0 += KCB) {
            for (int idx = tid; idx < KCB * kC / 4; idx += NT)
                reinterpret_cast<float4*>(sW)[idx] = ldg4(prm.v[iWc(i)] + nb0 * kC + idx * 4);
            __syncthreads();
            tile_fma<2>(cur + nb0 * TMP, TMP, sW, kC, KCB, m0, kc0, dc_acc);
            __syncthreads();
        }
        // (d) da_i = dy_i * act'(a_i)   (in place)
        for (int idx = tid; idx < H * (TM / 4); idx += NT) {
            const int n = idx / (TM / 4), q = idx % (TM / 4);
            const float4 a = ldg4(act_tile + (2 * i) * H * TM + n * TM + q * 4);
            float4* gp = reinterpret_cast<float4*>(cur + n * TMP + q * 4);
            float4 g = *gp;
            g.x *= act_bwd<C::RELU>(a.x);
            g.y *= act_bwd<C::RELU>(a.y);
            g.z *= act_bwd<C::RELU>(a.z);
            g.w *= act_bwd<C::RELU>(a.w);
            *gp = g;
        }
        // (f) hidden part of x_i = y_{i-1}
        if (i >= 1 && want_w) {
            for (int idx = tid; idx < H * (TM / 4); idx += NT) {
                const int n = idx / (TM / 4), q = idx % (TM / 4);
                *reinterpret_cast<float4*>(sX + (E + n) * TMP + q * 4) =
                    ldg4(act_tile + (2 * (i - 1) + 1) * H * TM + n * TM + q * 4);
            }
        }
        __syncthreads();
        // (e) db_i, (g) dW_i
        if (want_w) {
            rowsum_atomic(cur, H, grd.v[iB(i)], tid);
            const float* xin = (i == 0 || i == 3) ? sX : sX + E * TMP;
            wgrad_tile<(H >= 128 ? 8 : 4), 4>(cur, H, xin, K, grd.v[iW(i)], tid);
        }
        // (h) dx = da_i . W_i  -> next dy (hidden columns) / dE (embedding columns)
        const int col_h0 = (i == 3) ? E : 0;      // first hidden column of W_i
        const bool has_h = (i != 0), has_e = (i == 0 || i == 3);
        for (int part = 0; part < 2; ++part) {
            if ((part == 0 && !has_h) || (part == 1 && !has_e)) continue;
            const int cbeg = part == 0 ? col_h0 : 0;
            const int cnum = part == 0 ? H : E;
            float* dst = part == 0 ? nxt : sDE;
            for (int c0 = 0; c0 < cnum; c0 += NCOL) {
                const int cw = min(NCOL, cnum - c0);
                float acc[RM][RN];
#pragma unroll
                for (int a = 0; a < RM; ++a)
#pragma unroll
                    for (int b = 0; b < RN; ++b) acc[a][b] = 0.f;
                for (int nb0 = 0; nb0 < H; nb0 += KCB) {
                    for (int idx = tid; idx < KCB * NCOL; idx += NT) {
                        const int r = idx / NCOL, cc = idx % NCOL;
                        sW[idx] = cc < cw ? __ldg(prm.v[iW(i)] + (int64_t)(nb0 + r) * K + cbeg + c0 + cc) : 0.f;
                    }
                    __syncthreads();
                    tile_fma<RN>(cur + nb0 * TMP, TMP, sW, NCOL, KCB, m0, ng * RN, acc);
                    __syncthreads();
                }
#pragma unroll
                for (int j = 0; j < RN; ++j) {
                    const int col = c0 + ng * RN + j;
                    if (col < cnum) {
                        float4* d4 = reinterpret_cast<float4*>(dst + col * TMP + m0);
                        if (part == 0) {
                            *d4 = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
                        } else {
                            float4 o = *d4;
                            o.x += acc[0][j]; o.y += acc[1][j]; o.z += acc[2][j]; o.w += acc[3][j];
                            *d4 = o;
                        }
                    }
                }
            }
        }
        __syncthreads();
        float* t = cur; cur = nxt; nxt = t;
    }

    // ---- dc ----
#pragma unroll
    for (int r = 0; r < RM; ++r) {
        const int64_t pt = m_base + m0 + r;
        if (pt < P) *reinterpret_cast<float2*>(dc + pt * kC + kc0) = make_float2(dc_acc[r][0], dc_acc[r][1]);
    }
    // ---- dE -> dp, dB ----
    float* sDP = sDraw;   // reuse: [3][TMP]
    __syncthreads();
    for (int idx = tid; idx < 3 * TMP; idx += NT) sDP[idx] = 0.f;
    __syncthreads();
    const bool want_b = (KIND == HPS_MLP_GEO) && grd.v[0] != nullptr;
    if (dp != nullptr || want_b) {
        for (int idx = tid; idx < NF * TM; idx += NT) {
            const int j = idx / TM, m = idx % TM;
            const int64_t pt = m_base + m;
            float x0 = 0.f, x1 = 0.f, x2 = 0.f;
            if (pt < P) {
                x0 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 0));
                x1 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 1));
                x2 = __fmul_rn(kTwoPi, __ldg(p + pt * 3 + 2));
            }
            const float b0 = __ldg(prm.v[0] + j), b1 = __ldg(prm.v[0] + NF + j), b2 = __ldg(prm.v[0] + 2 * NF + j);
            const float arg = fmaf(x2, b2, fmaf(x1, b1, x0 * b0));
            float s, co;
            sincosf(arg, &s, &co);
            float g = sDE[j * TMP + m] * co;
            if (C::CONCAT) g -= sDE[(NF + j) * TMP + m] * s;
            if (pt >= P) g = 0.f;
            if (dp != nullptr) {
                atomicAdd(&sDP[0 * TMP + m], g * kTwoPi * b0);
                atomicAdd(&sDP[1 * TMP + m], g * kTwoPi * b1);
                atomicAdd(&sDP[2 * TMP + m], g * kTwoPi * b2);
            }
            if (want_b) {   // a warp covers 32 consecutive m of one j
                float g0 = g * x0, g1 = g * x1, g2 = g * x2;
                for (int o = 16; o; o >>= 1) {
                    g0 += __shfl_xor_sync(kFull, g0, o);
                    g1 += __shfl_xor_sync(kFull, g1, o);
                    g2 += __shfl_xor_sync(kFull, g2, o);
                }
                if ((tid & 31) == 0) {
                    red_add(grd.v[0] + j, g0);
                    red_add(grd.v[0] + NF + j, g1);
                    red_add(grd.v[0] + 2 * NF + j, g2);
                }
            }
        }
        __syncthreads();
        if (dp != nullptr && tid < TM) {
            const int64_t pt = m_base + tid;
            if (pt < P) {
                dp[pt * 3 + 0] = sDP[0 * TMP + tid];
                dp[pt * 3 + 1] = sDP[1 * TMP + tid];
                dp[pt * 3 + 2] = sDP[2 * TMP + tid];
            }
        }
    }
}

template <int KIND> int launch_pack(const float* const* params, float* packed, cudaStream_t s) {
    MlpPtrs prm;
    for (int i = 0; i < HPS_MLP_NPARAM; ++i) prm.v[i] = params[i];
    mlp_pack_kernel<KIND><<<64, 256, 0, s>>>(prm, packed);
    return 0;
}

template <int KIND>
int launch_fwd(const float* const* params, const float* packed, const float* p, const float* c,
               const float* affine, int64_t P, float* out, float* act, cudaStream_t s) {
    MlpPtrs prm;
    for (int i = 0; i < HPS_MLP_NPARAM; ++i) prm.v[i] = params[i];
    const size_t smem = FwdSmem<KIND>::floats * sizeof(float);
    static bool configured = false;
    if (!configured) {
        HPS_CUDA_OK(cudaFuncSetAttribute(mlp_fwd_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    mlp_fwd_kernel<KIND><<<(unsigned)ceil_div(P, TM), NT, smem, s>>>(prm, packed, p, c, affine, P, out, act);
    return 0;
}

template <int KIND>
int launch_bwd(const float* const* params, const float* p, const float* c, const float* affine,
               const float* act, const float* dout, int64_t P, float* const* dparams, float* dc, float* dp,
               float* daffine, cudaStream_t s) {
    MlpPtrs prm;
    MlpGradPtrs grd;
    for (int i = 0; i < HPS_MLP_NPARAM; ++i) {
        prm.v[i] = params[i];
        grd.v[i] = dparams ? dparams[i] : nullptr;
    }
    if (dparams) {
        // wgrad is all-or-nothing for the Linear layers (B is optional)
        for (int i = 1; i < HPS_MLP_NPARAM; ++i)
            HPS_REQUIRE(grd.v[i] != nullptr, "dparams: all Linear gradients or none");
    }
    const size_t smem = BwdSmem<KIND>::floats * sizeof(float);
    static bool configured = false;
    if (!configured) {
        HPS_CUDA_OK(cudaFuncSetAttribute(mlp_bwd_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    mlp_bwd_kernel<KIND><<<(unsigned)ceil_div(P, TM), NT, smem, s>>>(prm, grd, p, c, affine, act, dout, P, dc, dp,
                                                                    daffine);
    return 0;
}

}  // namespace hps

extern "C" int64_t hps_mlp_act_floats(int kind) {
    return kind == HPS_MLP_GEO ? 10 * hps::MlpCfg<HPS_MLP_GEO>::H : 10 * hps::MlpCfg<HPS_MLP_COL>::H;
}
extern "C" int64_t hps_mlp_packed_floats(int kind) {
    return kind == HPS_MLP_GEO ? hps::packed_total<HPS_MLP_GEO>() : hps::packed_total<HPS_MLP_COL>();
}

extern "C" int hps_mlp_pack(int kind, const float* const* params, float* packed, void* stream) {
    using namespace hps;
    HPS_REQUIRE(kind == HPS_MLP_GEO || kind == HPS_MLP_COL, "bad decoder kind");
    HPS_REQUIRE(params && packed, "null buffer");
    for (int i = 0; i < HPS_MLP_NPARAM; ++i) HPS_REQUIRE(params[i] != nullptr, "null parameter pointer");
    int rc = kind == HPS_MLP_GEO ? launch_pack<HPS_MLP_GEO>(params, packed, (cudaStream_t)stream)
                                 : launch_pack<HPS_MLP_COL>(params, packed, (cudaStream_t)stream);
    if (rc) return rc;
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_mlp_fwd(int kind, const float* const* params, const float* packed, const float* p,
                           const float* c, const float* affine, int64_t P, float* out, float* act,
                           void* stream) {
    using namespace hps;
    HPS_REQUIRE(kind == HPS_MLP_GEO || kind == HPS_MLP_COL, "bad decoder kind");
    HPS_REQUIRE(P >= 0, "negative point count");
    if (P == 0) return 0;
    HPS_REQUIRE(params && packed && p && c && out, "null buffer");
    HPS_REQUIRE(kind == HPS_MLP_COL || affine == nullptr, "affine is colour-only");
    for (int i = 0; i < HPS_MLP_NPARAM; ++i) HPS_REQUIRE(params[i] != nullptr, "null parameter pointer");
    int rc = kind == HPS_MLP_GEO
                 ? launch_fwd<HPS_MLP_GEO>(params, packed, p, c, affine, P, out, act, (cudaStream_t)stream)
                 : launch_fwd<HPS_MLP_COL>(params, packed, p, c, affine, P, out, act, (cudaStream_t)stream);
    if (rc) return rc;
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_mlp_bwd(int kind, const float* const* params, const float* p, const float* c,
                           const float* affine, const float* act, const float* dout, int64_t P,
                           float* const* dparams, float* dc, float* dp, float* daffine, void* stream) {
    using namespace hps;
    HPS_REQUIRE(kind == HPS_MLP_GEO || kind == HPS_MLP_COL, "bad decoder kind");
    HPS_REQUIRE(P >= 0, "negative point count");
    if (P == 0) return 0;
    HPS_REQUIRE(params && p && c && act && dout && dc, "null buffer");
    for (int i = 0; i < HPS_MLP_NPARAM; ++i) HPS_REQUIRE(params[i] != nullptr, "null parameter pointer");
    int rc = kind == HPS_MLP_GEO ? launch_bwd<HPS_MLP_GEO>(params, p, c, affine, act, dout, P, dparams, dc, dp,
                                                           daffine, (cudaStream_t)stream)
                                 : launch_bwd<HPS_MLP_COL>(params, p, c, affine, act, dout, P, dparams, dc, dp,
                                                           daffine, (cudaStream_t)stream);
    if (rc) return rc;
    HPS_LAUNCH_OK();
    return 0;
}
