// Fused inverse-distance-weighted feature gather and its backward scatter (kernel 2 of the north star).
//
// Restates MLP_geometry.get_feature_at_pos / MLP_color.get_feature_at_pos
// (reference src/conv_onet/models/decoder.py:195-272 and :491-559) for nn_weighting='distance':
//     w_j = 1/(D_j + 1e-10), w_j = 0 where D_j > r^2, w = w / max(sum|w|, 1e-12),
//     c   = sum_j w_j * feats[I_j];  c = fill where count < min_nn.
// One neighbour list feeds both feature tables (the reference queries twice with identical inputs,
// SURVEY A4).  The (P,8,32) gather the reference materialises never exists here: one warp per sample
// point, lane = (neighbour slot mod 4, float4 chunk), rows are read as 8 x 16-byte coalesced loads
// (one 128-byte line per row) and reduced with two shuffles.
//
// Backward: d feats[I_j] += w_j * dc as red.global.add.v4.f32 (the reference's index_put_ accumulate);
// tracker variant additionally dL/dD_j = -(w_j^2/W) <dc, f_j - c>,  dL/dp = sum_j dL/dD_j * 2 (p - x_j)
// (SURVEY A5).
#include "hps_common.cuh"

namespace hps {

constexpr int kInterpWarps = 8;

__global__ void __launch_bounds__(kInterpWarps * 32)
interp_fwd_kernel(const float* __restrict__ p, const int32_t* __restrict__ I, const float* __restrict__ D,
                  const float* __restrict__ t_dn, const int32_t* __restrict__ count,
                  const float* __restrict__ cloud_pos, const float* __restrict__ feats_a,
                  const float* __restrict__ feats_b, const float* __restrict__ fill_a,
                  const float* __restrict__ fill_b, int64_t P, int min_nn, float* __restrict__ w_out,
                  float* __restrict__ c_a, float* __restrict__ c_b, uint8_t* __restrict__ has_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t pi = (int64_t)blockIdx.x * kInterpWarps + warp;
    if (pi >= P) return;
    const int nb = lane >> 3, ch = lane & 7;

    int idx = -1;
    float d = FLT_MAX;
    if (lane < kK) {
        idx = __ldg(I + pi * kK + lane);
        d = __ldg(D + pi * kK + lane);
    }
    const float tdn = __ldg(t_dn + pi);
    if (cloud_pos != nullptr && lane < kK && idx >= 0) {   // tracker: D from positions (decoder.py:209-214)
        const float px = __ldg(p + pi * 3 + 0), py = __ldg(p + pi * 3 + 1), pz = __ldg(p + pi * 3 + 2);
        const float* x = cloud_pos + (int64_t)idx * 3;
        d = sqdist_rn(__ldg(x + 0), __ldg(x + 1), __ldg(x + 2), px, py, pz);
    }
    const bool inr = (idx >= 0) && !(d > tdn);
    float w = inr ? __fdiv_rn(1.0f, __fadd_rn(d, 1e-10f)) : 0.f;
    float W = w;
    W += __shfl_xor_sync(kFull, W, 1);
    W += __shfl_xor_sync(kFull, W, 2);
    W += __shfl_xor_sync(kFull, W, 4);
    w = __fdiv_rn(w, fmaxf(W, 1e-12f));
    const bool has = __ldg(count + pi) >= min_nn;

    float4 acc_a = make_float4(0.f, 0.f, 0.f, 0.f), acc_b = acc_a;
#pragma unroll
    for (int it = 0; it < kK / 4; ++it) {
        const int j = it * 4 + nb;
        const float wj = __shfl_sync(kFull, w, j);
        const int ij = __shfl_sync(kFull, idx, j);
        if (wj != 0.f) {
            const float4 va = ldg4(feats_a + (int64_t)ij * kC + ch * 4);
            acc_a.x += wj * va.x; acc_a.y += wj * va.y; acc_a.z += wj * va.z; acc_a.w += wj * va.w;
            if (feats_b != nullptr) {
                const float4 vb = ldg4(feats_b + (int64_t)ij * kC + ch * 4);
                acc_b.x += wj * vb.x; acc_b.y += wj * vb.y; acc_b.z += wj * vb.z; acc_b.w += wj * vb.w;
            }
        }
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        acc_a.x += __shfl_xor_sync(kFull, acc_a.x, o); acc_a.y += __shfl_xor_sync(kFull, acc_a.y, o);
        acc_a.z += __shfl_xor_sync(kFull, acc_a.z, o); acc_a.w += __shfl_xor_sync(kFull, acc_a.w, o);
        acc_b.x += __shfl_xor_sync(kFull, acc_b.x, o); acc_b.y += __shfl_xor_sync(kFull, acc_b.y, o);
        acc_b.z += __shfl_xor_sync(kFull, acc_b.z, o); acc_b.w += __shfl_xor_sync(kFull, acc_b.w, o);
    }
    if (nb == 0) {
        float4 oa = has ? acc_a : ldg4(fill_a + ch * 4);
        *reinterpret_cast<float4*>(c_a + pi * kC + ch * 4) = oa;
        if (feats_b != nullptr) {
            float4 ob = has ? acc_b : ldg4(fill_b + ch * 4);
            *reinterpret_cast<float4*>(c_b + pi * kC + ch * 4) = ob;
        }
    }
    if (lane < kK) w_out[pi * kK + lane] = w;
    if (lane == 0) has_out[pi] = has ? 1 : 0;
}

__global__ void __launch_bounds__(kInterpWarps * 32)
interp_bwd_kernel(const float* __restrict__ p, const int32_t* __restrict__ I, const float* __restrict__ w_in,
                  const uint8_t* __restrict__ has_in, const float* __restrict__ cloud_pos,
                  const float* __restrict__ feats_a, const float* __restrict__ feats_b,
                  const float* __restrict__ c_a, const float* __restrict__ c_b,
                  const float* __restrict__ dc_a, const float* __restrict__ dc_b, int64_t P,
                  float* __restrict__ dfeats_a, float* __restrict__ dfeats_b, float* __restrict__ dp,
                  const int32_t* __restrict__ row_map) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t pi = (int64_t)blockIdx.x * kInterpWarps + warp;
    if (pi >= P) return;
    const int nb = lane >> 3, ch = lane & 7;
    const bool has = has_in[pi] != 0;
    if (!has) {   // c was the constant fill vector: no gradient to features or position
        if (dp != nullptr && lane < 3) dp[pi * 3 + lane] = 0.f;
        return;
    }
    int idx = -1;
    float w = 0.f;
    if (lane < kK) {
        idx = __ldg(I + pi * kK + lane);
        w = __ldg(w_in + pi * kK + lane);
    }
    const bool tracker = (dp != nullptr);
    float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga, ca = ga, cb = ga;
    if (dc_a != nullptr) ga = ldg4(dc_a + pi * kC + ch * 4);
    if (dc_b != nullptr) gb = ldg4(dc_b + pi * kC + ch * 4);
    if (tracker) {
        if (dc_a != nullptr) ca = ldg4(c_a + pi * kC + ch * 4);
        if (dc_b != nullptr) cb = ldg4(c_b + pi * kC + ch * 4);
    }
    // Every lane reads the weight and index of its two neighbour slots itself (eight lanes share an address: one
    // broadcast transaction) and the row-map entries right behind them, all issued before the first reduction: the
    // r01o capture had 31 % of this kernel's stall samples on the shuffles that used to hand them out from lanes 0..7.
    float wjs[kK / 4];
    int ijs[kK / 4], gjs[kK / 4];
#pragma unroll
    for (int it = 0; it < kK / 4; ++it) {
        wjs[it] = __ldg(w_in + pi * kK + it * 4 + nb);
        ijs[it] = __ldg(I + pi * kK + it * 4 + nb);
    }
#pragma unroll
    for (int it = 0; it < kK / 4; ++it)      // gradient row: the table row itself, or its slot in the compact live-row
        gjs[it] = (row_map != nullptr && wjs[it] != 0.f) ? __ldg(row_map + ijs[it]) : ijs[it];   // buffer (-1 = frozen row)
    float gdot[kK / 4];
#pragma unroll
    for (int it = 0; it < kK / 4; ++it) {
        const float wj = wjs[it];
        const int ij = ijs[it];
        float part = 0.f;
        if (wj != 0.f) {
            const int gj = gjs[it];
            if (dc_a != nullptr) {
                if (dfeats_a != nullptr && gj >= 0)
                    red_add_v4(dfeats_a + (int64_t)gj * kC + ch * 4,
                               make_float4(wj * ga.x, wj * ga.y, wj * ga.z, wj * ga.w));
                if (tracker) {
                    const float4 f = ldg4(feats_a + (int64_t)ij * kC + ch * 4);
                    part += ga.x * (f.x - ca.x) + ga.y * (f.y - ca.y) + ga.z * (f.z - ca.z) + ga.w * (f.w - ca.w);
                }
            }
            if (dc_b != nullptr) {
                if (dfeats_b != nullptr && gj >= 0)
                    red_add_v4(dfeats_b + (int64_t)gj * kC + ch * 4,
                               make_float4(wj * gb.x, wj * gb.y, wj * gb.z, wj * gb.w));
                if (tracker) {
                    const float4 f = ldg4(feats_b + (int64_t)ij * kC + ch * 4);
                    part += gb.x * (f.x - cb.x) + gb.y * (f.y - cb.y) + gb.z * (f.z - cb.z) + gb.w * (f.w - cb.w);
                }
            }
        }
        if (tracker) {   // warp-uniform; the mapper has no use for <dc, f_j - c> (13 % of this kernel's stall samples, r01o)
            part += __shfl_xor_sync(kFull, part, 1);
            part += __shfl_xor_sync(kFull, part, 2);
            part += __shfl_xor_sync(kFull, part, 4);
        }
        gdot[it] = part;
    }
    if (!tracker) return;
    // lane j < 8 fetches <dc, f_j - c> from group (j & 3), iteration (j >> 2)
    float g = 0.f;
#pragma unroll
    for (int it = 0; it < kK / 4; ++it) {
        const float v = __shfl_sync(kFull, gdot[it], (lane & 3) * 8);
        if ((lane >> 2) == it) g = v;
    }
    const float px = __ldg(p + pi * 3 + 0), py = __ldg(p + pi * 3 + 1), pz = __ldg(p + pi * 3 + 2);
    float ex = 0.f, ey = 0.f, ez = 0.f, wr = 0.f;
    if (lane < kK && w != 0.f) {
        const float* x = cloud_pos + (int64_t)idx * 3;
        const float xx = __ldg(x + 0), xy = __ldg(x + 1), xz = __ldg(x + 2);
        const float d = sqdist_rn(xx, xy, xz, px, py, pz);
        wr = __fdiv_rn(1.0f, __fadd_rn(d, 1e-10f));
        ex = px - xx; ey = py - xy; ez = pz - xz;
    }
    float W = wr;
    W += __shfl_xor_sync(kFull, W, 1);
    W += __shfl_xor_sync(kFull, W, 2);
    W += __shfl_xor_sync(kFull, W, 4);
    const float dD = (lane < kK && w != 0.f) ? -(wr * wr / fmaxf(W, 1e-12f)) * g : 0.f;
    float gx = 2.f * dD * ex, gy = 2.f * dD * ey, gz = 2.f * dD * ez;
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1) {
        gx += __shfl_xor_sync(kFull, gx, o);
        gy += __shfl_xor_sync(kFull, gy, o);
        gz += __shfl_xor_sync(kFull, gz, o);
    }
    if (lane == 0) {
        dp[pi * 3 + 0] = gx;
        dp[pi * 3 + 1] = gy;
        dp[pi * 3 + 2] = gz;
    }
}

// table[idx[r]] = rows[r] for up to two (table, rows, idx) triples in one launch: the per-iteration write-back of the
// optimised leaves into the full feature tables (Mapper.py:615-630).  Eight lanes move one 128-byte row as float4.
__global__ void __launch_bounds__(256) rows_store_kernel(float4* __restrict__ ta, const float4* __restrict__ ra,
                                                         const int64_t* __restrict__ ia, int64_t na,
                                                         float4* __restrict__ tb, const float4* __restrict__ rb,
                                                         const int64_t* __restrict__ ib, int64_t nb) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t row = t >> 3;
    const int lane = (int)(t & 7);
    if (row < na) {
        ta[__ldg(ia + row) * (kC / 4) + lane] = __ldg(ra + row * (kC / 4) + lane);
    } else if ((row -= na) < nb) {
        tb[__ldg(ib + row) * (kC / 4) + lane] = __ldg(rb + row * (kC / 4) + lane);
    }
}

}  // namespace hps

extern "C" int hps_interp_fwd(const float* p, const int32_t* I, const float* D, const float* t_dn,
                              const int32_t* count, const float* cloud_pos, const float* feats_a,
                              const float* feats_b, const float* fill_a, const float* fill_b, int64_t P,
                              int min_nn, float* w, float* c_a, float* c_b, uint8_t* has, void* stream) {
    using namespace hps;
    HPS_REQUIRE(P >= 0, "negative point count");
    if (P == 0) return 0;
    HPS_REQUIRE(I && D && t_dn && count && feats_a && fill_a && w && c_a && has, "null buffer");
    HPS_REQUIRE(cloud_pos == nullptr || p != nullptr, "tracker variant needs p");
    HPS_REQUIRE(feats_b == nullptr || (fill_b && c_b), "second table needs fill_b and c_b");
    interp_fwd_kernel<<<(unsigned)ceil_div(P, kInterpWarps), kInterpWarps * 32, 0, (cudaStream_t)stream>>>(
        p, I, D, t_dn, count, cloud_pos, feats_a, feats_b, fill_a, fill_b, P, min_nn, w, c_a, c_b, has);
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_rows_store(float* table_a, const float* rows_a, const int64_t* idx_a, int64_t n_a, float* table_b,
                              const float* rows_b, const int64_t* idx_b, int64_t n_b, void* stream) {
    using namespace hps;
    HPS_REQUIRE(n_a >= 0 && n_b >= 0, "negative row count");
    HPS_REQUIRE(n_a == 0 || (table_a && rows_a && idx_a), "null buffer (first table)");
    HPS_REQUIRE(n_b == 0 || (table_b && rows_b && idx_b), "null buffer (second table)");
    if (n_a + n_b == 0) return 0;
    rows_store_kernel<<<(unsigned)ceil_div((n_a + n_b) * (kC / 4), 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4*>(table_a), reinterpret_cast<const float4*>(rows_a), idx_a, n_a,
        reinterpret_cast<float4*>(table_b), reinterpret_cast<const float4*>(rows_b), idx_b, n_b);
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_interp_bwd(const float* p, const int32_t* I, const float* w, const uint8_t* has,
                              const float* cloud_pos, const float* feats_a, const float* feats_b,
                              const float* c_a, const float* c_b, const float* dc_a, const float* dc_b,
                              int64_t P, float* dfeats_a, float* dfeats_b, float* dp, const int32_t* row_map,
                              void* stream) {
    using namespace hps;
    HPS_REQUIRE(P >= 0, "negative point count");
    if (P == 0) return 0;
    HPS_REQUIRE(I && w && has, "null buffer");
    HPS_REQUIRE(dp == nullptr || (p && cloud_pos), "dp needs p and cloud_pos");
    HPS_REQUIRE(dp == nullptr || dc_a == nullptr || (feats_a && c_a), "dp needs feats_a and c_a");
    HPS_REQUIRE(dp == nullptr || dc_b == nullptr || (feats_b && c_b), "dp needs feats_b and c_b");
    HPS_SHARE_SM(interp_bwd_kernel);
    interp_bwd_kernel<<<(unsigned)ceil_div(P, kInterpWarps), kInterpWarps * 32, 0, (cudaStream_t)stream>>>(
        p, I, w, has, cloud_pos, feats_a, feats_b, c_a, c_b, dc_a, dc_b, P, dfeats_a, dfeats_b, dp, row_map);
    HPS_LAUNCH_OK();
    return 0;
}
