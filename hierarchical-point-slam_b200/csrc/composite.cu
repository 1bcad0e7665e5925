// Alpha compositing along rays and its backward (kernel 4 of the north star).
//
// Restates raw2outputs_nerf_color with occupancy=True (reference src/common.py:385-428) fused with the
// no-neighbour overwrite raw[~point_mask, -1] = -100 (src/utils/Renderer.py:237-239):
//     alpha_s = sigmoid(coef * occ_s),  T_s = prod_{j<s} (1 - alpha_j + 1e-10),  w_s = alpha_s * T_s
//     rgb = sum w rgb / (sum w + 1e-10), depth = sum w z / (sum w + 1e-10), var = sum w (z - depth)^2
// One lane per sample, G lanes per ray (G = 8 for the reference's S = 5, G = 32 for S up to 32), so the
// packed (R,S,4) raw tensor is read with unit-stride, fully coalesced loads.  The exclusive transmittance
// product and the sums are taken in sample order through shuffles (same association as the reference's
// sequential cumprod / sum, so results agree to the last bits).
//
// Backward (only depth and rgb carry gradient):
//     gw_s = (g_depth (z_s - depth) + sum_c g_c (rgb_sc - rgb_c)) / wsum
//     d alpha_s = gw_s T_s - (sum_{k>s} gw_k w_k) / (1 - alpha_s + 1e-10)
//     d occ_s = d alpha_s * coef * alpha_s (1 - alpha_s)       (straight-through for masked points, A6)
#include "hps_common.cuh"

namespace hps {

struct RaySample {
    float alpha, trans, w, z, r, g, b;
};

template <int G>
__device__ __forceinline__ float group_bcast(float v, int src, int lane_base) {
    return __shfl_sync(kFull, v, lane_base + src);
}

// Forward quantities shared by fwd and bwd.  Every lane of the group returns the same sums.
template <int G>
__device__ __forceinline__ void composite_core(const float* __restrict__ rgb, int rgb_stride,
                                               const float* __restrict__ occ, int occ_stride,
                                               const uint8_t* __restrict__ point_mask,
                                               const float* __restrict__ z, int64_t ray, int S, int s,
                                               bool active, float coef, int lane_base, RaySample& me,
                                               float& wsum, float& depth, float& cr, float& cg, float& cb,
                                               float& var) {
    me = RaySample{0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float t = 1.f;
    if (active) {
        const int64_t i = ray * S + s;
        float o = __ldg(occ + i * occ_stride);
        if (point_mask != nullptr && point_mask[i] == 0) o = -100.0f;
        me.alpha = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-__fmul_rn(coef, o))));
        t = __fadd_rn(__fsub_rn(1.0f, me.alpha), 1e-10f);
        me.z = __ldg(z + i);
        if (rgb != nullptr) {
            me.r = __ldg(rgb + i * rgb_stride + 0);
            me.g = __ldg(rgb + i * rgb_stride + 1);
            me.b = __ldg(rgb + i * rgb_stride + 2);
        }
    }
    float T = 1.f;
    for (int j = 0; j < S - 1; ++j) {
        const float tj = group_bcast<G>(t, j, lane_base);
        if (j < s) T = __fmul_rn(T, tj);
    }
    me.trans = T;
    me.w = active ? __fmul_rn(me.alpha, T) : 0.f;
    float sw = 0.f, swz = 0.f, sr = 0.f, sg = 0.f, sb = 0.f;
    for (int j = 0; j < S; ++j) {
        const float wj = group_bcast<G>(me.w, j, lane_base);
        sw = __fadd_rn(sw, wj);
        swz = __fadd_rn(swz, __fmul_rn(wj, group_bcast<G>(me.z, j, lane_base)));
        sr = __fadd_rn(sr, __fmul_rn(wj, group_bcast<G>(me.r, j, lane_base)));
        sg = __fadd_rn(sg, __fmul_rn(wj, group_bcast<G>(me.g, j, lane_base)));
        sb = __fadd_rn(sb, __fmul_rn(wj, group_bcast<G>(me.b, j, lane_base)));
    }
    wsum = __fadd_rn(sw, 1e-10f);
    depth = __fdiv_rn(swz, wsum);
    cr = __fdiv_rn(sr, wsum);
    cg = __fdiv_rn(sg, wsum);
    cb = __fdiv_rn(sb, wsum);
    const float dz = __fsub_rn(me.z, depth);
    const float mine = __fmul_rn(__fmul_rn(me.w, dz), dz);
    var = 0.f;
    for (int j = 0; j < S; ++j) var = __fadd_rn(var, group_bcast<G>(mine, j, lane_base));
}

template <int G>
__global__ void __launch_bounds__(256)
composite_fwd_kernel(const float* __restrict__ rgb, int rgb_stride, const float* __restrict__ occ,
                     int occ_stride, const uint8_t* __restrict__ point_mask, const float* __restrict__ z,
                     int64_t R, int S, float coef, float* __restrict__ depth_out, float* __restrict__ var_out,
                     float* __restrict__ color_out, float* __restrict__ weights_out, int min_valid,
                     uint8_t* __restrict__ valid_ray, const float* __restrict__ gt_depth,
                     float* __restrict__ tracker_sum, int zero_missing) {
    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int sub = lane / G, s = lane % G;
    const int64_t ray_raw = warp * RPW + sub;
    if (warp * RPW >= R) return;
    const int64_t ray = ray_raw < R ? ray_raw : R - 1;   // keep whole warp converged for shuffles
    const bool live = ray_raw < R;
    const bool active = s < S;
    RaySample me;
    float wsum, depth, cr, cg, cb, var;
    composite_core<G>(rgb, rgb_stride, occ, occ_stride, point_mask, z, ray, S, s, active, coef, sub * G, me,
                      wsum, depth, cr, cg, cb, var);
    int n_has = 0;
    if (valid_ray != nullptr) {   // decoder.py:310-311: a ray is valid if >= min_valid samples have neighbours
        const bool pm = active && (point_mask == nullptr || point_mask[ray * S + s] != 0);
        const unsigned grp = (G == 32) ? kFull : (((1u << G) - 1u) << (sub * G));
        n_has = __popc(__ballot_sync(kFull, pm) & grp);
    }
    if (tracker_sum != nullptr) {   // Tracker.py:217: sum over the batch of |gt - depth| / sqrt(var + 1e-10)
        float t = 0.f;
        if (live && s == 0) {
            const float gd = __ldg(gt_depth + ray);
            // rays without a depth reading: the renderer reports depth 0 for them (Renderer.py:250-251)
            t = fabsf(gd - ((zero_missing && !(gd > 0.f)) ? 0.f : depth)) / sqrtf(var + 1e-10f);
        }
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(kFull, t, o);
        if (lane == 0) red_add(tracker_sum, t);
    }
    if (!live) return;
    if (s == 0 && valid_ray != nullptr) valid_ray[ray] = n_has >= min_valid ? 1 : 0;
    if (active && weights_out != nullptr) weights_out[ray * S + s] = me.w;
    if (s == 0) {
        depth_out[ray] = depth;
        var_out[ray] = var;
        color_out[ray * 3 + 0] = cr;
        color_out[ray * 3 + 1] = cg;
        color_out[ray * 3 + 2] = cb;
    }
}

template <int G>
__global__ void __launch_bounds__(256)
composite_bwd_kernel(const float* __restrict__ rgb, int rgb_stride, const float* __restrict__ occ,
                     int occ_stride, const uint8_t* __restrict__ point_mask, const float* __restrict__ z,
                     int64_t R, int S, float coef, const float* __restrict__ g_depth,
                     const float* __restrict__ g_color, const float* __restrict__ gscale,
                     float* __restrict__ d_rgb, int d_rgb_stride,
                     float* __restrict__ d_occ, int d_occ_stride, unsigned* __restrict__ amax_rgb,
                     unsigned* __restrict__ amax_occ) {
    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int sub = lane / G, s = lane % G;
    const int64_t ray_raw = warp * RPW + sub;
    if (warp * RPW >= R) return;
    const int64_t ray = ray_raw < R ? ray_raw : R - 1;
    const bool live = ray_raw < R;
    const bool active = s < S;
    const int lane_base = sub * G;
    RaySample me;
    float wsum, depth, cr, cg, cb, var;
    composite_core<G>(rgb, rgb_stride, occ, occ_stride, point_mask, z, ray, S, s, active, coef, lane_base, me,
                      wsum, depth, cr, cg, cb, var);
    const float gs = gscale != nullptr ? __ldg(gscale) : 1.0f;      // upstream gradient of a scalar loss
    const float gd = g_depth != nullptr ? __ldg(g_depth + ray) * gs : 0.f;
    float gr = 0.f, gg = 0.f, gb = 0.f;
    if (g_color != nullptr) {
        gr = __ldg(g_color + ray * 3 + 0) * gs;
        gg = __ldg(g_color + ray * 3 + 1) * gs;
        gb = __ldg(g_color + ray * 3 + 2) * gs;
    }
    const float gw = active ? (gd * (me.z - depth) + gr * (me.r - cr) + gg * (me.g - cg) + gb * (me.b - cb)) / wsum
                            : 0.f;
    const float gww = gw * me.w;
    float suffix = 0.f;
    for (int j = S - 1; j > 0; --j) {
        const float v = group_bcast<G>(gww, j, lane_base);
        if (j > s) suffix += v;
    }
    const bool store = live && active;
    const int64_t i = ray * S + s;
    const float t = __fadd_rn(__fsub_rn(1.0f, me.alpha), 1e-10f);
    const float dalpha = gw * me.trans - suffix / t;
    const float v_occ = store ? dalpha * coef * me.alpha * (1.0f - me.alpha) : 0.f;
    const float k = me.w / wsum;
    const float v_r = store ? gr * k : 0.f, v_g = store ? gg * k : 0.f, v_b = store ? gb * k : 0.f;
    if (store) {
        d_occ[i * d_occ_stride] = v_occ;
        if (d_rgb != nullptr) {
            d_rgb[i * d_rgb_stride + 0] = v_r;
            d_rgb[i * d_rgb_stride + 1] = v_g;
            d_rgb[i * d_rgb_stride + 2] = v_b;
        }
    }
    // optional: max |gradient| of the two outputs (the tensor-core decoder backward scales by it; NaN/inf ignored)
    if (amax_occ != nullptr || amax_rgb != nullptr) {
        float mo = fabsf(v_occ), mc = fmaxf(fabsf(v_r), fmaxf(fabsf(v_g), fabsf(v_b)));
        mo = mo <= 3.0e38f ? mo : 0.f;
        mc = mc <= 3.0e38f ? mc : 0.f;
        for (int o = 16; o; o >>= 1) {
            mo = fmaxf(mo, __shfl_xor_sync(0xffffffffu, mo, o));
            mc = fmaxf(mc, __shfl_xor_sync(0xffffffffu, mc, o));
        }
        if (lane == 0) {
            if (amax_occ != nullptr && mo > 0.f) atomicMax(amax_occ, __float_as_uint(mo));
            if (amax_rgb != nullptr && d_rgb != nullptr && mc > 0.f) atomicMax(amax_rgb, __float_as_uint(mc));
        }
    }
}

// Compositing forward fused with the mapper loss (reference src/Mapper.py:791-856): the ray's lane 0 holds depth and
// colour right after the sums, so the masked L1 depth term, the sigmoid(colour x affine) L1 colour term and their
// gradients wrt depth / colour / the exposure affine are evaluated in place; one block reduction and 13 atomics
// per block replace the separate loss kernel (same expressions as mapper_loss_kernel in loss.cu).
template <int G>
__global__ void __launch_bounds__(256)
composite_mapper_loss_kernel(const float* __restrict__ rgb, int rgb_stride, const float* __restrict__ occ,
                             int occ_stride, const uint8_t* __restrict__ point_mask, const float* __restrict__ z,
                             int64_t R, int S, float coef, int min_valid, const float* __restrict__ gt_depth,
                             const float* __restrict__ gt_color, const float* __restrict__ affine,
                             const int32_t* __restrict__ frame, float w_color, float* __restrict__ depth_out, float* __restrict__ var_out,
                             float* __restrict__ color_out, uint8_t* __restrict__ valid_ray,
                             float* __restrict__ loss, float* __restrict__ g_depth, float* __restrict__ g_color,
                             float* __restrict__ d_affine, float* __restrict__ d_rgb, float* __restrict__ d_occ,
                             unsigned* __restrict__ amax_rgb, unsigned* __restrict__ amax_occ,
                             float4* __restrict__ zero, int64_t zero_n4) {
    constexpr int RPW = 32 / G, NACC = 13;
    __shared__ float red[NACC][8];
    __shared__ float red_max[2][8];
    // the backward's gradient buffers are cleared here, beside the compositing (16 bytes per thread and round), instead
    // of by a memset between the loss and the first dgrad kernel
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < zero_n4; i += (int64_t)gridDim.x * blockDim.x)
        zero[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int sub = lane / G, s = lane % G;
    const int64_t ray_raw = warp * RPW + sub;
    float acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
    float mo = 0.f, mc = 0.f;          // max |d occ|, max |d rgb| of this warp's samples
    const bool use_aff = gt_color != nullptr && affine != nullptr;
    // rays of several keyframes in one batch (Mapper.py:805-822): frame[ray] selects the ray's exposure affine
    int f = 0, f_block = 0;
    if (frame != nullptr && use_aff) {
        f_block = __ldg(frame + min(((int64_t)blockIdx.x * blockDim.x >> 5) * RPW, R - 1));
        f = f_block;
    }
    if (warp * RPW < R) {                                  // warp-uniform: the shuffles below stay converged
        const int64_t ray = ray_raw < R ? ray_raw : R - 1;
        const bool live = ray_raw < R;
        const bool active = s < S;
        // the loss operands of the ray's leading lane are requested before the compositing loads, not after the sums
        const bool lead = live && s == 0;
        float gd = 0.f, gtc[3] = {0.f, 0.f, 0.f}, af[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) af[k] = 0.f;
        if (lead) {
            gd = __ldg(gt_depth + ray);
            if (gt_color != nullptr) {
                gtc[0] = __ldg(gt_color + ray * 3 + 0);
                gtc[1] = __ldg(gt_color + ray * 3 + 1);
                gtc[2] = __ldg(gt_color + ray * 3 + 2);
            }
            if (use_aff) {
                if (frame != nullptr) f = __ldg(frame + ray);
#pragma unroll
                for (int k = 0; k < 12; ++k) af[k] = __ldg(affine + 12 * (int64_t)f + k);
            }
        }
        RaySample me;
        float wsum, depth, cr, cg, cb, var;
        composite_core<G>(rgb, rgb_stride, occ, occ_stride, point_mask, z, ray, S, s, active, coef, sub * G, me,
                          wsum, depth, cr, cg, cb, var);
        const bool pm = active && (point_mask == nullptr || point_mask[ray * S + s] != 0);
        const unsigned grp = (G == 32) ? kFull : (((1u << G) - 1u) << (sub * G));
        const int n_has = __popc(__ballot_sync(kFull, pm) & grp);
        float bgd = 0.f, bgr = 0.f, bgg = 0.f, bgb = 0.f;       // the lead lane's d loss / d depth, d colour (for the group)
        if (lead) {
            const bool ok = n_has >= min_valid;            // decoder.py:310-311
            valid_ray[ray] = ok ? 1 : 0;
            depth_out[ray] = depth;
            var_out[ray] = var;
            color_out[ray * 3 + 0] = cr;
            color_out[ray * 3 + 1] = cg;
            color_out[ray * 3 + 2] = cb;
            const bool m = (gd > 0.f) && ok && !isnan(depth);
            float gdep = 0.f;
            if (m) {
                acc[0] = fabsf(gd - depth);
                gdep = depth > gd ? 1.f : (depth < gd ? -1.f : 0.f);
            }
            g_depth[ray] = gdep;
            if (gt_color != nullptr) {
                float gc[3] = {0.f, 0.f, 0.f};
                if (m) {
                    const float c[3] = {cr, cg, cb};
                    float dl[3];
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) {
                        float lin = c[cc];
                        if (use_aff) lin = c[0] * af[cc] + c[1] * af[3 + cc] + c[2] * af[6 + cc] + af[9 + cc];
                        const float sg = 1.0f / (1.0f + expf(-lin));
                        const float g = gtc[cc];
                        acc[0] += w_color * fabsf(g - sg);
                        dl[cc] = w_color * (sg > g ? 1.f : (sg < g ? -1.f : 0.f)) * sg * (1.0f - sg);
                    }
                    if (use_aff) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            gc[k] = dl[0] * af[3 * k] + dl[1] * af[3 * k + 1] + dl[2] * af[3 * k + 2];
#pragma unroll
                            for (int cc = 0; cc < 3; ++cc) acc[1 + k * 3 + cc] = c[k] * dl[cc];
                        }
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) acc[10 + cc] = dl[cc];
                    } else {
                        gc[0] = dl[0]; gc[1] = dl[1]; gc[2] = dl[2];
                    }
                }
                g_color[ray * 3 + 0] = gc[0];
                g_color[ray * 3 + 1] = gc[1];
                g_color[ray * 3 + 2] = gc[2];
                bgr = gc[0]; bgg = gc[1]; bgb = gc[2];
            }
            bgd = gdep;
        }
        if (d_occ != nullptr) {
            // The compositing backward right here (composite_bwd_kernel's expressions for an upstream gradient of 1): the
            // loss is a scalar, so d loss / d raw for ANY upstream gradient is this times that scalar - the decoders'
            // backward multiplies it in (hps_mlp_tc_bwd gscale) - and the separate compositing-backward launch between the
            // decoders' forward and backward disappears from the critical path.
            const int lane_base = sub * G;
            const float gd_ = __shfl_sync(kFull, bgd, lane_base), gr = __shfl_sync(kFull, bgr, lane_base),
                        gg = __shfl_sync(kFull, bgg, lane_base), gb = __shfl_sync(kFull, bgb, lane_base);
            const float gw = active ? (gd_ * (me.z - depth) + gr * (me.r - cr) + gg * (me.g - cg) + gb * (me.b - cb)) / wsum : 0.f;
            const float gww = gw * me.w;
            float suffix = 0.f;
            for (int j = S - 1; j > 0; --j) {
                const float v = group_bcast<G>(gww, j, lane_base);
                if (j > s) suffix += v;
            }
            const bool store = live && active;
            const int64_t i = ray * S + s;
            const float t = __fadd_rn(__fsub_rn(1.0f, me.alpha), 1e-10f);
            const float dalpha = gw * me.trans - suffix / t;
            const float v_occ = store ? dalpha * coef * me.alpha * (1.0f - me.alpha) : 0.f;
            const float k = me.w / wsum;
            const float v_r = store ? gr * k : 0.f, v_g = store ? gg * k : 0.f, v_b = store ? gb * k : 0.f;
            if (store) {
                d_occ[i] = v_occ;
                if (d_rgb != nullptr) {
                    d_rgb[i * 3 + 0] = v_r;
                    d_rgb[i * 3 + 1] = v_g;
                    d_rgb[i * 3 + 2] = v_b;
                }
            }
            mo = fabsf(v_occ), mc = fmaxf(fabsf(v_r), fmaxf(fabsf(v_g), fabsf(v_b)));
            mo = mo <= 3.0e38f ? mo : 0.f;
            mc = mc <= 3.0e38f ? mc : 0.f;
            for (int o = 16; o; o >>= 1) {
                mo = fmaxf(mo, __shfl_xor_sync(kFull, mo, o));
                mc = fmaxf(mc, __shfl_xor_sync(kFull, mc, o));
            }
        }
    }
    if (d_occ != nullptr && lane == 0) {       // one atomicMax pair per BLOCK below (per warp: 2,500 on one L2 line)
        red_max[0][wib] = mo;
        red_max[1][wib] = mc;
    }
    int n_acc = (use_aff && d_affine != nullptr) ? NACC : 1;
    if (n_acc == NACC && !__syncthreads_and(f == f_block)) {
        // the block straddles a frame boundary: per-ray atomics for the affine gradient, block sum for the loss only
#pragma unroll
        for (int k = 1; k < NACC; ++k)
            if (acc[k] != 0.f) red_add(d_affine + 12 * (int64_t)f + (k - 1), acc[k]);
        n_acc = 1;
    }
    for (int k = 0; k < n_acc; ++k) {
        float v = acc[k];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if (lane == 0) red[k][wib] = v;
    }
    __syncthreads();
    if (threadIdx.x < n_acc) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[threadIdx.x][w];
        red_add(threadIdx.x == 0 ? loss : d_affine + 12 * (int64_t)f_block + (threadIdx.x - 1), t);
    }
    if (d_occ != nullptr && (threadIdx.x == 32 || threadIdx.x == 64)) {
        const int which = threadIdx.x == 32 ? 0 : 1;
        float m = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red_max[which][w]);
        unsigned* dst = which == 0 ? amax_occ : (d_rgb != nullptr ? amax_rgb : nullptr);
        if (dst != nullptr && m > 0.f) atomicMax(dst, __float_as_uint(m));
    }
}

}  // namespace hps

extern "C" int hps_composite_mapper_loss(const float* rgb, int rgb_stride, const float* occ, int occ_stride,
                                         const uint8_t* point_mask, const float* z, int64_t R, int S, float coef,
                                         int min_valid, const float* gt_depth, const float* gt_color,
                                         const float* affine, const int32_t* frame, float w_color, float* depth,
                                         float* var, float* color,
                                         uint8_t* valid_ray, float* loss, float* g_depth, float* g_color,
                                         float* d_affine, float* d_rgb, float* d_occ, uint32_t* absmax_rgb,
                                         uint32_t* absmax_occ, float* zero, int64_t zero_floats, void* stream) {
    using namespace hps;
    HPS_REQUIRE(R >= 0 && S >= 1 && S <= 32, "need 1 <= S <= 32");
    if (R == 0) return 0;
    HPS_REQUIRE(occ && z && depth && var && color && valid_ray && gt_depth && loss && g_depth, "null buffer");
    HPS_REQUIRE(gt_color == nullptr || (rgb != nullptr && g_color != nullptr), "colour term needs rgb and g_color");
    HPS_REQUIRE(affine == nullptr || d_affine != nullptr, "affine needs d_affine");
    HPS_REQUIRE(d_rgb == nullptr || (d_occ != nullptr && rgb != nullptr), "d_rgb needs d_occ and rgb");
    HPS_REQUIRE(zero_floats >= 0 && zero_floats % 4 == 0 && (zero_floats == 0 || zero != nullptr) &&
                    (reinterpret_cast<uintptr_t>(zero) & 15) == 0,
                "zero: 16-byte aligned buffer of a multiple of 4 floats");
    const int tb = 256;
    float4* const zero4 = reinterpret_cast<float4*>(zero);
    const int64_t zero_n4 = zero_floats / 4;
    if (S <= 8) {
        const int64_t warps = ceil_div(R, 4);
        composite_mapper_loss_kernel<8><<<(unsigned)ceil_div(warps * 32, tb), tb, 0, (cudaStream_t)stream>>>(
            rgb, rgb_stride, occ, occ_stride, point_mask, z, R, S, coef, min_valid, gt_depth, gt_color, affine, frame,
            w_color, depth, var, color, valid_ray, loss, g_depth, g_color, d_affine, d_rgb, d_occ, absmax_rgb, absmax_occ,
            zero4, zero_n4);
    } else {
        composite_mapper_loss_kernel<32><<<(unsigned)ceil_div(R * 32, tb), tb, 0, (cudaStream_t)stream>>>(
            rgb, rgb_stride, occ, occ_stride, point_mask, z, R, S, coef, min_valid, gt_depth, gt_color, affine, frame,
            w_color, depth, var, color, valid_ray, loss, g_depth, g_color, d_affine, d_rgb, d_occ, absmax_rgb, absmax_occ,
            zero4, zero_n4);
    }
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_composite_fwd(const float* rgb, int rgb_stride, const float* occ, int occ_stride,
                                 const uint8_t* point_mask, const float* z, int64_t R, int S, float coef,
                                 float* depth, float* var, float* color, float* weights, int min_valid,
                                 uint8_t* valid_ray, const float* gt_depth, float* tracker_sum, int zero_missing,
                                 void* stream) {
    using namespace hps;
    HPS_REQUIRE(R >= 0 && S >= 1 && S <= 32, "need 1 <= S <= 32");
    if (R == 0) return 0;
    HPS_REQUIRE(occ && z && depth && var && color, "null buffer");
    HPS_REQUIRE(tracker_sum == nullptr || gt_depth != nullptr, "tracker_sum needs gt_depth");
    const int tb = 256;
    if (S <= 8) {
        const int64_t warps = ceil_div(R, 4);
        composite_fwd_kernel<8><<<(unsigned)ceil_div(warps * 32, tb), tb, 0, (cudaStream_t)stream>>>(
            rgb, rgb_stride, occ, occ_stride, point_mask, z, R, S, coef, depth, var, color, weights, min_valid,
            valid_ray, gt_depth, tracker_sum, zero_missing);
    } else {
        composite_fwd_kernel<32><<<(unsigned)ceil_div(R * 32, tb), tb, 0, (cudaStream_t)stream>>>(
            rgb, rgb_stride, occ, occ_stride, point_mask, z, R, S, coef, depth, var, color, weights, min_valid,
            valid_ray, gt_depth, tracker_sum, zero_missing);
    }
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_composite_bwd(const float* rgb, int rgb_stride, const float* occ, int occ_stride,
                                 const uint8_t* point_mask, const float* z, int64_t R, int S, float coef,
                                 const float* g_depth, const float* g_color, const float* gscale, float* d_rgb,
                                 int d_rgb_stride, float* d_occ, int d_occ_stride, uint32_t* absmax_rgb,
                                 uint32_t* absmax_occ, void* stream) {
    using namespace hps;
    HPS_REQUIRE(R >= 0 && S >= 1 && S <= 32, "need 1 <= S <= 32");
    if (R == 0) return 0;
    HPS_REQUIRE(occ && z && d_occ, "null buffer");
    HPS_REQUIRE(d_rgb == nullptr || rgb != nullptr, "d_rgb needs rgb");
    const int tb = 256;
    if (S <= 8) {
        const int64_t warps = ceil_div(R, 4);
        composite_bwd_kernel<8><<<(unsigned)ceil_div(warps * 32, tb), tb, 0, (cudaStream_t)stream>>>(
            rgb, rgb_stride, occ, occ_stride, point_mask, z, R, S, coef, g_depth, g_color, gscale, d_rgb,
            d_rgb_stride, d_occ, d_occ_stride, absmax_rgb, absmax_occ);
    } else {
        composite_bwd_kernel<32><<<(unsigned)ceil_div(R * 32, tb), tb, 0, (cudaStream_t)stream>>>(
            rgb, rgb_stride, occ, occ_stride, point_mask, z, R, S, coef, g_depth, g_color, gscale, d_rgb,
            d_rgb_stride, d_occ, d_occ_stride, absmax_rgb, absmax_occ);
    }
    HPS_LAUNCH_OK();
    return 0;
}
