// Mapper and tracker losses fused with their gradients (no host synchronisation, no boolean indexing).
//
// Mapper (reference src/Mapper.py:791-856, one frame's exposure affine per call):
//     mask = (gt_depth > 0) & valid_ray & ~isnan(depth)
//     loss = sum_mask |gt_depth - depth|  +  w * sum_mask |gt_color - sigmoid(color @ A[:9].view(3,3) + A[9:])|
// Tracker (reference src/Tracker.py:213-231, handle_dynamic, use_color_in_tracking):
//     tmp  = |gt_depth - depth| / sqrt(var + 1e-10);  mask = (tmp < 10 * mean(tmp)) & (gt_depth > 0) & ~nan
//     loss = sum_mask clamp(tmp, 0, 1e3)  +  w * sum_mask |gt_color - color|
// The loss is a plain sum, so the gradient w.r.t. depth / colour / affine is produced in the same pass.
#include "hps_common.cuh"

namespace hps {

__device__ __forceinline__ float sgn(float x) { return (x > 0.f) - (x < 0.f); }

__device__ __forceinline__ float block_sum(float v, float* scratch) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float t = (threadIdx.x < (blockDim.x >> 5)) ? scratch[threadIdx.x] : 0.f;
    if (warp == 0)
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(kFull, t, o);
    return t;   // valid in thread 0
}

__global__ void __launch_bounds__(256)
mapper_loss_kernel(const float* __restrict__ depth, const float* __restrict__ color,
                   const uint8_t* __restrict__ valid, const float* __restrict__ gt_depth,
                   const float* __restrict__ gt_color, const float* __restrict__ affine,
                   const int32_t* __restrict__ frame, float w_color, int64_t R,
                   float* __restrict__ loss, float* __restrict__ g_depth, float* __restrict__ g_color,
                   float* __restrict__ d_affine) {
    __shared__ float scratch[32];
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f;
    float da[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) da[k] = 0.f;
    // rays of several keyframes in one batch (Mapper.py:805-822): frame[r] selects the ray's exposure affine
    int f = 0, f_block = 0;
    if (frame != nullptr && affine != nullptr) {
        f_block = __ldg(frame + min((int64_t)blockIdx.x * blockDim.x, R - 1));
        f = r < R ? __ldg(frame + r) : f_block;
        affine += 12 * (int64_t)f;
    }
    if (r < R) {
        const float d = depth[r], gd = gt_depth[r];
        const bool m = (gd > 0.f) && (valid[r] != 0) && !isnan(d);
        float gdep = 0.f;
        if (m) {
            l = fabsf(gd - d);
            gdep = sgn(d - gd);
        }
        g_depth[r] = gdep;
        if (color != nullptr) {
            float gc[3] = {0.f, 0.f, 0.f};
            if (m) {
                const float c[3] = {color[r * 3 + 0], color[r * 3 + 1], color[r * 3 + 2]};
                float dl[3];
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    float lin = c[cc];
                    if (affine != nullptr)
                        lin = c[0] * __ldg(affine + cc) + c[1] * __ldg(affine + 3 + cc) + c[2] * __ldg(affine + 6 + cc) +
                              __ldg(affine + 9 + cc);
                    const float s = 1.0f / (1.0f + expf(-lin));
                    const float g = gt_color[r * 3 + cc];
                    l += w_color * fabsf(g - s);
                    dl[cc] = w_color * sgn(s - g) * s * (1.0f - s);
                }
                if (affine != nullptr) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        gc[k] = dl[0] * __ldg(affine + 3 * k) + dl[1] * __ldg(affine + 3 * k + 1) + dl[2] * __ldg(affine + 3 * k + 2);
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) da[k * 3 + cc] = c[k] * dl[cc];
                    }
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) da[9 + cc] = dl[cc];
                } else {
                    gc[0] = dl[0]; gc[1] = dl[1]; gc[2] = dl[2];
                }
            }
            g_color[r * 3 + 0] = gc[0];
            g_color[r * 3 + 1] = gc[1];
            g_color[r * 3 + 2] = gc[2];
        }
    }
    const float bl = block_sum(l, scratch);
    if (threadIdx.x == 0) red_add(loss, bl);
    if (color != nullptr && affine != nullptr && d_affine != nullptr) {
        // one frame per block (the usual case: frames are contiguous runs of rays): block sums; else per-ray atomics
        if (__syncthreads_and(f == f_block)) {
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const float s = block_sum(da[k], scratch);
                if (threadIdx.x == 0) red_add(d_affine + 12 * (int64_t)f_block + k, s);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 12; ++k)
                if (da[k] != 0.f) red_add(d_affine + 12 * (int64_t)f + k, da[k]);
        }
    }
}

__global__ void __launch_bounds__(256)
tracker_stats_kernel(const float* __restrict__ depth, const float* __restrict__ var,
                     const float* __restrict__ gt_depth, int64_t R, float* __restrict__ sum_tmp) {
    __shared__ float scratch[32];
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float t = 0.f;
    if (r < R) t = fabsf(gt_depth[r] - depth[r]) / sqrtf(var[r] + 1e-10f);
    const float s = block_sum(t, scratch);
    if (threadIdx.x == 0) red_add(sum_tmp, s);
}

__global__ void __launch_bounds__(256)
tracker_loss_kernel(const float* __restrict__ depth, const float* __restrict__ var, const float* __restrict__ color,
                    const float* __restrict__ gt_depth, const float* __restrict__ gt_color,
                    const float* __restrict__ sum_tmp, float count, float w_color, int64_t R,
                    float* __restrict__ loss, float* __restrict__ g_depth, float* __restrict__ g_color) {
    __shared__ float scratch[32];
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f;
    if (r < R) {
        const float d = depth[r], v = var[r], gd = gt_depth[r];
        const float inv = 1.0f / sqrtf(v + 1e-10f);
        const float tmp = fabsf(gd - d) * inv;
        const float thr = 10.0f * (__ldg(sum_tmp) / count);
        const bool m = (tmp < thr) && (gd > 0.f) && !isnan(d) && !isnan(v);
        float gdep = 0.f, gc[3] = {0.f, 0.f, 0.f};
        if (m) {
            l = fminf(fmaxf(tmp, 0.f), 1e3f);
            if (tmp <= 1e3f) gdep = sgn(d - gd) * inv;
            if (color != nullptr) {
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    const float c = color[r * 3 + cc], g = gt_color[r * 3 + cc];
                    l += w_color * fabsf(g - c);
                    gc[cc] = w_color * sgn(c - g);
                }
            }
        }
        g_depth[r] = gdep;
        if (color != nullptr) {
            g_color[r * 3 + 0] = gc[0];
            g_color[r * 3 + 1] = gc[1];
            g_color[r * 3 + 2] = gc[2];
        }
    }
    const float bl = block_sum(l, scratch);
    if (threadIdx.x == 0) red_add(loss, bl);
}

}  // namespace hps

extern "C" int hps_mapper_loss(const float* depth, const float* color, const uint8_t* valid, const float* gt_depth,
                               const float* gt_color, const float* affine, const int32_t* frame, float w_color,
                               int64_t R, float* loss, float* g_depth, float* g_color, float* d_affine, void* stream) {
    using namespace hps;
    HPS_REQUIRE(R >= 0, "negative ray count");
    if (R == 0) return 0;
    HPS_REQUIRE(depth && valid && gt_depth && loss && g_depth, "null buffer");
    HPS_REQUIRE(color == nullptr || (gt_color && g_color), "colour term needs gt_color and g_color");
    mapper_loss_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, (cudaStream_t)stream>>>(
        depth, color, valid, gt_depth, gt_color, affine, frame, w_color, R, loss, g_depth, g_color, d_affine);
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_tracker_loss_stats(const float* depth, const float* var, const float* gt_depth, int64_t R,
                                      float* sum_tmp, void* stream) {
    using namespace hps;
    HPS_REQUIRE(R >= 0, "negative ray count");
    if (R == 0) return 0;
    HPS_REQUIRE(depth && var && gt_depth && sum_tmp, "null buffer");
    tracker_stats_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, (cudaStream_t)stream>>>(depth, var, gt_depth, R, sum_tmp);
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_tracker_loss(const float* depth, const float* var, const float* color, const float* gt_depth,
                                const float* gt_color, const float* sum_tmp, double count, float w_color, int64_t R,
                                float* loss, float* g_depth, float* g_color, void* stream) {
    using namespace hps;
    HPS_REQUIRE(R >= 0 && count > 0, "bad sizes");
    if (R == 0) return 0;
    HPS_REQUIRE(depth && var && gt_depth && sum_tmp && loss && g_depth, "null buffer");
    HPS_REQUIRE(color == nullptr || (gt_color && g_color), "colour term needs gt_color and g_color");
    tracker_loss_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, (cudaStream_t)stream>>>(
        depth, var, color, gt_depth, gt_color, sum_tmp, (float)count, w_color, R, loss, g_depth, g_color);
    HPS_LAUNCH_OK();
    return 0;
}


// ------------------------------------------------------------------------------------------------
// Exposure MLP (reference src/conv_onet/models/decoder.py:380-396): one CTA, one thread per hidden unit.
// ------------------------------------------------------------------------------------------------
namespace hps {

__global__ void exposure_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W1,
                                    const float* __restrict__ b1, const float* __restrict__ W2,
                                    const float* __restrict__ b2, int n_in, int hidden, int n_out,
                                    float* __restrict__ a1, float* __restrict__ out) {
    extern __shared__ float sh[];       // h[hidden]
    const int j = threadIdx.x;
    if (j < hidden) {
        float a = b1[j];
        for (int k = 0; k < n_in; ++k) a = fmaf(W1[j * n_in + k], x[k], a);
        a1[j] = a;
        const float t = a * 100.0f;     // nn.Softplus(beta=100, threshold=20)
        sh[j] = t > 20.0f ? a : log1pf(expf(t)) / 100.0f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int o = warp; o < n_out; o += nw) {
        float s = 0.f;
        for (int k = lane; k < hidden; k += 32) s = fmaf(W2[o * hidden + k], sh[k], s);
        for (int m = 16; m; m >>= 1) s += __shfl_xor_sync(kFull, s, m);
        if (lane == 0) out[o] = s + b2[o];
    }
}

__global__ void exposure_bwd_kernel(const float* __restrict__ x, const float* __restrict__ W1,
                                    const float* __restrict__ W2, const float* __restrict__ a1,
                                    const float* __restrict__ dout, const float* __restrict__ gscale, int n_in,
                                    int hidden, int n_out, float* __restrict__ dW1, float* __restrict__ db1, float* __restrict__ dW2,
                                    float* __restrict__ db2, float* __restrict__ dx) {
    extern __shared__ float sh[];       // da[hidden]
    const int j = threadIdx.x;
    const float gs = gscale != nullptr ? __ldg(gscale) : 1.0f;      // upstream gradient of a scalar loss
    if (j < hidden) {
        const float a = a1[j];
        const float t = a * 100.0f;
        const float h = t > 20.0f ? a : log1pf(expf(t)) / 100.0f;
        float dh = 0.f;
        for (int o = 0; o < n_out; ++o) {
            const float g = dout[o] * gs;
            dW2[o * hidden + j] = g * h;
            dh = fmaf(W2[o * hidden + j], g, dh);
        }
        float sg = 1.f;                 // d softplus = sigmoid(100 a) below the threshold
        if (t <= 20.0f) {
            const float z = expf(t);
            sg = z / (z + 1.0f);
        }
        const float da = dh * sg;
        db1[j] = da;
        sh[j] = da;
        for (int k = 0; k < n_in; ++k) dW1[j * n_in + k] = da * x[k];
    }
    if (j < n_out) db2[j] = dout[j] * gs;
    __syncthreads();
    if (dx != nullptr) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        for (int k = warp; k < n_in; k += nw) {
            float s = 0.f;
            for (int i = lane; i < hidden; i += 32) s = fmaf(W1[i * n_in + k], sh[i], s);
            for (int m = 16; m; m >>= 1) s += __shfl_xor_sync(kFull, s, m);
            if (lane == 0) dx[k] = s;
        }
    }
}

}  // namespace hps

extern "C" int hps_exposure_fwd(const float* x, const float* W1, const float* b1, const float* W2, const float* b2,
                                int n_in, int hidden, int n_out, float* a1, float* out, void* stream) {
    using namespace hps;
    HPS_REQUIRE(n_in >= 1 && n_in <= 64 && hidden >= 1 && hidden <= 1024 && n_out >= 1 && n_out <= 32, "exposure MLP size");
    HPS_REQUIRE(x && W1 && b1 && W2 && b2 && a1 && out, "null buffer");
    const int tb = (max(hidden, 32) + 31) / 32 * 32;
    exposure_fwd_kernel<<<1, tb, hidden * sizeof(float), (cudaStream_t)stream>>>(x, W1, b1, W2, b2, n_in, hidden, n_out, a1,
                                                                              out);
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_exposure_bwd(const float* x, const float* W1, const float* W2, const float* a1, const float* dout,
                                const float* gscale, int n_in, int hidden, int n_out, float* dW1, float* db1, float* dW2, float* db2,
                                float* dx, void* stream) {
    using namespace hps;
    HPS_REQUIRE(n_in >= 1 && n_in <= 64 && hidden >= 1 && hidden <= 1024 && n_out >= 1 && n_out <= 32, "exposure MLP size");
    HPS_REQUIRE(x && W1 && W2 && a1 && dout && dW1 && db1 && dW2 && db2, "null buffer");
    const int tb = (max(hidden, 32) + 31) / 32 * 32;
    HPS_SHARE_SM(exposure_bwd_kernel);      // (it is issued beside the decoders' dgrad kernels and the backward tail)
    exposure_bwd_kernel<<<1, tb, hidden * sizeof(float), (cudaStream_t)stream>>>(x, W1, W2, a1, dout, gscale, n_in, hidden,
                                                                              n_out, dW1, db1, dW2, db2, dx);
    HPS_LAUNCH_OK();
    return 0;
}
