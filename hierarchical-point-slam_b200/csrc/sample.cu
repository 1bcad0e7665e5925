// Depth-guided sampling along rays, per-point radius thresholds, and the backward into the rays.
//
// Restates the inline sampling block of Renderer.render_batch_ray (reference src/utils/Renderer.py:163-229,
// N_samples = 0):  t = linspace(0,1,S);  z = near*d*(1-t) + far*d*t;  pts = rays_o + rays_d * z;
// the per-ray float64 query radius is repeated per sample and turned into the two fp32 thresholds that
// reproduce the reference's float64 comparisons bit for bit:
//     D <  r^2 (f64)  <=>  D <  t_up      (neural_point.py:412, counting)
//     D >  r^2 (f64)  <=>  D >  t_dn      (decoder.py:236, weight zeroing)
// Every product and sum is rounded separately (no FMA contraction) in the reference's operation order, so
// sample points - and therefore neighbour sets - are bit-identical to the reference's.
#include "sample_point.cuh"

namespace hps {

__global__ void sample_rays_kernel(const RaySource src, int64_t R, float* __restrict__ z, float* __restrict__ pts,
                                   float* __restrict__ t_up, float* __restrict__ t_dn) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * src.S) return;
    const int64_t r = i / src.S;
    const SamplePoint sp = sample_point(src, r, (int)(i - r * src.S));
    z[i] = sp.z;
    pts[i * 3 + 0] = sp.x;
    pts[i * 3 + 1] = sp.y;
    pts[i * 3 + 2] = sp.zc;
    t_dn[i] = sp.t_dn;
    t_up[i] = sp.t_up;
}

// g_rays_o[r] = sum_s dp[r,s],  g_rays_d[r] = sum_s dp[r,s] * z[r,s]
// (dp = dp + dp2 + dp3 where given: the two decoders' and the gather's contributions, summed on the fly)
__global__ void sample_rays_bwd_kernel(const float* __restrict__ dp, const float* __restrict__ dp2,
                                       const float* __restrict__ dp3, const float* __restrict__ z, int64_t R, int S,
                                       float* __restrict__ g_o, float* __restrict__ g_d) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * 3) return;
    const int64_t r = i / 3;
    const int a = (int)(i - r * 3);
    float so = 0.f, sd = 0.f;
    for (int s = 0; s < S; ++s) {
        const int64_t e = (r * S + s) * 3 + a;
        float g = __ldg(dp + e);
        if (dp2 != nullptr) g += __ldg(dp2 + e);
        if (dp3 != nullptr) g += __ldg(dp3 + e);
        so += g;
        sd += g * __ldg(z + r * S + s);
    }
    g_o[i] = so;
    g_d[i] = sd;
}

}  // namespace hps

extern "C" int hps_sample_rays(const float* rays_o, const float* rays_d, const float* gt_depth, const double* radius,
                               double radius_scalar, const float* z_zero, int64_t z_zero_stride, int64_t R, int S,
                               float near, float far, float* z, float* pts, float* t_up, float* t_dn, void* stream) {
    using namespace hps;
    HPS_REQUIRE(R >= 0 && S >= 1 && S <= 32, "need R >= 0, 1 <= S <= 32");
    HPS_REQUIRE(z_zero_stride == 0 || z_zero_stride >= S, "z_zero_stride must be 0 (shared row) or >= S");
    if (R == 0) return 0;
    HPS_REQUIRE(rays_o && rays_d && gt_depth && z && pts && t_up && t_dn, "null buffer");
    const RaySource src{rays_o, rays_d, gt_depth, radius, radius_scalar, z_zero, z_zero_stride, S, near, far};
    HPS_SHARE_SM(sample_rays_kernel);
    sample_rays_kernel<<<(unsigned)ceil_div(R * S, 256), 256, 0, (cudaStream_t)stream>>>(src, R, z, pts, t_up, t_dn);
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_sample_rays_bwd(const float* dp, const float* dp2, const float* dp3, const float* z, int64_t R, int S,
                                   float* g_rays_o, float* g_rays_d, void* stream) {
    using namespace hps;
    HPS_REQUIRE(R >= 0 && S >= 1, "bad sizes");
    if (R == 0) return 0;
    HPS_REQUIRE(dp && z && g_rays_o && g_rays_d, "null buffer");
    sample_rays_bwd_kernel<<<(unsigned)ceil_div(R * 3, 256), 256, 0, (cudaStream_t)stream>>>(dp, dp2, dp3, z, R, S,
                                                                                             g_rays_o, g_rays_d);
    HPS_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Pose helpers of the tracker (reference src/common.py:304-346 get_camera_from_tensor / quad2rotation and
// :103-119 get_rays_from_uv) as single kernels with analytic backward: the op-by-op torch versions are ~100
// launches per tracker iteration (forward + autograd), more than the whole render path.
// Arithmetic follows the reference expression by expression (separately rounded, no FMA), so the rays - and with
// them the sample points and neighbour sets - are bit-identical to the torch ops.
// ------------------------------------------------------------------------------------------------
namespace hps {

// RT (3,4) row-major = [quad2rotation(q) | T] for cam = [qr, qi, qj, qk, tx, ty, tz]; q is NOT normalised
__global__ void camera_from_tensor_kernel(const float* __restrict__ cam, float* __restrict__ RT) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float qr = cam[0], qi = cam[1], qj = cam[2], qk = cam[3];
    // torch's CUDA reduction over the 4 squares associates (q0^2 + q2^2) + (q1^2 + q3^2) (probed on the device)
    const float n = __fadd_rn(__fadd_rn(__fmul_rn(qr, qr), __fmul_rn(qj, qj)), __fadd_rn(__fmul_rn(qi, qi), __fmul_rn(qk, qk)));
    const float s = __fdiv_rn(2.0f, n);
    const float rr = __fmul_rn(qr, qr), ii = __fmul_rn(qi, qi), jj = __fmul_rn(qj, qj), kk = __fmul_rn(qk, qk);
    (void)rr;
    RT[0] = __fsub_rn(1.0f, __fmul_rn(s, __fadd_rn(jj, kk)));
    RT[1] = __fmul_rn(s, __fsub_rn(__fmul_rn(qi, qj), __fmul_rn(qk, qr)));
    RT[2] = __fmul_rn(s, __fadd_rn(__fmul_rn(qi, qk), __fmul_rn(qj, qr)));
    RT[3] = cam[4];
    RT[4] = __fmul_rn(s, __fadd_rn(__fmul_rn(qi, qj), __fmul_rn(qk, qr)));
    RT[5] = __fsub_rn(1.0f, __fmul_rn(s, __fadd_rn(ii, kk)));
    RT[6] = __fmul_rn(s, __fsub_rn(__fmul_rn(qj, qk), __fmul_rn(qi, qr)));
    RT[7] = cam[5];
    RT[8] = __fmul_rn(s, __fsub_rn(__fmul_rn(qi, qk), __fmul_rn(qj, qr)));
    RT[9] = __fmul_rn(s, __fadd_rn(__fmul_rn(qj, qk), __fmul_rn(qi, qr)));
    RT[10] = __fsub_rn(1.0f, __fmul_rn(s, __fadd_rn(ii, jj)));
    RT[11] = cam[6];
}

// g_cam (7) from g_RT (3,4): R = I + s P(q), s = 2 / |q|^2  =>  dL/dq = s * <M, dP/dq> - s^2 <M, P> q
__global__ void camera_from_tensor_bwd_kernel(const float* __restrict__ cam, const float* __restrict__ g,
                                              float* __restrict__ g_cam) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float qr = cam[0], qi = cam[1], qj = cam[2], qk = cam[3];
    const float n = qr * qr + qi * qi + qj * qj + qk * qk, s = 2.0f / n;
    const float M00 = g[0], M01 = g[1], M02 = g[2], M10 = g[4], M11 = g[5], M12 = g[6], M20 = g[8], M21 = g[9], M22 = g[10];
    const float P00 = -(qj * qj + qk * qk), P01 = qi * qj - qk * qr, P02 = qi * qk + qj * qr;
    const float P10 = qi * qj + qk * qr, P11 = -(qi * qi + qk * qk), P12 = qj * qk - qi * qr;
    const float P20 = qi * qk - qj * qr, P21 = qj * qk + qi * qr, P22 = -(qi * qi + qj * qj);
    const float MP = M00 * P00 + M01 * P01 + M02 * P02 + M10 * P10 + M11 * P11 + M12 * P12 + M20 * P20 + M21 * P21 + M22 * P22;
    // <M, dP/dq_x>
    const float dr = -M01 * qk + M02 * qj + M10 * qk - M12 * qi - M20 * qj + M21 * qi;
    const float di = M01 * qj + M02 * qk + M10 * qj - 2.f * M11 * qi - M12 * qr + M20 * qk + M21 * qr - 2.f * M22 * qi;
    const float dj = -2.f * M00 * qj + M01 * qi + M02 * qr + M10 * qi + M12 * qk - M20 * qr + M21 * qk - 2.f * M22 * qj;
    const float dk = -2.f * M00 * qk - M01 * qr + M02 * qi + M10 * qr - 2.f * M11 * qk + M12 * qj + M20 * qi + M21 * qj;
    const float c = s * s * MP;
    g_cam[0] = s * dr - c * qr;
    g_cam[1] = s * di - c * qi;
    g_cam[2] = s * dj - c * qj;
    g_cam[3] = s * dk - c * qk;
    g_cam[4] = g[3];
    g_cam[5] = g[7];
    g_cam[6] = g[11];
}

// rays_d[r] = sum_b dirs[r][b] * c2w[a][b] (torch.sum(dirs * c2w[:3,:3], -1)), rays_o[r] = c2w[:3, 3]; c2w rows have
// `pitch` floats (4 for a (3,4) or (4,4) matrix)
__global__ void rays_from_uv_kernel(const float* __restrict__ pi, const float* __restrict__ pj,
                                    const float* __restrict__ c2w, int pitch, int64_t R, float fx, float fy, float cx,
                                    float cy, float* __restrict__ ro, float* __restrict__ rd) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    // torch divides a tensor by a python scalar as a multiplication with the fp32 reciprocal
    const float d0 = __fmul_rn(__fsub_rn(pi[r], cx), __fdiv_rn(1.0f, fx));
    const float d1 = -__fmul_rn(__fsub_rn(pj[r], cy), __fdiv_rn(1.0f, fy)), d2 = -1.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float* row = c2w + a * pitch;
        // torch's CUDA reduction over the 3 products associates (p0 + p2) + p1 (probed on the device)
        rd[r * 3 + a] = __fadd_rn(__fadd_rn(__fmul_rn(d0, row[0]), __fmul_rn(d2, row[2])), __fmul_rn(d1, row[1]));
        ro[r * 3 + a] = row[3];
    }
}

// g_c2w (3,4), zero-initialised by the caller: [sum_r g_d[r] (x) dirs[r] | sum_r g_o[r]]
__global__ void __launch_bounds__(256)
rays_from_uv_bwd_kernel(const float* __restrict__ pi, const float* __restrict__ pj, const float* __restrict__ g_o,
                        const float* __restrict__ g_d, int64_t R, float fx, float fy, float cx, float cy,
                        float* __restrict__ g_c2w) {
    __shared__ float part[8][12];
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = 0.f;
    if (r < R) {
        const float d[3] = {(pi[r] - cx) / fx, -(pj[r] - cy) / fy, -1.0f};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float gd = g_d != nullptr ? g_d[r * 3 + a] : 0.f;
#pragma unroll
            for (int b = 0; b < 3; ++b) v[a * 4 + b] = gd * d[b];
            v[a * 4 + 3] = g_o != nullptr ? g_o[r * 3 + a] : 0.f;
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        float s = v[k];
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) part[warp][k] = s;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w][threadIdx.x];
        red_add(g_c2w + threadIdx.x, s);
    }
}

}  // namespace hps

extern "C" int hps_camera_from_tensor(const float* cam7, float* RT, void* stream) {
    using namespace hps;
    HPS_REQUIRE(cam7 && RT, "null buffer");
    camera_from_tensor_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(cam7, RT);
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_camera_from_tensor_bwd(const float* cam7, const float* g_RT, float* g_cam7, void* stream) {
    using namespace hps;
    HPS_REQUIRE(cam7 && g_RT && g_cam7, "null buffer");
    camera_from_tensor_bwd_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(cam7, g_RT, g_cam7);
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_rays_from_uv(const float* pi, const float* pj, const float* c2w, int pitch, int64_t R, float fx,
                                float fy, float cx, float cy, float* rays_o, float* rays_d, void* stream) {
    using namespace hps;
    HPS_REQUIRE(R >= 0 && pitch >= 4, "bad sizes");
    if (R == 0) return 0;
    HPS_REQUIRE(pi && pj && c2w && rays_o && rays_d, "null buffer");
    rays_from_uv_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, (cudaStream_t)stream>>>(pi, pj, c2w, pitch, R, fx, fy, cx, cy,
                                                                                   rays_o, rays_d);
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_rays_from_uv_bwd(const float* pi, const float* pj, const float* g_rays_o, const float* g_rays_d,
                                    int64_t R, float fx, float fy, float cx, float cy, float* g_c2w, void* stream) {
    using namespace hps;
    HPS_REQUIRE(R >= 0, "bad sizes");
    if (R == 0) return 0;
    HPS_REQUIRE(pi && pj && g_c2w, "null buffer");
    rays_from_uv_bwd_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, (cudaStream_t)stream>>>(pi, pj, g_rays_o, g_rays_d, R, fx,
                                                                                       fy, cx, cy, g_c2w);
    HPS_LAUNCH_OK();
    return 0;
}
