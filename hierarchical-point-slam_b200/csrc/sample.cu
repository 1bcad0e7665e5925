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
#include "hps_common.cuh"

namespace hps {

__device__ __forceinline__ float linspace01(int i, int S) {
    if (S == 1) return 0.f;
    const float step = 1.0f / (float)(S - 1);
    return i < S / 2 ? __fmul_rn(step, (float)i) : __fsub_rn(1.0f, __fmul_rn(step, (float)(S - 1 - i)));
}

__global__ void sample_rays_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                   const float* __restrict__ gt_depth, const double* __restrict__ radius,
                                   double radius_scalar, const float* __restrict__ z_zero, int64_t R, int S,
                                   float near, float far, float* __restrict__ z, float* __restrict__ pts,
                                   float* __restrict__ t_up, float* __restrict__ t_dn) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * S) return;
    const int64_t r = i / S;
    const int s = (int)(i - r * S);
    const float d = __ldg(gt_depth + r);
    float zz;
    if (d > 0.f) {
        const float t = linspace01(s, S);
        zz = __fadd_rn(__fmul_rn(__fmul_rn(near, d), __fsub_rn(1.0f, t)), __fmul_rn(__fmul_rn(far, d), t));
    } else {
        zz = z_zero ? __ldg(z_zero + s) : 0.f;
    }
    z[i] = zz;
#pragma unroll
    for (int a = 0; a < 3; ++a)
        pts[i * 3 + a] = __fadd_rn(__ldg(rays_o + r * 3 + a), __fmul_rn(__ldg(rays_d + r * 3 + a), zz));
    const double rad = radius ? radius[r] : radius_scalar;
    const double r2 = rad * rad;
    const float t = __double2float_rn(r2);
    const float dn = ((double)t > r2) ? nextafterf(t, -INFINITY) : t;
    const float up = ((double)t < r2) ? nextafterf(t, INFINITY) : t;
    t_dn[i] = dn;
    t_up[i] = up;
}

// g_rays_o[r] = sum_s dp[r,s],  g_rays_d[r] = sum_s dp[r,s] * z[r,s]
__global__ void sample_rays_bwd_kernel(const float* __restrict__ dp, const float* __restrict__ z, int64_t R, int S,
                                       float* __restrict__ g_o, float* __restrict__ g_d) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * 3) return;
    const int64_t r = i / 3;
    const int a = (int)(i - r * 3);
    float so = 0.f, sd = 0.f;
    for (int s = 0; s < S; ++s) {
        const float g = __ldg(dp + (r * S + s) * 3 + a);
        so += g;
        sd += g * __ldg(z + r * S + s);
    }
    g_o[i] = so;
    g_d[i] = sd;
}

}  // namespace hps

extern "C" int hps_sample_rays(const float* rays_o, const float* rays_d, const float* gt_depth, const double* radius,
                               double radius_scalar, const float* z_zero, int64_t R, int S, float near, float far,
                               float* z, float* pts, float* t_up, float* t_dn, void* stream) {
    using namespace hps;
    HPS_REQUIRE(R >= 0 && S >= 1 && S <= 32, "need R >= 0, 1 <= S <= 32");
    if (R == 0) return 0;
    HPS_REQUIRE(rays_o && rays_d && gt_depth && z && pts && t_up && t_dn, "null buffer");
    sample_rays_kernel<<<(unsigned)ceil_div(R * S, 256), 256, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, gt_depth, radius, radius_scalar, z_zero, R, S, near, far, z, pts, t_up, t_dn);
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_sample_rays_bwd(const float* dp, const float* z, int64_t R, int S, float* g_rays_o, float* g_rays_d,
                                   void* stream) {
    using namespace hps;
    HPS_REQUIRE(R >= 0 && S >= 1, "bad sizes");
    if (R == 0) return 0;
    HPS_REQUIRE(dp && z && g_rays_o && g_rays_d, "null buffer");
    sample_rays_bwd_kernel<<<(unsigned)ceil_div(R * 3, 256), 256, 0, (cudaStream_t)stream>>>(dp, z, R, S, g_rays_o,
                                                                                             g_rays_d);
    HPS_LAUNCH_OK();
    return 0;
}
