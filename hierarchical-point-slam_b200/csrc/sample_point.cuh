// One sample point of a ray, exactly as the reference computes it (src/utils/Renderer.py:163-229, N_samples = 0):
// shared by the stand-alone sampling kernel (sample.cu) and the neighbour search that samples on the fly (grid.cu).
// Every product and sum is rounded separately (no FMA contraction), in the reference's operation order.
#pragma once
#include "hps_common.cuh"

namespace hps {

struct RaySource {
    const float* rays_o;       // (R,3)
    const float* rays_d;       // (R,3)
    const float* gt_depth;     // (R)
    const double* radius;      // (R) float64 per-ray query radius, or NULL -> radius_scalar
    double radius_scalar;
    const float* z_zero;       // z for rays without a depth reading: shared row (stride 0) or one row per ray; NULL -> 0
    int64_t z_zero_stride;
    int S;
    float near, far;
};

struct SamplePoint {
    float z, x, y, zc, t_up, t_dn;
};

__device__ __forceinline__ float linspace01(int i, int S) {
    if (S == 1) return 0.f;
    const float step = 1.0f / (float)(S - 1);
    return i < S / 2 ? __fmul_rn(step, (float)i) : __fsub_rn(1.0f, __fmul_rn(step, (float)(S - 1 - i)));
}

// sample s of ray r: z = near*d*(1-t) + far*d*t (t = linspace(0,1,S)[s]) or z_zero; p = o + d*z; the two fp32
// thresholds that reproduce the float64 comparisons against radius^2
__device__ __forceinline__ SamplePoint sample_point(const RaySource& src, int64_t r, int s) {
    SamplePoint sp;
    const float d = __ldg(src.gt_depth + r);
    if (d > 0.f) {
        const float t = linspace01(s, src.S);
        sp.z = __fadd_rn(__fmul_rn(__fmul_rn(src.near, d), __fsub_rn(1.0f, t)), __fmul_rn(__fmul_rn(src.far, d), t));
    } else {
        sp.z = src.z_zero ? __ldg(src.z_zero + r * src.z_zero_stride + s) : 0.f;
    }
    sp.x = __fadd_rn(__ldg(src.rays_o + r * 3 + 0), __fmul_rn(__ldg(src.rays_d + r * 3 + 0), sp.z));
    sp.y = __fadd_rn(__ldg(src.rays_o + r * 3 + 1), __fmul_rn(__ldg(src.rays_d + r * 3 + 1), sp.z));
    sp.zc = __fadd_rn(__ldg(src.rays_o + r * 3 + 2), __fmul_rn(__ldg(src.rays_d + r * 3 + 2), sp.z));
    const double rad = src.radius ? src.radius[r] : src.radius_scalar;
    const double r2 = rad * rad;
    const float t = __double2float_rn(r2);
    sp.t_dn = ((double)t > r2) ? nextafterf(t, -INFINITY) : t;
    sp.t_up = ((double)t < r2) ? nextafterf(t, INFINITY) : t;
    return sp;
}

}  // namespace hps
