// Shared helpers for the hps_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <float.h>

#include "../../include/hps_b200.h"

namespace hps {

void set_error(const char* fmt, ...);

#define HPS_CUDA_OK(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            hps::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,             \
                           cudaGetErrorString(_e));                                         \
            return -2;                                                                      \
        }                                                                                   \
    } while (0)

#define HPS_REQUIRE(cond, msg)                                                              \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            hps::set_error("%s (%s) at %s:%d", msg, #cond, __FILE__, __LINE__);             \
            return -1;                                                                      \
        }                                                                                   \
    } while (0)

#define HPS_LAUNCH_OK()                                                                     \
    do {                                                                                    \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess) {                                                            \
            hps::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,         \
                           cudaGetErrorString(_e));                                         \
            return -3;                                                                      \
        }                                                                                   \
    } while (0)

// Shared-memory carve-out of kernels that are meant to run beside other kernels.  An SM is configured (L1 / shared
// split) for the first CTA that lands on it and is not re-split while CTAs are resident; a CTA whose kernel wants another
// split waits until the SM has drained.  Round 2 timelines: behind a kernel without shared memory (the gradient exchange,
// the feature scatter) the neighbour search of the next batch (13 KB per CTA) started when that kernel ENDED, not beside
// it; behind the 185 KB weight-gradient CTAs (split 196 KB by default) likewise.  Every kernel of the backward tail,
// the exchange and the search therefore ask for the same split - the largest, 228 KB shared - once before their first
// launch: HPS_SHARE_SM(kernel).
#define HPS_SHARE_SM(kernel)                                                                                      \
    do {                                                                                                          \
        static bool _shared_sm_done = false;                                                                      \
        if (!_shared_sm_done) {                                                                                   \
            HPS_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,              \
                                             cudaSharedmemCarveoutMaxShared));                                    \
            _shared_sm_done = true;                                                                               \
        }                                                                                                         \
    } while (0)

constexpr int kK = HPS_K;   // neighbours per query
constexpr int kC = HPS_C;   // feature channels
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNumSMs = 148;

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Squared distance exactly as the oracle states it: fl(fl(fl(dx*dx)+fl(dy*dy))+fl(dz*dz)),
// dx = fl(q-p); the _rn intrinsics are never contracted into FMAs.
__device__ __forceinline__ float sqdist_rn(float qx, float qy, float qz, float px, float py, float pz) {
    const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// red.global.add.v4.f32 (sm_90+): one 16-byte reduction instead of four scalar ones.
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void red_add(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

}  // namespace hps
