// Adam over the selected (live) feature rows, fused with the write-back into the full table.
//
// Reference: the mapper clones the frustum-selected rows of every feature table into leaf tensors, runs
// torch.optim.Adam on them and writes them back into the tables each iteration (src/Mapper.py:512-536, 592-603,
// 615-630, 859, 862-874).  Here one kernel per table does the Adam update of the compact (M,C) leaves AND stores the
// new rows into the (N,C) table the gather kernels read, so neither the optimiser's foreach kernels nor the
// per-iteration index copy remain.  The step counter lives on the device (CUDA-graph replayable): the kernel
// reads t = step + 1 and the last block to finish stores it.
//
// Update rule = torch.optim.Adam (no amsgrad, no weight decay, maximize=False):
//     m <- m + (g - m)(1 - b1);  v <- b2 v + (1 - b2) g g
//     p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "hps_common.cuh"

namespace hps {

__global__ void __launch_bounds__(256)
adam_rows_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n4, int c4, float lr, float b1, float b2, float eps, int* __restrict__ step,
                 float* __restrict__ table, const int64_t* __restrict__ rows) {
    const int t = step[0] + 1;
    const float bc1 = 1.0f - powf(b1, (float)t), bc2 = 1.0f - powf(b2, (float)t);
    const float step_size = lr / bc1, inv_sqrt_bc2 = 1.0f / sqrtf(bc2);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i], pp = reinterpret_cast<float4*>(p)[i];
        float* gp = const_cast<float*>(reinterpret_cast<const float*>(&gg));
        float *mp = reinterpret_cast<float*>(&mm), *vp = reinterpret_cast<float*>(&vv), *ppp = reinterpret_cast<float*>(&pp);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gr = gp[k];
            mp[k] = mp[k] + (gr - mp[k]) * (1.0f - b1);
            vp[k] = vp[k] * b2 + (1.0f - b2) * gr * gr;
            const float denom = sqrtf(vp[k]) * inv_sqrt_bc2 + eps;
            ppp[k] = ppp[k] - step_size * (mp[k] / denom);
        }
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
        reinterpret_cast<float4*>(p)[i] = pp;
        if (table != nullptr) {
            const int64_t row = i / c4, q = i - row * c4;
            reinterpret_cast<float4*>(table)[rows[row] * c4 + q] = pp;
        }
    }
    // the last block to finish publishes the new step count (every block has read the old one by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int done = atomicAdd(step + 1, 1) + 1;
        if (done == (int)gridDim.x) {
            step[1] = 0;
            step[0] = t;
        }
    }
}

// The same update for a handful of small dense tensors (tracker: pose quaternion / translation, exposure latent and
// exposure MLP, reference src/Tracker.py:383-416) in ONE launch: torch.optim.Adam's capturable foreach path is ~20
// launches of 1-2 us each, more than a tenth of a tracker iteration.  Tensor descriptors travel by value in the
// kernel arguments.  zero_grad != 0: the gradient is cleared after use (the next backward accumulates into it), so
// the caller's optimizer.zero_grad() disappears as well.
constexpr int kAdamMaxTensors = 24;
struct AdamPack {
    float* p[kAdamMaxTensors];
    float* g[kAdamMaxTensors];
    float* m[kAdamMaxTensors];
    float* v[kAdamMaxTensors];
    float lr[kAdamMaxTensors];
    int n[kAdamMaxTensors];
    int block_begin[kAdamMaxTensors + 1];
    int count;
};

__global__ void __launch_bounds__(256)
adam_multi_kernel(const AdamPack pk, float b1, float b2, float eps, int* __restrict__ step, int zero_grad) {
    const int t = step[0] + 1;
    const float bc1 = 1.0f - powf(b1, (float)t), bc2 = 1.0f - powf(b2, (float)t);
    const float inv_sqrt_bc2 = 1.0f / sqrtf(bc2);
    int ti = 0;
    while (ti + 1 < pk.count && (int)blockIdx.x >= pk.block_begin[ti + 1]) ++ti;
    const int i = ((int)blockIdx.x - pk.block_begin[ti]) * blockDim.x + threadIdx.x;
    if (i < pk.n[ti]) {
        const float gr = pk.g[ti][i];
        const float mm = pk.m[ti][i] + (gr - pk.m[ti][i]) * (1.0f - b1);
        const float vv = pk.v[ti][i] * b2 + (1.0f - b2) * gr * gr;
        const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
        pk.m[ti][i] = mm;
        pk.v[ti][i] = vv;
        pk.p[ti][i] = pk.p[ti][i] - (pk.lr[ti] / bc1) * (mm / denom);
        if (zero_grad) pk.g[ti][i] = 0.f;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int done = atomicAdd(step + 1, 1) + 1;
        if (done == (int)gridDim.x) {
            step[1] = 0;
            step[0] = t;
        }
    }
}

// adam_multi with ONE STEP COUNTER PER TENSOR - torch.optim.Adam keeps `state[p]['step']` per parameter and skips
// parameters whose gradient is None, so in the mapper (src/Mapper.py:592-603, 637-660: one optimiser over all four
// decoders, four feature tables, poses and the exposure latent, of which a stage touches only some) every tensor has its
// own bias correction.  steps[i] points at that tensor's counter; `done` is a scratch word (0 between launches).
struct AdamStepPack {
    int* step[kAdamMaxTensors];
};

__global__ void __launch_bounds__(256)
adam_multi_own_steps_kernel(const AdamPack pk, const AdamStepPack sp, float b1, float b2, float eps,
                            int* __restrict__ done_word, int zero_grad) {
    int ti = 0;
    while (ti + 1 < pk.count && (int)blockIdx.x >= pk.block_begin[ti + 1]) ++ti;
    const int t = *sp.step[ti] + 1;
    const float bc1 = 1.0f - powf(b1, (float)t), bc2 = 1.0f - powf(b2, (float)t);
    const float inv_sqrt_bc2 = 1.0f / sqrtf(bc2);
    const int i = ((int)blockIdx.x - pk.block_begin[ti]) * blockDim.x + threadIdx.x;
    if (i < pk.n[ti]) {
        const float gr = pk.g[ti][i];
        const float mm = pk.m[ti][i] + (gr - pk.m[ti][i]) * (1.0f - b1);
        const float vv = pk.v[ti][i] * b2 + (1.0f - b2) * gr * gr;
        const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
        pk.m[ti][i] = mm;
        pk.v[ti][i] = vv;
        pk.p[ti][i] = pk.p[ti][i] - (pk.lr[ti] / bc1) * (mm / denom);
        if (zero_grad) pk.g[ti][i] = 0.f;
    }
    // the last block to finish advances every counter of the launch (all blocks have read theirs by then)
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = (atomicAdd(done_word, 1) + 1 == (int)gridDim.x);
    }
    __syncthreads();
    if (last) {
        if ((int)threadIdx.x < pk.count) *sp.step[threadIdx.x] += 1;
        if (threadIdx.x == 0) *done_word = 0;
    }
}

}  // namespace hps

extern "C" int hps_adam_multi(int n_tensors, float* const* params, float* const* grads, float* const* exp_avg,
                              float* const* exp_avg_sq, const int64_t* numel, const float* lr, float beta1,
                              float beta2, float eps, int32_t* step, int zero_grad, void* stream) {
    using namespace hps;
    HPS_REQUIRE(n_tensors >= 0 && n_tensors <= kAdamMaxTensors, "at most 24 tensors per call");
    if (n_tensors == 0) return 0;
    HPS_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel && lr && step, "null buffer");
    AdamPack pk;
    pk.count = n_tensors;
    int blocks = 0;
    for (int i = 0; i < n_tensors; ++i) {
        HPS_REQUIRE(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i], "null tensor pointer");
        HPS_REQUIRE(numel[i] > 0 && numel[i] < (1ll << 24), "tensor size out of range for the small-tensor Adam");
        pk.p[i] = params[i];
        pk.g[i] = grads[i];
        pk.m[i] = exp_avg[i];
        pk.v[i] = exp_avg_sq[i];
        pk.lr[i] = lr[i];
        pk.n[i] = (int)numel[i];
        pk.block_begin[i] = blocks;
        blocks += (int)ceil_div(numel[i], 256);
    }
    pk.block_begin[n_tensors] = blocks;
    adam_multi_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pk, beta1, beta2, eps, step, zero_grad);
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_adam_rows(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t M, int C,
                             float lr, float beta1, float beta2, float eps, int32_t* step, float* table,
                             const int64_t* rows, void* stream) {
    using namespace hps;
    HPS_REQUIRE(M >= 0 && C > 0 && C % 4 == 0, "need M >= 0 and C a multiple of 4");
    if (M == 0) return 0;
    HPS_REQUIRE(param && grad && exp_avg && exp_avg_sq && step, "null buffer");
    HPS_REQUIRE(table == nullptr || rows != nullptr, "table write-back needs the row indices");
    const int64_t n4 = M * (C / 4);
    adam_rows_kernel<<<(unsigned)ceil_div(n4, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n4,
                                                                                C / 4, lr, beta1, beta2, eps, step, table, rows);
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_adam_multi_own_steps(int n_tensors, float* const* params, float* const* grads, float* const* exp_avg,
                                        float* const* exp_avg_sq, const int64_t* numel, const float* lr,
                                        int32_t* const* steps, float beta1, float beta2, float eps, int32_t* done_word,
                                        int zero_grad, void* stream) {
    using namespace hps;
    HPS_REQUIRE(n_tensors >= 0 && n_tensors <= kAdamMaxTensors, "at most 24 tensors per call");
    if (n_tensors == 0) return 0;
    HPS_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel && lr && steps && done_word, "null buffer");
    AdamPack pk;
    AdamStepPack sp;
    pk.count = n_tensors;
    int blocks = 0;
    for (int i = 0; i < n_tensors; ++i) {
        HPS_REQUIRE(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i] && steps[i], "null tensor pointer");
        HPS_REQUIRE(numel[i] > 0 && numel[i] < (1ll << 24), "tensor size out of range for the small-tensor Adam");
        for (int j = 0; j < i; ++j) HPS_REQUIRE(steps[j] != steps[i], "two tensors share a step counter");
        pk.p[i] = params[i];
        pk.g[i] = grads[i];
        pk.m[i] = exp_avg[i];
        pk.v[i] = exp_avg_sq[i];
        pk.lr[i] = lr[i];
        pk.n[i] = (int)numel[i];
        pk.block_begin[i] = blocks;
        sp.step[i] = steps[i];
        blocks += (int)ceil_div(numel[i], 256);
    }
    pk.block_begin[n_tensors] = blocks;
    adam_multi_own_steps_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pk, sp, beta1, beta2, eps, done_word, zero_grad);
    HPS_LAUNCH_OK();
    return 0;
}
