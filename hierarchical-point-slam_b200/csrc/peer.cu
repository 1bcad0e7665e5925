// Gradient exchange over NVLink peer memory: one kernel per iteration instead of a library collective.
//
// The mapper's exchange step (point-feature + decoder gradients of the ray shards summed over the ranks,
// the data-parallel form of Mapper.py:783-860) moves 3.4 MB per rank and iteration at the bench workload - a payload
// for which a library all-reduce is dominated by its launch and protocol latency, and which (on this stack) could not be
// captured into the iteration's CUDA graph.  Here every rank's gradient buffer lives in SYMMETRIC memory (same size,
// peer-mapped into every rank's address space, optionally bound to an NVLS multicast address); the fused backward writes
// its gradients straight into it (RenderSpec.grad_buffer), and this kernel sums the buffers in place:
//
//   barrier   (all ranks' gradients are complete): CTA b of rank r raises flag [b][r] in every peer's signal pad and
//             waits for flag [b][p] of every peer p in its own pad
//   reduce    rank r owns slice r of the buffer.  For every 16-byte vector of the slice it sums the W copies in rank
//             order - P2P: W vector loads over NVLink; multicast: ONE multimem.ld_reduce, the NVSwitch adds - and
//   broadcast stores the sum into all W copies - P2P: W vector stores; multicast: one multimem.st -
//   barrier   (every rank's buffer is complete) before the kernel ends.
//
// The owner computes each sum once and everybody receives those bits, so all ranks hold bit-identical gradients
// (replicated optimiser state stays in sync without a broadcast).  The flags are monotone epoch numbers kept on the
// device, so the kernel carries no epoch argument and replays from a CUDA graph.  48 CTAs: the kernel is bound by the
// latency of its two barriers, not by bandwidth (3.4 MB), and it is meant to leave the SMs to the kernel beside it.
#include "hps_common.cuh"

namespace hps {

constexpr int kPeerMaxWorld = 8;
constexpr int kPeerThreads = 256;

struct PeerPtrs {
    float* buf[kPeerMaxWorld];
    uint32_t* pad[kPeerMaxWorld];
};

__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long kPeerTimeoutNs = 2000000000ull;   // a peer that never arrives must not hang the GPU

// Flags are monotone epoch numbers: barrier number e of CTA b (two per launch, counted in a word of this rank's own pad,
// so the kernel carries no epoch argument and replays from a CUDA graph) is "every peer's flag [b][p] >= e".  Raising a
// flag is ONE posted store to the peer (st.release.sys: everything this CTA wrote - and, through the CTA barrier in
// front of it, everything its threads observed - is visible system-wide first); waiting spins on LOCAL memory.  No
// read-modify-write crosses NVLink: a barrier costs one one-way hop.
__device__ __forceinline__ void put_flag(uint32_t* addr, uint32_t epoch) {
    asm volatile("st.global.release.sys.b32 [%0], %1;" ::"l"(addr), "r"(epoch) : "memory");
}
// gives up after kPeerTimeoutNs and raises *status (bit 1): the sums are then garbage and PeerAllReduce.check() reports
// it - a rank that died or took another code path costs an error message, not a wedged device
__device__ __forceinline__ void wait_flag(const uint32_t* addr, uint32_t epoch, uint32_t* status) {
    uint32_t v;
    const unsigned long long t0 = now_ns();
    while (true) {
        asm volatile("ld.global.acquire.sys.b32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
        if ((int32_t)(v - epoch) >= 0) return;              // (wrap-around safe)
        if (now_ns() - t0 > kPeerTimeoutNs) {
            atomicOr(status, 2u);
            return;
        }
        __nanosleep(40);
    }
}

// CTA-level barrier across the ranks: thread p < world talks to peer p.  Flag [cta][r] of rank p's pad is written by rank r.
__device__ __forceinline__ void rank_barrier(const PeerPtrs& pp, int rank, int world, int cta, uint32_t epoch,
                                             uint32_t* status) {
    __threadfence_system();      // every thread's own stores first, then the CTA barrier, then one flag per peer
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const int p = threadIdx.x;
        put_flag(pp.pad[p] + (size_t)cta * kPeerMaxWorld + rank, epoch);
        wait_flag(pp.pad[rank] + (size_t)cta * kPeerMaxWorld + p, epoch, status);
    }
    __syncthreads();
}

__device__ __forceinline__ float4 ld_peer(const float* p) {
    float4 v;
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer(float* p, float4 v) {
    asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_reduce_mc(const float* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_mc(float* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// n4 = number of float4 vectors in the buffer; slice of rank r = vectors [r * per, min((r + 1) * per, n4)).
// A thread keeps 4 (multicast) / 2 x world (P2P) vectors in flight (the reduce phase is bound by the NVLink round trip, ~2 us: 3.4 MB need
// ~1.8 MB in flight to fill the link).
constexpr int kPeerUnrollMc = 4, kPeerUnrollP2p = 2;

template <bool MULTICAST>
__global__ void __launch_bounds__(kPeerThreads)
peer_allreduce_kernel(PeerPtrs pp, float* mc, int rank, int world, int64_t n4) {
    const int cta = blockIdx.x;
    uint32_t* status = pp.pad[rank] + (size_t)gridDim.x * kPeerMaxWorld;      // the word behind this rank's flags
    // this CTA's launch counter lives behind the status word (local memory, touched by this CTA only)
    __shared__ uint32_t s_epoch;
    if (threadIdx.x == 0) {
        uint32_t* counter = status + 4 + cta;
        s_epoch = *counter + 2u;
        *counter = s_epoch;
    }
    __syncthreads();
    const uint32_t epoch = s_epoch;                                           // barriers epoch - 1 and epoch of this launch
    rank_barrier(pp, rank, world, cta, epoch - 1u, status);
    const int64_t per = (n4 + world - 1) / world;
    const int64_t lo = (int64_t)rank * per, hi = lo + per < n4 ? lo + per : n4;
    const int64_t stride = (int64_t)gridDim.x * kPeerThreads;
    constexpr int kPeerUnroll = MULTICAST ? kPeerUnrollMc : kPeerUnrollP2p;
    for (int64_t i0 = lo + (int64_t)cta * kPeerThreads + threadIdx.x; i0 < hi; i0 += stride * kPeerUnroll) {
        float4 acc[kPeerUnroll];
        if constexpr (MULTICAST) {
#pragma unroll
            for (int u = 0; u < kPeerUnroll; ++u)
                if (i0 + u * stride < hi) acc[u] = ld_reduce_mc(mc + 4 * (i0 + u * stride));
#pragma unroll
            for (int u = 0; u < kPeerUnroll; ++u)
                if (i0 + u * stride < hi) st_mc(mc + 4 * (i0 + u * stride), acc[u]);
        } else {
            float4 v[kPeerUnroll][kPeerMaxWorld];
#pragma unroll
            for (int u = 0; u < kPeerUnroll; ++u)
#pragma unroll
                for (int p = 0; p < kPeerMaxWorld; ++p)
                    if (p < world && i0 + u * stride < hi) v[u][p] = ld_peer(pp.buf[p] + 4 * (i0 + u * stride));
#pragma unroll
            for (int u = 0; u < kPeerUnroll; ++u) {
                if (i0 + u * stride < hi) {
                    acc[u] = v[u][0];
#pragma unroll
                    for (int p = 1; p < kPeerMaxWorld; ++p) {          // rank order: every owner adds in the same order
                        if (p < world) {
                            acc[u].x += v[u][p].x; acc[u].y += v[u][p].y; acc[u].z += v[u][p].z; acc[u].w += v[u][p].w;
                        }
                    }
#pragma unroll
                    for (int p = 0; p < kPeerMaxWorld; ++p)
                        if (p < world) st_peer(pp.buf[p] + 4 * (i0 + u * stride), acc[u]);
                }
            }
        }
    }
    rank_barrier(pp, rank, world, cta, epoch, status);
}

}  // namespace hps

extern "C" int hps_peer_allreduce_ctas(void) { return 48; }

// flags [ctas][8] | status word + 3 pad | per-CTA launch counters [ctas]
extern "C" int64_t hps_peer_allreduce_pad_words(void) { return (int64_t)hps_peer_allreduce_ctas() * (hps::kPeerMaxWorld + 1) + 4; }

extern "C" int hps_peer_allreduce(float* const* bufs, uint32_t* const* pads, float* multicast, int rank, int world,
                                  int64_t n_floats, void* stream) {
    using namespace hps;
    HPS_REQUIRE(world >= 1 && world <= kPeerMaxWorld, "world size must be 1..8");
    HPS_REQUIRE(rank >= 0 && rank < world, "rank out of range");
    HPS_REQUIRE(n_floats >= 0 && n_floats % 4 == 0, "n_floats must be a multiple of 4");
    if (n_floats == 0 || world == 1) return 0;
    HPS_REQUIRE(bufs && pads, "null pointer table");
    PeerPtrs pp;
    for (int p = 0; p < kPeerMaxWorld; ++p) {
        pp.buf[p] = p < world ? bufs[p] : nullptr;
        pp.pad[p] = p < world ? pads[p] : nullptr;
        if (p < world) {
            HPS_REQUIRE(bufs[p] != nullptr && pads[p] != nullptr, "null peer pointer");
            HPS_REQUIRE((reinterpret_cast<uintptr_t>(bufs[p]) & 15) == 0, "buffers must be 16-byte aligned");
        }
    }
    HPS_REQUIRE(multicast == nullptr || (reinterpret_cast<uintptr_t>(multicast) & 15) == 0, "multicast address must be 16-byte aligned");
    HPS_SHARE_SM(peer_allreduce_kernel<true>);
    HPS_SHARE_SM(peer_allreduce_kernel<false>);
    const int ctas = hps_peer_allreduce_ctas();
    if (multicast != nullptr)
        peer_allreduce_kernel<true><<<ctas, kPeerThreads, 0, (cudaStream_t)stream>>>(pp, multicast, rank, world, n_floats / 4);
    else
        peer_allreduce_kernel<false><<<ctas, kPeerThreads, 0, (cudaStream_t)stream>>>(pp, nullptr, rank, world, n_floats / 4);
    HPS_LAUNCH_OK();
    return 0;
}
