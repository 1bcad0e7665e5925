// Hashed voxel grid + exact radius-limited 8-NN (kernel 1 of the north star).
//
// Replaces faiss.IndexIVFFlat train/add/search as used by NeuralPointCloud.find_neighbors_faiss
// (reference src/neural_point.py:211-214, 322-327, 366-417).  The reference's index is approximate
// (nprobe 4 of 400 lists); this one is exact, i.e. equals the reference forced to exhaustive search on
// every slot that lies inside the query radius (the only slots that ever carry weight,
// decoder.py:235-236).
//
// Data layout in HBM (per level):
//   sorted[n]   float4 (x, y, z, original index as int bits), grouped by voxel (cell-sorted)
//   table[T]    16-byte open-addressing hash slots {u64 key, i32 start, i32 count}, T = pow2 >= 2n
//   meta        origin (bbox min), max cell coordinate per axis, status word, #occupied cells
// Build = bbox reduce -> hash insert + per-cell rank -> warp-aggregated range allocation -> scatter.
// Query = one warp per query: lanes probe the voxels overlapping the query ball's bounding cube in
// parallel, the (start,count) ranges are staged in shared memory, then the warp sweeps the flattened
// candidate list with coalesced float4 loads and keeps the running top-8 in lanes 0..7
// (insertion by ballot + shfl_up, ordered by (D, original index)).
#include "hps_common.cuh"
#include "sample_point.cuh"

namespace hps {

struct alignas(16) Cell {
    unsigned long long key;
    int start;
    int count;
};

struct GridMeta {
    int min_enc[3];
    int max_enc[3];
    float origin[3];
    int maxc[3];
    int minc[3];
    int status;     // bit 0: a point outside the 21-bit cell range; appends: bit 2 `sorted` is full (the index is unusable
                    // until it is rebuilt)
    int n_cells;
    int cursor;
};

// The grid's origin sits kMarginCells voxels below the bounding box of the points it was built over, so that points
// appended later (hps_grid_append: a SLAM map grows in every direction) still get non-negative cell coordinates.
constexpr int kMarginCells = 1024;
// A cell's range in `sorted` is allocated with room to grow: count + max(2, count / 2) entries.
__device__ __forceinline__ int cell_capacity(int count) { return count + max(2, count >> 1); }

constexpr unsigned long long kEmptyKey = ~0ull;
constexpr int kCoordBits = 21;
constexpr int kCoordMax = (1 << kCoordBits) - 1;

__device__ __forceinline__ int float_enc(float f) {
    int b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float float_dec(int e) { return __int_as_float(e >= 0 ? e : e ^ 0x7fffffff); }

// The one and only point->cell-coordinate map (monotone non-decreasing in x for inv > 0).
__device__ __forceinline__ int cell_coord(float x, float o, float inv) {
    return __float2int_rd(__fmul_rn(__fsub_rn(x, o), inv));
}
__device__ __forceinline__ unsigned long long cell_key(int cx, int cy, int cz) {
    return (unsigned long long)cx | ((unsigned long long)cy << kCoordBits) |
           ((unsigned long long)cz << (2 * kCoordBits));
}
__device__ __forceinline__ unsigned hash_key(unsigned long long k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (unsigned)k;
}

__global__ void grid_init_kernel(Cell* table, int64_t T, GridMeta* meta) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < T) {
        Cell c;
        c.key = kEmptyKey;
        c.start = 0;
        c.count = 0;
        table[i] = c;
    }
    if (i == 0) {
        for (int a = 0; a < 3; ++a) {
            meta->min_enc[a] = 0x7fffffff;
            meta->max_enc[a] = (int)0x80000000;
            meta->origin[a] = 0.f;
            meta->maxc[a] = -1;
            meta->minc[a] = 0;
        }
        meta->status = 0;
        meta->n_cells = 0;
        meta->cursor = 0;
    }
}

__global__ void grid_bbox_kernel(const float* __restrict__ pos, int64_t n, GridMeta* meta) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float v = pos[i * 3 + a];
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(kFull, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(kFull, mx[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&meta->min_enc[a], float_enc(mn[a]));
            atomicMax(&meta->max_enc[a], float_enc(mx[a]));
        }
    }
}

__global__ void grid_origin_kernel(GridMeta* meta, float inv_cell) {
    if (threadIdx.x < 3) {
        int a = threadIdx.x;
        const float lo = float_dec(meta->min_enc[a]);
        float m = float_dec(meta->max_enc[a]);
        const float o = lo - (float)kMarginCells / inv_cell;
        meta->origin[a] = o;
        int c = cell_coord(m, o, inv_cell);
        if (c > kCoordMax || c < 0) {
            atomicOr(&meta->status, 1);
            c = kCoordMax;
        }
        meta->maxc[a] = c;
        meta->minc[a] = max(cell_coord(lo, o, inv_cell), 0);
    }
}

__global__ void grid_insert_kernel(const float* __restrict__ pos, int64_t n, float inv_cell, Cell* table,
                                   unsigned mask, GridMeta* meta, int* __restrict__ pt_slot,
                                   int* __restrict__ pt_rank) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        int v = cell_coord(pos[i * 3 + a], meta->origin[a], inv_cell);
        if (v < 0 || v > kCoordMax) {
            atomicOr(&meta->status, 1);
            v = min(max(v, 0), kCoordMax);
        }
        c[a] = v;
    }
    const unsigned long long key = cell_key(c[0], c[1], c[2]);
    unsigned h = hash_key(key) & mask;
    while (true) {
        unsigned long long prev = atomicCAS(&table[h].key, kEmptyKey, key);
        if (prev == kEmptyKey) {
            atomicAdd(&meta->n_cells, 1);
            break;
        }
        if (prev == key) break;
        h = (h + 1) & mask;
    }
    pt_slot[i] = (int)h;
    pt_rank[i] = atomicAdd(&table[h].count, 1);
}

// Give every occupied cell a contiguous range of `sorted`: warp-aggregated bump allocation (the
// order of cells in memory is irrelevant to the search result).
__global__ void grid_alloc_kernel(Cell* table, int64_t T, GridMeta* meta, int* __restrict__ cell_cap) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int count = (i < T) ? table[i].count : 0;
    int cnt = count > 0 ? cell_capacity(count) : 0;          // entries reserved for the cell (room for appends)
    if (i < T) cell_cap[i] = cnt;
    int incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
    }
    int total = __shfl_sync(kFull, incl, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(&meta->cursor, total);
    base = __shfl_sync(kFull, base, 31);
    if (i < T && cnt > 0) table[i].start = base + incl - cnt;
}

__global__ void grid_scatter_kernel(const float* __restrict__ pos, int64_t n, const Cell* __restrict__ table,
                                    const int* __restrict__ pt_slot, const int* __restrict__ pt_rank,
                                    float4* __restrict__ sorted) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int dst = table[pt_slot[i]].start + pt_rank[i];
    sorted[dst] = make_float4(pos[i * 3 + 0], pos[i * 3 + 1], pos[i * 3 + 2], __int_as_float((int)i));
}

// ------------------------------------------------------------------------------------------------
// append (index.add on an existing index, neural_point.py:327): hash ONLY the new points into the built structure
// ------------------------------------------------------------------------------------------------
// New point i (global index n_old + i): find or create its cell, take the next rank in it.
__global__ void grid_append_insert_kernel(const float* __restrict__ pos, int64_t n_old, int64_t n_add, float inv_cell,
                                          Cell* table, unsigned mask, GridMeta* meta, int* __restrict__ pt_slot,
                                          int* __restrict__ pt_rank) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_add) return;
    const float* p = pos + (n_old + i) * 3;
    int c[3];
    bool ok = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        c[a] = cell_coord(p[a], meta->origin[a], inv_cell);
        ok = ok && c[a] >= 0 && c[a] <= kCoordMax;
    }
    if (!ok) {                      // outside the coordinate range of this origin: the host rebuilds
        atomicOr(&meta->status, 1);
        pt_slot[i] = -1;
        return;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        atomicMax(&meta->maxc[a], c[a]);
        atomicMin(&meta->minc[a], c[a]);
    }
    const unsigned long long key = cell_key(c[0], c[1], c[2]);
    unsigned h = hash_key(key) & mask;
    while (true) {
        unsigned long long prev = atomicCAS(&table[h].key, kEmptyKey, key);
        if (prev == kEmptyKey) {
            atomicAdd(&meta->n_cells, 1);
            break;
        }
        if (prev == key) break;
        h = (h + 1) & mask;
    }
    pt_slot[i] = (int)h;
    pt_rank[i] = atomicAdd(&table[h].count, 1);
}

// The last arrival of a cell (rank == count - 1) settles its range: a new cell gets one at the cursor; a cell that
// outgrew its range MOVES to a larger one there (the old entries are copied, the old range becomes dead space until the
// next rebuild compacts it - frontier cells of a growing map are sparse when first seen and fill up a frame later).  Only
// a full `sorted` array is reported (status bit 2): the host rebuilds.
__global__ void grid_append_alloc_kernel(int64_t n_add, Cell* table, int* __restrict__ cell_cap, GridMeta* meta,
                                         const int* __restrict__ pt_slot, const int* __restrict__ pt_rank,
                                         float4* __restrict__ sorted, int64_t sorted_cap) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_add) return;
    const int h = pt_slot[i];
    if (h < 0) return;
    const int cnt = table[h].count;
    const int old_cap = cell_cap[h];
    if (pt_rank[i] != cnt - 1 || cnt <= old_cap) return;
    const int want = max(cell_capacity(cnt), 2 * old_cap);
    const int start = atomicAdd(&meta->cursor, want);
    if ((int64_t)start + want > sorted_cap) {
        atomicOr(&meta->status, 4);
        return;
    }
    // entries [0, old count) are the cell's old points; [old count, count) belong to this append and are written by the
    // scatter pass; copying min(old_cap, count) entries moves every old point (and some slack that gets overwritten)
    const int old_start = table[h].start;
    const int n_copy = min(old_cap, cnt);
    for (int j = 0; j < n_copy; ++j) sorted[start + j] = sorted[old_start + j];
    table[h].start = start;
    cell_cap[h] = want;
}

__global__ void grid_append_scatter_kernel(const float* __restrict__ pos, int64_t n_old, int64_t n_add,
                                           const Cell* __restrict__ table, const int* __restrict__ cell_cap,
                                           const int* __restrict__ pt_slot, const int* __restrict__ pt_rank,
                                           float4* __restrict__ sorted) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_add) return;
    const int h = pt_slot[i];
    if (h < 0 || pt_rank[i] >= cell_cap[h]) return;          // (reported by the alloc pass)
    const float* p = pos + (n_old + i) * 3;
    sorted[table[h].start + pt_rank[i]] = make_float4(p[0], p[1], p[2], __int_as_float((int)(n_old + i)));
}

// ------------------------------------------------------------------------------------------------
// query
// ------------------------------------------------------------------------------------------------
constexpr int kQueryWarps = 8;     // warps (= queries in flight) per CTA
// Shared memory and registers of the search CTA are sized for what runs BESIDE it: the mapper iteration issues the next
// batch's search beside this batch's backward tail, where a weight-gradient CTA holds 186 KB of the SM's 228 KB of shared
// memory and 18 K of its 64 K registers.  64 staged voxel ranges per pass and warp (9 KB per CTA; a fine-level ball's
// bounding cube is 27-64 voxels, the mid level's up to 125 = two passes) and 40 registers (launch bound 6 CTAs per SM,
// 40 bytes of spills) put FOUR search CTAs beside a weight-gradient CTA instead of three: the iteration 0.2515 -> 0.2475
// ms, the search alone 44.8 -> 46.0 us, the mid level unchanged (163-167 us); 128 ranges / 5 CTAs: -DHPS_KNN_STAGE_CELLS=128
// -DHPS_KNN_CTAS_PER_SM=5.
#ifndef HPS_KNN_STAGE_CELLS
#define HPS_KNN_STAGE_CELLS 64
#endif
#ifndef HPS_KNN_CTAS_PER_SM
#define HPS_KNN_CTAS_PER_SM 6
#endif
constexpr int kStageCells = HPS_KNN_STAGE_CELLS;

struct GridView {
    const float4* sorted;
    const Cell* table;
    const GridMeta* meta;
    unsigned mask;
    float inv_cell;
};

// A candidate is the 64-bit key (bits(D) << 32) | index: squared distances are non-negative, so unsigned key
// order == the reference's (D, index) order and one compare replaces the two-level test.  ~0 = empty slot.
constexpr unsigned long long kNoKey = ~0ull;

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
    return ((unsigned long long)__shfl_xor_sync(kFull, (unsigned)(v >> 32), m) << 32) | __shfl_xor_sync(kFull, (unsigned)v, m);
}
__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, int src) {
    return ((unsigned long long)__shfl_sync(kFull, (unsigned)(v >> 32), src) << 32) | __shfl_sync(kFull, (unsigned)v, src);
}

// ascending bitonic sort of one key per lane (15 compare-exchange steps)
__device__ __forceinline__ unsigned long long warp_sort32(unsigned long long v, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const unsigned long long o = shfl_xor_u64(v, j);
            const bool keep_min = ((lane & j) == 0) == ((lane & k) == 0);
            v = (keep_min == (v < o)) ? v : o;
        }
    }
    return v;
}

// best: ascending top-8 in lanes 0..7 (other lanes kNoKey); blk: ascending in lanes 0..31.  Returns the merged top-8.
__device__ __forceinline__ unsigned long long merge_top8(unsigned long long best, unsigned long long blk, int lane) {
    const unsigned long long rev = shfl_u64(blk, 15 - (lane & 15));     // lanes 8..15 <- blk[7..0]
    unsigned long long v = lane < 8 ? best : rev;                        // lanes 0..15: bitonic sequence
#pragma unroll
    for (int j = 8; j > 0; j >>= 1) {
        const unsigned long long o = shfl_xor_u64(v, j);
        v = (((lane & j) == 0) == (v < o)) ? v : o;
    }
    return lane < kK ? v : kNoKey;
}

// One warp per GROUP of G consecutive queries (G = 1: a single query; G = 5: the samples of one ray, whose search
// balls overlap heavily, so the voxel probes, the candidate loads and the owner look-ups are shared by the group).
// (1) the voxels overlapping the bounding box of the group's search balls are looked up in the hash table, one per
// lane; the non-empty ones are compacted into shared memory as (first point, exclusive prefix of the counts);
// (2) the concatenation of their point ranges is swept 32 candidates at a time - the owner voxel of every candidate
// comes from a start-bit mask (redux.or + popc), not from a search; (3) per query, candidates inside the radius and
// better than the current 8th are appended to a per-warp buffer, and every 32 of them are sorted by a bitonic
// network and merged into the running top-8.  Results do not depend on G (exact search).
// RAYS: the queries are the sample points of rays and are computed here (sample_point.cuh) instead of being read -
// the warp also writes z / point / thresholds for the kernels downstream, which saves the sampling launch.
template <int G, bool RAYS>
__global__ void __launch_bounds__(kQueryWarps * 32, G == 1 ? HPS_KNN_CTAS_PER_SM : 4)      // G = 1: 40 registers (see kStageCells above); G = 5 would take 80 otherwise
knn_query_kernel(GridView g, const float* __restrict__ q, const float* __restrict__ t_up,
                 const float* __restrict__ t_dn, int64_t nq, int32_t* __restrict__ I, float* __restrict__ D,
                 int32_t* __restrict__ count, int32_t* __restrict__ cand, const RaySource src,
                 float* __restrict__ z_out, float* __restrict__ pts_out, float* __restrict__ tup_out,
                 float* __restrict__ tdn_out) {
    __shared__ int s_start[kQueryWarps][kStageCells];
    __shared__ int s_excl[kQueryWarps][kStageCells + 32];
    __shared__ unsigned long long s_buf[kQueryWarps][G][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q0 = ((int64_t)blockIdx.x * kQueryWarps + warp) * G;
    if (q0 >= nq) return;

    const float ox = g.meta->origin[0], oy = g.meta->origin[1], oz = g.meta->origin[2];
    float qx[G], qy[G], qz[G], tdn[G];
    float tup[RAYS ? G : 1];
    int x0 = 0x7fffffff, y0 = 0x7fffffff, z0 = 0x7fffffff, x1 = -1, y1 = -1, z1 = -1;
#pragma unroll
    for (int k = 0; k < G; ++k) {
        const bool live = q0 + k < nq;
        const int64_t qi = live ? q0 + k : q0;
        if constexpr (RAYS) {
            const int64_t r = qi / src.S;
            const SamplePoint sp = sample_point(src, r, (int)(qi - r * src.S));
            qx[k] = sp.x;
            qy[k] = sp.y;
            qz[k] = sp.zc;
            tdn[k] = live ? sp.t_dn : -1.0f;
            tup[k] = sp.t_up;
            if (live && lane == 0) {
                z_out[qi] = sp.z;
                pts_out[qi * 3 + 0] = sp.x;
                pts_out[qi * 3 + 1] = sp.y;
                pts_out[qi * 3 + 2] = sp.zc;
                tup_out[qi] = sp.t_up;
                tdn_out[qi] = sp.t_dn;
            }
        } else {
            qx[k] = __ldg(q + qi * 3 + 0);
            qy[k] = __ldg(q + qi * 3 + 1);
            qz[k] = __ldg(q + qi * 3 + 2);
            tdn[k] = live ? __ldg(t_dn + qi) : -1.0f;
        }
        if (tdn[k] >= 0.f) {
            // conservative search radius: sqrt(t_dn) plus slack that dominates every fp32 rounding involved
            const float rr = sqrtf(tdn[k]) * 1.00001f + 1e-6f * (fabsf(qx[k]) + fabsf(qy[k]) + fabsf(qz[k]) + 1.0f);
            x0 = min(x0, cell_coord(qx[k] - rr, ox, g.inv_cell));
            x1 = max(x1, cell_coord(qx[k] + rr, ox, g.inv_cell));
            y0 = min(y0, cell_coord(qy[k] - rr, oy, g.inv_cell));
            y1 = max(y1, cell_coord(qy[k] + rr, oy, g.inv_cell));
            z0 = min(z0, cell_coord(qz[k] - rr, oz, g.inv_cell));
            z1 = max(z1, cell_coord(qz[k] + rr, oz, g.inv_cell));
        }
    }
    x0 = max(x0, g.meta->minc[0]); y0 = max(y0, g.meta->minc[1]); z0 = max(z0, g.meta->minc[2]);
    x1 = min(x1, g.meta->maxc[0]); y1 = min(y1, g.meta->maxc[1]); z1 = min(z1, g.meta->maxc[2]);
    const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, nz = z1 - z0 + 1;
    const bool any = (nx > 0) && (ny > 0) && (nz > 0);
    const int64_t ncell = any ? (int64_t)nx * ny * nz : 0;

    unsigned long long best[G];   // running top-8 of query k in lanes 0..7, ascending
    unsigned long long thr[G];    // a candidate is interesting for query k iff key < thr[k]
    int nbuf[G];                  // warp-uniform fill of s_buf[warp][k]
#pragma unroll
    for (int k = 0; k < G; ++k) {
        best[k] = kNoKey;
        // D <= t_dn  <=>  key < (bits(t_dn) + 1) << 32 ; dead queries accept nothing
        thr[k] = tdn[k] >= 0.f ? ((unsigned long long)__float_as_uint(tdn[k]) + 1ull) << 32 : 0ull;
        nbuf[k] = 0;
    }
    int examined = 0;

    int* my_start = s_start[warp];
    int* my_excl = s_excl[warp];

    for (int64_t pass0 = 0; pass0 < ncell; pass0 += kStageCells) {
        const int npass = (int)min((int64_t)kStageCells, ncell - pass0);
        const int nchunk = (npass + 31) >> 5;
        int running = 0, nne = 0;              // candidates / non-empty voxels staged so far
        for (int t = 0; t < nchunk; ++t) {
            const int64_t ci = pass0 + t * 32 + lane;
            int cnt = 0, st = 0;
            if (ci < pass0 + npass) {
                int ix, iy, iz;
                if (ncell <= 0x7fffffff) {     // the usual case: 32-bit index arithmetic
                    const unsigned c32 = (unsigned)ci, r2 = c32 / (unsigned)nx;
                    ix = x0 + (int)(c32 - r2 * (unsigned)nx);
                    const unsigned r3 = r2 / (unsigned)ny;
                    iy = y0 + (int)(r2 - r3 * (unsigned)ny);
                    iz = z0 + (int)r3;
                } else {
                    ix = x0 + (int)(ci % nx);
                    const int64_t r2 = ci / nx;
                    iy = y0 + (int)(r2 % ny);
                    iz = z0 + (int)(r2 / ny);
                }
                const unsigned long long key = cell_key(ix, iy, iz);
                unsigned h = hash_key(key) & g.mask;
                while (true) {
                    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(g.table + h));
                    const unsigned long long ck = (unsigned long long)raw.x | ((unsigned long long)raw.y << 32);
                    if (ck == key) {
                        st = (int)raw.z;
                        cnt = (int)raw.w;
                        break;
                    }
                    if (ck == kEmptyKey) break;
                    h = (h + 1) & g.mask;
                }
            }
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += v;
            }
            const unsigned ne = __ballot_sync(kFull, cnt > 0);
            if (cnt > 0) {
                const int r = nne + __popc(ne & ((1u << lane) - 1));
                my_start[r] = st;
                my_excl[r] = running + incl - cnt;
            }
            nne += __popc(ne);
            running += __shfl_sync(kFull, incl, 31);
        }
        my_excl[nne + lane] = 0x7fffffff;      // sentinels for the start-mask look-ahead
        __syncwarp();
        const int total = running;
        examined += total;
        int R = 0;                             // non-empty voxels whose range starts before the window
        for (int t0 = 0; t0 < total; t0 += 32) {
            const int t = t0 + lane;
            // voxels starting inside [t0, t0+32): consecutive compact ranks R, R+1, ... (each holds >= 1 point)
            const unsigned rel = (unsigned)(my_excl[R + lane] - t0);
            const unsigned starts = __reduce_or_sync(kFull, rel < 32u ? (1u << rel) : 0u);
            const int owner = R - 1 + __popc(starts & (0xffffffffu >> (31 - lane)));
            R += __popc(starts);
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            const bool have = t < total;
            if (have) p = __ldg(g.sorted + (my_start[owner] + (t - my_excl[owner])));
#pragma unroll
            for (int k = 0; k < G; ++k) {
                unsigned long long key = kNoKey;
                if (have) {
                    const float d = sqdist_rn(qx[k], qy[k], qz[k], p.x, p.y, p.z);
                    key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(p.w);
                }
                const bool want = key < thr[k];
                const unsigned m = __ballot_sync(kFull, want);
                if (m) {
                    unsigned long long* buf = s_buf[warp][k];
                    if (want) buf[nbuf[k] + __popc(m & ((1u << lane) - 1))] = key;
                    nbuf[k] += __popc(m);
                    __syncwarp();
                    if (nbuf[k] >= 32) {
                        const unsigned long long blk = buf[lane];
                        const unsigned long long rest = buf[32 + lane];
                        __syncwarp();
                        nbuf[k] -= 32;
                        if (lane < nbuf[k]) buf[lane] = rest;
                        __syncwarp();
                        best[k] = merge_top8(best[k], warp_sort32(blk, lane), lane);
                        const unsigned long long tau = shfl_u64(best[k], kK - 1);
                        thr[k] = tau < thr[k] ? tau : thr[k];
                    }
                }
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < G; ++k) {
        if (q0 + k >= nq) break;
        if (nbuf[k] > 0) best[k] = merge_top8(best[k], warp_sort32(lane < nbuf[k] ? s_buf[warp][k][lane] : kNoKey, lane), lane);
        const int64_t qi = q0 + k;
        const bool used = best[k] != kNoKey;
        const float bD = __uint_as_float((unsigned)(best[k] >> 32));
        float up;
        if constexpr (RAYS) up = tup[k];
        else up = __ldg(t_up + qi);
        const unsigned inr = __ballot_sync(kFull, (lane < kK) && used && (bD < up));
        if (lane < kK) {
            I[qi * kK + lane] = used ? (int)(unsigned)best[k] : -1;
            D[qi * kK + lane] = used ? bD : FLT_MAX;
        }
        if (lane == 0) {
            count[qi] = __popc(inr);
            if (cand) cand[qi] = examined;
        }
    }
}


}  // namespace hps

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
struct hps_grid {
    int64_t n = 0, cap_n = 0, T = 0, cap_T = 0;
    int64_t sorted_cap = 0;      // entries of `sorted` (cells keep room to grow: <= 3.5 n at build, plus room for appends)
    float cell = 0.f;
    float4* sorted = nullptr;
    hps::Cell* table = nullptr;
    int* cell_cap = nullptr;     // [cap_T] entries reserved per cell (build / append only; the query never reads it)
    int* pt_slot = nullptr;
    int* pt_rank = nullptr;
    hps::GridMeta* meta = nullptr;
    int device = -1;
};

extern "C" int hps_grid_create(hps_grid_t** out) {
    HPS_REQUIRE(out != nullptr, "null out pointer");
    hps_grid* g = new hps_grid();
    HPS_CUDA_OK(cudaGetDevice(&g->device));
    cudaError_t e = cudaMalloc(&g->meta, sizeof(hps::GridMeta));
    if (e != cudaSuccess) {
        delete g;
        hps::set_error("cudaMalloc(meta) failed: %s", cudaGetErrorString(e));
        return -2;
    }
    *out = g;
    return 0;
}

extern "C" int hps_grid_destroy(hps_grid_t* g) {
    if (!g) return 0;
    cudaFree(g->sorted);
    cudaFree(g->table);
    cudaFree(g->cell_cap);
    cudaFree(g->pt_slot);
    cudaFree(g->pt_rank);
    cudaFree(g->meta);
    delete g;
    return 0;
}

extern "C" int hps_grid_build(hps_grid_t* g, const float* pos, int64_t n, float cell_size, void* stream_) {
    using namespace hps;
    HPS_REQUIRE(g != nullptr, "null grid");
    HPS_REQUIRE(n >= 0 && n < (1ll << 30), "point count out of range");
    HPS_REQUIRE(cell_size > 0.f, "cell_size must be positive");
    HPS_REQUIRE(n == 0 || pos != nullptr, "null positions");
    cudaStream_t stream = (cudaStream_t)stream_;
    // capacities: room for half as many points again before an append has to give up (status 8) and the host rebuilds
    const int64_t want_cap = n + n / 2 + 1024;
    int64_t T = 1024;
    while (T < 2 * want_cap) T <<= 1;
    if (want_cap > g->cap_n) {
        HPS_CUDA_OK(cudaStreamSynchronize(stream));
        cudaFree(g->sorted);
        cudaFree(g->pt_slot);
        cudaFree(g->pt_rank);
        g->sorted = nullptr; g->pt_slot = nullptr; g->pt_rank = nullptr; g->cap_n = 0;
        int64_t cap = want_cap;
        // sum of the cells' reserved ranges <= n + n / 2 + 2 * cells <= 3.5 n; appended points need at most 3 entries each
        // when every one of them opens a cell (capacity(1) = 3)
        g->sorted_cap = 3 * n + n / 2 + 3 * (cap - n) + 65536;
        HPS_CUDA_OK(cudaMalloc(&g->sorted, g->sorted_cap * sizeof(float4)));
        HPS_CUDA_OK(cudaMalloc(&g->pt_slot, cap * sizeof(int)));
        HPS_CUDA_OK(cudaMalloc(&g->pt_rank, cap * sizeof(int)));
        g->cap_n = cap;
    }
    if (T > g->cap_T) {
        HPS_CUDA_OK(cudaStreamSynchronize(stream));
        cudaFree(g->table);
        cudaFree(g->cell_cap);
        g->table = nullptr; g->cell_cap = nullptr; g->cap_T = 0;
        HPS_CUDA_OK(cudaMalloc(&g->table, T * sizeof(Cell)));
        HPS_CUDA_OK(cudaMalloc(&g->cell_cap, T * sizeof(int)));
        g->cap_T = T;
    }
    g->n = n;
    g->T = T;
    g->cell = cell_size;
    const float inv = 1.0f / cell_size;
    const int tb = 256;
    grid_init_kernel<<<(unsigned)ceil_div(T, tb), tb, 0, stream>>>(g->table, T, g->meta);
    HPS_LAUNCH_OK();
    HPS_CUDA_OK(cudaMemsetAsync(g->cell_cap, 0, T * sizeof(int), stream));
    if (n > 0) {
        int64_t nb64 = ceil_div(n, tb);
        int nb = (int)(nb64 < kNumSMs * 8 ? nb64 : kNumSMs * 8);
        grid_bbox_kernel<<<nb, tb, 0, stream>>>(pos, n, g->meta);
        HPS_LAUNCH_OK();
        grid_origin_kernel<<<1, 32, 0, stream>>>(g->meta, inv);
        HPS_LAUNCH_OK();
        grid_insert_kernel<<<(unsigned)ceil_div(n, tb), tb, 0, stream>>>(pos, n, inv, g->table, (unsigned)(T - 1),
                                                                         g->meta, g->pt_slot, g->pt_rank);
        HPS_LAUNCH_OK();
        grid_alloc_kernel<<<(unsigned)ceil_div(T, tb), tb, 0, stream>>>(g->table, T, g->meta, g->cell_cap);
        HPS_LAUNCH_OK();
        grid_scatter_kernel<<<(unsigned)ceil_div(n, tb), tb, 0, stream>>>(pos, n, g->table, g->pt_slot, g->pt_rank,
                                                                          g->sorted);
        HPS_LAUNCH_OK();
    }
    return 0;
}

extern "C" int hps_grid_append(hps_grid_t* g, const float* pos, int64_t n_old, int64_t n_new, int* status, void* stream_) {
    using namespace hps;
    HPS_REQUIRE(g != nullptr && g->T > 0, "grid not built");
    HPS_REQUIRE(status != nullptr, "null status pointer");
    HPS_REQUIRE(n_old == g->n, "n_old must be the point count the grid holds");
    HPS_REQUIRE(n_new >= n_old && n_new < (1ll << 30), "point count out of range");
    *status = 0;
    const int64_t n_add = n_new - n_old;
    if (n_add == 0) return 0;
    HPS_REQUIRE(pos != nullptr, "null positions");
    if (n_new > g->cap_n || 2 * n_new > g->T) {      // the build-time capacities are spent: the caller rebuilds
        *status = 8;
        return 0;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    const float inv = 1.0f / g->cell;
    const int tb = 256;
    const unsigned nb = (unsigned)ceil_div(n_add, tb);
    grid_append_insert_kernel<<<nb, tb, 0, stream>>>(pos, n_old, n_add, inv, g->table, (unsigned)(g->T - 1), g->meta,
                                                     g->pt_slot, g->pt_rank);
    HPS_LAUNCH_OK();
    grid_append_alloc_kernel<<<nb, tb, 0, stream>>>(n_add, g->table, g->cell_cap, g->meta, g->pt_slot, g->pt_rank,
                                                    g->sorted, g->sorted_cap);
    HPS_LAUNCH_OK();
    grid_append_scatter_kernel<<<nb, tb, 0, stream>>>(pos, n_old, n_add, g->table, g->cell_cap, g->pt_slot, g->pt_rank,
                                                      g->sorted);
    HPS_LAUNCH_OK();
    GridMeta h;
    HPS_CUDA_OK(cudaMemcpyAsync(&h, g->meta, sizeof(h), cudaMemcpyDeviceToHost, stream));
    HPS_CUDA_OK(cudaStreamSynchronize(stream));
    *status = h.status;
    if (h.status == 0) g->n = n_new;
    return 0;
}

extern "C" int hps_grid_info(const hps_grid_t* g, int64_t* n_points, int64_t* table_size, int64_t* n_cells,
                             float* cell_size, int* status, void* stream_) {
    HPS_REQUIRE(g != nullptr, "null grid");
    hps::GridMeta h;
    HPS_CUDA_OK(cudaMemcpyAsync(&h, g->meta, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
    HPS_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream_));
    if (n_points) *n_points = g->n;
    if (table_size) *table_size = g->T;
    if (n_cells) *n_cells = h.n_cells;
    if (cell_size) *cell_size = g->cell;
    if (status) *status = h.status;
    return 0;
}

extern "C" int hps_knn_query(const hps_grid_t* g, const float* q, const float* t_up, const float* t_dn, int64_t nq,
                             int32_t* I, float* D, int32_t* count, int32_t* cand, int group, void* stream_) {
    using namespace hps;
    HPS_REQUIRE(g != nullptr && g->T > 0, "grid not built");
    HPS_REQUIRE(nq >= 0, "negative query count");
    HPS_REQUIRE(group == 1 || group == 5, "group must be 1 or 5");
    if (nq == 0) return 0;
    HPS_REQUIRE(q && t_up && t_dn && I && D && count, "null buffer");
    GridView v;
    v.sorted = g->sorted;
    v.table = g->table;
    v.meta = g->meta;
    v.mask = (unsigned)(g->T - 1);
    v.inv_cell = 1.0f / g->cell;
    const RaySource none{};
    HPS_SHARE_SM((knn_query_kernel<5, false>));
    HPS_SHARE_SM((knn_query_kernel<1, false>));
    if (group == 5) {     // the samples of one ray share a sweep
        knn_query_kernel<5, false><<<(unsigned)ceil_div(ceil_div(nq, 5), kQueryWarps), kQueryWarps * 32, 0, (cudaStream_t)stream_>>>(
            v, q, t_up, t_dn, nq, I, D, count, cand, none, nullptr, nullptr, nullptr, nullptr);
    } else {
        knn_query_kernel<1, false><<<(unsigned)ceil_div(nq, kQueryWarps), kQueryWarps * 32, 0, (cudaStream_t)stream_>>>(
            v, q, t_up, t_dn, nq, I, D, count, cand, none, nullptr, nullptr, nullptr, nullptr);
    }
    HPS_LAUNCH_OK();
    return 0;
}

extern "C" int hps_sample_knn(const hps_grid_t* g, const float* rays_o, const float* rays_d, const float* gt_depth,
                              const double* radius, double radius_scalar, const float* z_zero, int64_t z_zero_stride,
                              int64_t R, int S, float near, float far, float* z, float* pts, float* t_up, float* t_dn,
                              int32_t* I, float* D, int32_t* count, int32_t* cand, void* stream_) {
    using namespace hps;
    HPS_REQUIRE(g != nullptr && g->T > 0, "grid not built");
    HPS_REQUIRE(R >= 0 && S >= 1 && S <= 32, "need R >= 0, 1 <= S <= 32");
    HPS_REQUIRE(z_zero_stride == 0 || z_zero_stride >= S, "z_zero_stride must be 0 (shared row) or >= S");
    if (R == 0) return 0;
    HPS_REQUIRE(rays_o && rays_d && gt_depth && z && pts && t_up && t_dn && I && D && count, "null buffer");
    GridView v;
    v.sorted = g->sorted;
    v.table = g->table;
    v.meta = g->meta;
    v.mask = (unsigned)(g->T - 1);
    v.inv_cell = 1.0f / g->cell;
    const RaySource src{rays_o, rays_d, gt_depth, radius, radius_scalar, z_zero, z_zero_stride, S, near, far};
    const int64_t nq = R * S;
    HPS_SHARE_SM((knn_query_kernel<1, true>));
    knn_query_kernel<1, true><<<(unsigned)ceil_div(nq, kQueryWarps), kQueryWarps * 32, 0, (cudaStream_t)stream_>>>(
        v, nullptr, nullptr, nullptr, nq, I, D, count, cand, src, z, pts, t_up, t_dn);
    HPS_LAUNCH_OK();
    return 0;
}
