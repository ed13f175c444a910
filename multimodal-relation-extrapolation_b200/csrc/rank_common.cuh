// Pieces shared by the TransE (FP32 pipe) and bilinear (tcgen05) rank kernels: candidate-group descriptors, the
// work-item decomposition, and the small pre-pass kernels.  The known-true filter lives in tile_filter.cu.
#pragma once
#include "device_utils.cuh"

namespace mre {

constexpr int TILE_Q = 128;  // queries per work item
constexpr int TILE_E = 128;  // candidate entities per work item (TransE); the bilinear kernel uses its own N

struct GroupDesc {
    int64_t q0, nq;      // query range of the group
    int64_t c0, nc;      // candidate range in cand_idx (c0 = 0, nc = E for the all-entity group)
    int64_t item0;       // first work item of the group
    int64_t s0;          // first query SLOT of the group (even): the TransE kernel's pair-interleaved query-vector layout
    int64_t pitem0;      // first CTA-PAIR work item of the group (two query tiles x one candidate tile): bilinear pair kernel
    int32_t n_qt, n_et;  // tiles along queries / candidates
};

struct RankParams {
    // tables after the pre-pass (normalised / padded as needed)
    const float *ent;  // [E, D]
    int64_t E, R, D;
    // queries
    const int64_t *q_h, *q_t, *q_r;
    const uint8_t *q_side;
    int32_t side;
    int64_t Q;
    const float *qvec;  // [Q, D] per-query vector
    const float2 *thr;  // [Q] (lo, hi): lt <=> acc < lo ; eq <=> !(lt) && acc < hi
    // groups
    const GroupDesc *groups;
    int32_t n_groups;
    int32_t all_entities;  // 1 => single group over [0, E), cand_idx unused
    const int64_t *cand_idx;
    int64_t total_items;
    int64_t total_slots;   // query slots over all groups (each group rounded up to an even count)
    int64_t total_pitems;  // CTA-pair work items over all groups
    // filter
    int32_t filter;
    const int64_t *hr_key, *hr_val, *tr_key, *tr_val;
    int64_t n_all;
    const int64_t *filt_ptr, *filt_idx;
    // known-true pairs per work item (tile_filter.cu): pairs[ptr[item] .. ptr[item+1]) = (row << 16 | col)
    const uint32_t *tf_ptr, *tf_pairs;
    // outputs [4][Q]
    int32_t *counts;
};

__device__ __forceinline__ int query_side(const RankParams &p, int64_t q) { return p.q_side ? (int)p.q_side[q] : p.side; }

// item -> (group, q-tile, e-tile); q-tiles vary fastest so CTAs running side by side share candidate tiles in L2
__device__ __forceinline__ void decode_item(const RankParams &p, int64_t item, int &g, int &qt, int &et) {
    int lo = 0, hi = p.n_groups;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (p.groups[mid].item0 <= item) lo = mid; else hi = mid;
    }
    g = lo;
    int64_t local = item - p.groups[g].item0;
    qt = (int)(local % p.groups[g].n_qt);
    et = (int)(local / p.groups[g].n_qt);
}

__device__ __forceinline__ int group_of_query(const RankParams &p, int64_t q) {
    int lo = 0, hi = p.n_groups;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (p.groups[mid].q0 <= q) lo = mid; else hi = mid;
    }
    return lo;
}

// candidate groups: copy the listed entity rows into one dense table so that the main kernel streams every
// candidate tile with the same TMA boxes as the all-entity case
static __global__ void gather_rows_kernel(const float *__restrict__ ent, int64_t D, const int64_t *__restrict__ idx, int64_t n,
                                   float *__restrict__ out) {
    const int64_t total = n * (D >> 2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / (D >> 2), c = i - row * (D >> 2);
        reinterpret_cast<float4 *>(out + row * D)[c] = reinterpret_cast<const float4 *>(ent + __ldg(idx + row) * D)[c];
    }
}

static __global__ void init_counts_kernel(int32_t *counts, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) counts[i] = 0;
}

static inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block < 148 * 32 ? (n + block - 1) / block : 148 * 32); }

}  // namespace mre
