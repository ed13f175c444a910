// Pieces shared by the TransE (FP32 pipe) and bilinear (tcgen05) rank kernels: candidate-group descriptors,
// the work-item decomposition and the known-true correction pass.
//
// Rank counts are computed as   filtered = dense - correction   where
//   dense      = #{j in S_q : s_j < s_true}          over the WHOLE candidate set (all E entities, or the group's list),
//                produced tile by tile without ever writing a score, and
//   correction = #{j in (known_q u {true_q}) n S_q : s_j < s_true}   over the query's short known-true list.
// That is Test.h:80-87's "l_s++ ; if (!_find(...)) l_filter_s++" with the membership test hoisted out of the
// E-long loop: instead of E binary searches in tripleList per query (Corrupt.h:166-177) the filter costs one
// lower_bound to locate the query's run in the packed-key CSR plus |known| extra scores.
#pragma once
#include "device_utils.cuh"

namespace mre {

constexpr int TILE_Q = 128;  // queries per work item
constexpr int TILE_E = 128;  // candidate entities per work item (TransE); the bilinear kernel uses its own N

struct GroupDesc {
    int64_t q0, nq;      // query range of the group
    int64_t c0, nc;      // candidate range in cand_idx (c0 = 0, nc = E for the all-entity group)
    int64_t item0;       // first work item of the group
    int32_t n_qt, n_et;  // tiles along queries / candidates
    int32_t pad0, pad1;
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

// Known-true correction for query q, executed by one full warp.  Score(q, entity) must return the value the
// dense phase compares against thr (the raw accumulator).  Returns through atomics on counts[2], counts[3].
template <bool NEED_EQ, bool INCLUDE_TRUTH = true, class ScoreFn>
__device__ __forceinline__ void correct_query(const RankParams &p, int64_t q, int lane, ScoreFn score) {
    const int side = query_side(p, q);
    const int64_t h = p.q_h[q], t = p.q_t[q], r = p.q_r[q];
    const int64_t truth = side ? t : h;
    const float2 th = p.thr[q];
    const int64_t *list = nullptr;
    int64_t lo = 0, hi = 0;
    if (p.filter == MRE_FILTER_INDEX) {
        const int64_t *keys = side ? p.hr_key : p.tr_key;
        const int64_t key = (side ? h : t) * p.R + r;
        lo = lower_bound_i64(keys, 0, p.n_all, key);
        hi = lower_bound_i64(keys, lo, p.n_all, key + 1);
        list = side ? p.hr_val : p.tr_val;
    } else if (p.filter == MRE_FILTER_CSR) {
        lo = p.filt_ptr[q];
        hi = p.filt_ptr[q + 1];
        list = p.filt_idx;
    }
    int64_t c0 = 0, c1 = 0;
    if (!p.all_entities) {
        const GroupDesc &gd = p.groups[group_of_query(p, q)];
        c0 = gd.c0;
        c1 = gd.c0 + gd.nc;
    }
    int n_lt = 0, n_eq = 0;
    // entries lo..hi-1 are the known entities; index hi stands for the true entity itself (when the dense phase
    // counted it: the TransE kernel does, the tensor-core kernel excludes it by index)
    const int64_t last = INCLUDE_TRUTH ? hi : hi - 1;
    for (int64_t base = lo; base <= last; base += 32) {
        int64_t i = base + lane;
        bool lt = false, eq = false;
        if (i <= last) {
            int64_t x = i < hi ? __ldg(list + i) : truth;
            bool use = (i == hi) || (x != truth);
            if (use && (x < 0 || x >= p.E)) use = false;
            if (use && !p.all_entities) use = contains_i64(p.cand_idx, c0, c1, x);
            if (use) {
                float s = score(q, x);
                lt = s < th.x;
                if (NEED_EQ) eq = !lt && s < th.y;
            }
        }
        n_lt += __popc(__ballot_sync(0xffffffffu, lt));
        if (NEED_EQ) n_eq += __popc(__ballot_sync(0xffffffffu, eq));
    }
    if (lane == 0) {
        if (n_lt) atomicSub(p.counts + 2 * p.Q + q, n_lt);
        if (NEED_EQ && n_eq) atomicSub(p.counts + 3 * p.Q + q, n_eq);
    }
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
