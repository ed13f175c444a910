// Pieces shared by the TransE (FP32 pipe) and bilinear (tcgen05) rank kernels: candidate-group descriptors, the
// work-item decomposition, the known-true list walk, and the small pre-pass kernels.
#pragma once
#include "common.h"
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
    int64_t filt_nnz;      // MRE_FILTER_CSR: upper bound on filt_ptr[Q] when the caller gave one, else 0
    // optional cost-balanced work-item order (TransE kernel, candidate groups): sched[k * gridDim.x + cta] = k-th item of the
    // CTA, -1 past its last; NULL => round-robin (item = cta + k * gridDim.x)
    const int32_t *sched;
    int32_t sched_rounds;
    // outputs [4][Q]
    int32_t *counts;
};

// k-th work item of this CTA, or -1 when it has none left
__device__ __forceinline__ int64_t cta_item(const RankParams &p, int64_t k) {
    if (p.sched) return k < p.sched_rounds ? (int64_t)__ldg(p.sched + k * gridDim.x + blockIdx.x) : -1;
    const int64_t item = blockIdx.x + k * (int64_t)gridDim.x;
    return item < p.total_items ? item : -1;
}

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

// ------------------------------------------------------------------------------------------ known-true correction
// The reference asks "_find(candidate)?" once per counted candidate (Test.h:80-87, Corrupt.h:166-177).  Here the question is
// inverted: the tile kernels add every candidate that beats the true entity to the raw AND the filtered counters, and this
// pass scores the few known-true entities of each query (and the true entity itself) with the scalar scorer's arithmetic --
// the arithmetic the tile kernels' decisions agree with bit for bit -- and takes them back OUT of the filtered counters.
// Integer atomics commute, so the pass may run before, beside or after the tile kernel.
//
// One warp per query; the hardware block scheduler balances the few long lists (FB15K237 head queries reach 4 364 known heads)
// against the many short ones.  The list is walked 32 entries at a time (entry i -> lane i % 32): distinct entities that lie in
// the query's candidate set survive (lists are sorted ascending -- the index's runs are, MRE_FILTER_CSR slices must be -- so
// repeated ids are adjacent and counted once; slot `hi` stands for the true entity itself).  A batch is scored in 32-wide
// chunks of d: the lanes first work ACROSS d, entry by entry (coalesced row reads; the element term -- one rounding, the scalar
// scorer's own -- is parked in shared memory), then ALONG d (lane e folds entry e's terms in order): 32 sequential sums side by
// side, bit-identical to the scalar scorer, without the 32-lines-per-load access pattern of a row per lane.
//   Op::query(q, ...)       per-query setup (row pointers);  Op::thresholds(q)  the query's thresholds (for classify)
//   Op::vec(d)              element d of the query vector (the lane's own d of the chunk)
//   Op::term(v, e)          the parked element term from the query element and the entity element
//   Op::fold(acc, term)     acc <- acc (+) term, the scalar scorer's accumulation step
//   Op::classify(acc, lt, eq)
constexpr int KNOWN_WARPS = 8;
constexpr int KNOWN_SEG = 32;        // entries of a run one warp claims at a time (shared-run pass)
constexpr int KNOWN_DIRECT = 8;     // segments this short are scored one entry per lane by the scalar scorer
template <class Op>
__device__ __forceinline__ void known_correction(const RankParams &p, Op op, float (*sT)[33], int64_t *sX) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * KNOWN_WARPS + (threadIdx.x >> 5);
    if (q >= p.Q) return;
    const int side = query_side(p, q);
    const int64_t h = p.q_h[q], t = p.q_t[q], r = p.q_r[q];
    const int64_t truth = side ? t : h;
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
    const GroupDesc &gd = p.groups[p.all_entities ? 0 : group_of_query(p, q)];
    if (q < gd.q0 || q - gd.q0 >= gd.nq) return;            // the query's own group was empty (dropped): nothing is scored for it
    op.query(q, side, h, t, r);
    op.thresholds(q);
    const int D = (int)p.D;
    int k_lt = 0, k_eq = 0;
    for (int64_t base = lo; base <= hi; base += 32) {       // warp-uniform trip count
        const int64_t i = base + lane;
        int64_t x = -1;
        if (i <= hi) {
            x = i < hi ? __ldg(list + i) : truth;
            if (i < hi && (x == truth || (i > lo && __ldg(list + i - 1) == x))) x = -1;   // the truth goes last; duplicates once
            if (x < 0 || x >= p.E) x = -1;
            if (x >= 0 && !p.all_entities) {
                const int64_t k = lower_bound_i64(p.cand_idx, gd.c0, gd.c0 + gd.nc, x);
                if (k >= gd.c0 + gd.nc || __ldg(p.cand_idx + k) != x) x = -1;
            }
        }
        const unsigned live = __ballot_sync(0xffffffffu, x >= 0);
        if (!live) continue;
        // the batch's surviving entities, compacted: entry k < n is scored by lane k
        const int n = __popc(live);
        if (x >= 0) sX[__popc(live & ((1u << lane) - 1u))] = x;
        __syncwarp();
        float acc = 0.f;
        for (int c0 = 0; c0 < D; c0 += 32) {
            const int d = c0 + lane;
            const float v = d < D ? op.vec(d) : 0.f;
            float val[32];
#pragma unroll
            for (int k = 0; k < 32; k++)                     // every row element of the chunk is requested before any is used
                if (k < n) val[k] = d < D ? __ldg(p.ent + sX[k] * p.D + d) : 0.f;
#pragma unroll
            for (int k = 0; k < 32; k++)
                if (k < n) sT[k][lane] = op.term(v, val[k]);
            __syncwarp();
            if (lane < n) {
                const int nd = min(32, D - c0);
#pragma unroll 8
                for (int dd = 0; dd < nd; dd++) acc = op.fold(acc, sT[lane][dd]);
            }
            __syncwarp();
        }
        if (lane < n) op.classify(acc, k_lt, k_eq);
        __syncwarp();                                        // sX is rewritten by the next batch
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        k_lt += __shfl_xor_sync(0xffffffffu, k_lt, m);
        k_eq += __shfl_xor_sync(0xffffffffu, k_eq, m);
    }
    if (lane == 0) {
        if (k_lt) atomicSub(p.counts + 2 * p.Q + q, k_lt);
        if (k_eq) atomicSub(p.counts + 3 * p.Q + q, k_eq);
    }
}

// The same pass FLATTENED over (query, list entry) pairs, one pair per lane, for jobs whose list prefix is known on the host
// (MRE_FILTER_CSR with filt_nnz, MRE_FILTER_NONE: only the true entities): the zero-shot test sets hold 2-3 known tails per
// query, so a warp per query would leave 28 lanes idle through every dependent memory round trip.  Pair g belongs to the query q
// with filt_ptr[q] + q <= g < filt_ptr[q + 1] + q + 1 (one binary search over the prefix; slot filt_ptr[q + 1] - filt_ptr[q] of a
// query is its true entity).  A warp scores its 32 pairs side by side: for every 32-wide chunk of d and every live pair, the
// lanes fetch that pair's query-vector elements and entity-row elements across d (two coalesced reads), park the term in
// shared memory, then lane e folds pair e's terms in order.
//   Op::flat_query(q, slot, side, h, t, r)   per-lane setup (thresholds); returns the base pointer of the query vector
//   Op::VSTRIDE                              element d of the query vector at base[d * VSTRIDE]
template <class Op>
__device__ __forceinline__ void known_correction_flat(const RankParams &p, Op op, float (*sT)[33], int64_t *sX, const float **sV) {
    const int lane = threadIdx.x & 31;
    const int64_t total = p.filt_nnz + p.Q;
    const int64_t g0 = ((int64_t)blockIdx.x * KNOWN_WARPS + (threadIdx.x >> 5)) * 32;
    if (g0 >= total) return;
    const int64_t g = g0 + lane;
    int64_t x = -1, q = 0;
    if (g < total) {
        int64_t lo = 0, hi = p.Q;                          // c(q) = filt_ptr[q] + q;  c(lo) <= g < c(hi)
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if ((p.filt_ptr ? __ldg(p.filt_ptr + mid) : 0) + mid <= g) lo = mid; else hi = mid;
        }
        q = lo;
        const int64_t b = p.filt_ptr ? __ldg(p.filt_ptr + q) : 0, e = p.filt_ptr ? __ldg(p.filt_ptr + q + 1) : 0;
        const int64_t i = b + (g - (b + q));
        const int side = query_side(p, q);
        const int64_t h = p.q_h[q], t = p.q_t[q], r = p.q_r[q];
        const int64_t truth = side ? t : h;
        if (i <= e) {                                      // i > e: filt_nnz was an over-estimate
            x = i < e ? __ldg(p.filt_idx + i) : truth;
            if (i < e && (x == truth || (i > b && __ldg(p.filt_idx + i - 1) == x))) x = -1;   // the truth goes last; duplicates once
            if (x < 0 || x >= p.E) x = -1;
        }
        const GroupDesc &gd = p.groups[p.all_entities ? 0 : group_of_query(p, q)];
        if (q < gd.q0 || q - gd.q0 >= gd.nq) x = -1;       // the query's own group was empty (dropped)
        if (x >= 0 && !p.all_entities) {
            const int64_t k = lower_bound_i64(p.cand_idx, gd.c0, gd.c0 + gd.nc, x);
            if (k >= gd.c0 + gd.nc || __ldg(p.cand_idx + k) != x) x = -1;
        }
        if (x >= 0) {
            sX[lane] = x;
            sV[lane] = op.flat_query(q, gd.s0 + (q - gd.q0), side, h, t, r);
        }
    }
    const unsigned live = __ballot_sync(0xffffffffu, x >= 0);
    if (!live) return;
    __syncwarp();
    const int D = (int)p.D;
    float acc = 0.f;
    for (int c0 = 0; c0 < D; c0 += 32) {
        const int d = c0 + lane;
        float val[32], vq[32];
#pragma unroll
        for (int k = 0; k < 32; k++)                         // every element of the chunk is requested before any is used
            if ((live >> k) & 1u) {
                val[k] = d < D ? __ldg(p.ent + sX[k] * p.D + d) : 0.f;
                vq[k] = d < D ? __ldg(sV[k] + (int64_t)d * Op::VSTRIDE) : 0.f;
            }
#pragma unroll
        for (int k = 0; k < 32; k++)
            if ((live >> k) & 1u) sT[k][lane] = op.term(vq[k], val[k]);
        __syncwarp();
        if (x >= 0) {
            const int nd = min(32, D - c0);
#pragma unroll 8
            for (int dd = 0; dd < nd; dd++) acc = op.fold(acc, sT[lane][dd]);
        }
        __syncwarp();
    }
    if (x >= 0) {
        int k_lt = 0, k_eq = 0;
        op.classify(acc, k_lt, k_eq);
        if (k_lt) atomicSub(p.counts + 2 * p.Q + q, 1);
        if (k_eq) atomicSub(p.counts + 3 * p.Q + q, 1);
    }
}

// MRE_FILTER_INDEX: the list of a query is the index's run of its (fixed entity, relation) -- a function of exactly the things
// the query VECTOR is a function of.  All queries that share a run share the scores of the run's entities, so each run is
// scored ONCE per job (known_score_runs: the first query to stamp the run's first slot with the job's epoch scores it, 32
// entries side by side as above, into a score column parallel to the index's payload column), and every query then only
// COMPARES its own thresholds against its run's stored scores (known_compare_runs: 12 bytes per list entry instead of a row).
// FB15K237's test set: 9.7 M (query, known entity) pairs, but only the ~0.3 M distinct entries of the touched runs are scored.
// lower_bound by a whole warp: 32 probes per round trip instead of one (a binary search over the 3 x 10^5 keys of FB15K237 is 18
// DEPENDENT loads; this is 4).  All lanes must call it with the same arguments; every lane gets the result.
__device__ __forceinline__ int64_t warp_lower_bound_i64(const int64_t *__restrict__ keys, int64_t lo, int64_t hi, int64_t key, int lane) {
    // invariant: keys[i] < key for i < lo, keys[i] >= key for i >= hi
    while (hi - lo > 32) {
        const int64_t stride = (hi - lo + 31) >> 5;                   // probes lo + (l + 1) * stride - 1, l = 0..31 (clipped to hi - 1)
        const int64_t pos = min(hi - 1, lo + (int64_t)(lane + 1) * stride - 1);
        const unsigned below = __ballot_sync(0xffffffffu, __ldg(keys + pos) < key);   // monotone: a prefix of the lanes
        const int c = __popc(below);
        const int64_t nlo = c == 0 ? lo : min(hi - 1, lo + (int64_t)c * stride - 1) + 1;
        const int64_t nhi = c == 32 ? hi : min(hi - 1, lo + (int64_t)(c + 1) * stride - 1);
        lo = nlo;
        hi = nhi;
    }
    const int64_t pos = lo + lane;
    const unsigned below = __ballot_sync(0xffffffffu, pos < hi && __ldg(keys + pos) < key);
    return lo + __popc(below);
}

struct KnownRuns {
    float *score0, *score1;            // [n_all] per orientation (0: (t, r) -> heads, 1: (h, r) -> tails): score of the payload entity
    unsigned int *stamp0, *stamp1;     // [n_all]: epoch of the job that scored the run starting at this slot
    int64_t *range;            // [Q][2]: the query's run [lo, hi) in its orientation's columns
    unsigned int epoch;
};

template <class Op>
__device__ __forceinline__ void known_score_runs(const RankParams &p, const KnownRuns kr, Op op, float (*sT)[33], int64_t *sX) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * KNOWN_WARPS + (threadIdx.x >> 5);
    if (q >= p.Q) return;
    const int side = query_side(p, q);
    const int64_t h = p.q_h[q], t = p.q_t[q], r = p.q_r[q];
    const int64_t *keys = side ? p.hr_key : p.tr_key, *list = side ? p.hr_val : p.tr_val;
    const int64_t key = (side ? h : t) * p.R + r;
    // the run's start by a warp-wide search; its end is nearly always within the next 32 slots (one more round trip)
    const int64_t lo = warp_lower_bound_i64(keys, 0, p.n_all, key, lane);
    int64_t hi = warp_lower_bound_i64(keys, lo, min(p.n_all, lo + 32), key + 1, lane);
    if (hi == lo + 32) hi = warp_lower_bound_i64(keys, hi, p.n_all, key + 1, lane);
    if (lane == 0) { kr.range[2 * q] = lo; kr.range[2 * q + 1] = hi; }
    if (hi == lo) return;
    const int D = (int)p.D;
    float *out = side ? kr.score1 : kr.score0;
    unsigned int *stamp = side ? kr.stamp1 : kr.stamp0;
    bool ready = false;                                       // op.query() done (once, by a warp that scores something)
    // A run is scored in segments of KNOWN_SEG entries, each claimed by stamping its first slot with the job's epoch: the queries
    // that share a long run (FB15K237: up to 4 364 known heads, shared by hundreds of test triples) split it between their warps.
    for (int64_t s0 = lo; s0 < hi; s0 += KNOWN_SEG) {
        unsigned int old = 0;
        if (lane == 0) old = atomicExch(stamp + s0, kr.epoch);
        if (__shfl_sync(0xffffffffu, old, 0) == kr.epoch) continue;      // another warp of this job scores the segment
        if (!ready) { op.query(q, side, h, t, r); ready = true; }
        const int64_t s1 = min(hi, s0 + KNOWN_SEG);
        if (Op::DIRECT_ONLY || s1 - s0 <= KNOWN_DIRECT) {
            // a handful of entries (the zero-shot test sets: 2-3 known tails per (h, r)): one entry per lane, scored by the scalar
            // scorer itself -- few lanes, so the row-per-lane access pattern costs nothing, and no instruction is spent on idle lanes
            const int64_t i = s0 + lane;
            if (i < s1) {
                const int64_t x = __ldg(list + i);
                out[i] = (x >= 0 && x < p.E) ? op.direct(x) : __int_as_float(0x7fc00000);
            }
            continue;
        }
        for (int64_t base = s0; base < s1; base += 32) {
            const int64_t i = base + lane;
            int64_t x = i < s1 ? __ldg(list + i) : -1;
            if (x >= p.E) x = -1;
            const unsigned live = __ballot_sync(0xffffffffu, x >= 0);
            if (x >= 0) sX[lane] = x;
            __syncwarp();
            float acc = 0.f;
            for (int c0 = 0; c0 < D; c0 += 32) {
                const int d = c0 + lane;
                const float v = d < D ? op.vec(d) : 0.f;
                float val[32];
#pragma unroll
                for (int k = 0; k < 32; k++)
                    if ((live >> k) & 1u) val[k] = d < D ? __ldg(p.ent + sX[k] * p.D + d) : 0.f;
#pragma unroll
                for (int k = 0; k < 32; k++)
                    if ((live >> k) & 1u) sT[k][lane] = op.term(v, val[k]);
                __syncwarp();
                if (x >= 0) {
                    const int nd = min(32, D - c0);
#pragma unroll 8
                    for (int dd = 0; dd < nd; dd++) acc = op.fold(acc, sT[lane][dd]);
                }
                __syncwarp();
            }
            if (i < s1) out[i] = x >= 0 ? acc : __int_as_float(0x7fc00000);      // NaN never counts
        }
    }
}

template <class Op>
__device__ __forceinline__ void known_compare_runs(const RankParams &p, const KnownRuns kr, Op op) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * KNOWN_WARPS + (threadIdx.x >> 5);
    if (q >= p.Q) return;
    const int side = query_side(p, q);
    const int64_t truth = side ? p.q_t[q] : p.q_h[q];
    const GroupDesc &gd = p.groups[p.all_entities ? 0 : group_of_query(p, q)];
    if (q < gd.q0 || q - gd.q0 >= gd.nq) return;            // the query's own group was empty (dropped): nothing is scored for it
    op.thresholds(q);
    const int64_t lo = kr.range[2 * q], hi = kr.range[2 * q + 1];
    const int64_t *list = side ? p.hr_val : p.tr_val;
    const float *sc = side ? kr.score1 : kr.score0;
    int k_lt = 0, k_eq = 0;
    for (int64_t i = lo + lane; i <= hi; i += 32) {         // slot hi stands for the true entity itself
        const int64_t x = i < hi ? __ldg(list + i) : truth;
        if (i < hi && (x == truth || (i > lo && __ldg(list + i - 1) == x))) continue;   // the truth goes last; duplicates once
        if (x < 0 || x >= p.E) continue;
        if (!p.all_entities) {
            const int64_t k = lower_bound_i64(p.cand_idx, gd.c0, gd.c0 + gd.nc, x);
            if (k >= gd.c0 + gd.nc || __ldg(p.cand_idx + k) != x) continue;
        }
        if (i < hi) op.classify(sc[i], k_lt, k_eq);
        else if (op.truth_ties()) k_eq++;                   // the true entity ties with itself (unless its score is NaN / inf)
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        k_lt += __shfl_xor_sync(0xffffffffu, k_lt, m);
        k_eq += __shfl_xor_sync(0xffffffffu, k_eq, m);
    }
    if (lane == 0) {
        if (k_lt) atomicSub(p.counts + 2 * p.Q + q, k_lt);
        if (k_eq) atomicSub(p.counts + 3 * p.Q + q, k_eq);
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
