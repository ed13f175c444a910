// Counter-based Philox Bernoulli negative sampler.
//
// Reference path replaced (paths relative to /root/reference):
//   OpenKE/openke/base/Base.cpp:78-159   getBatch: positive draw, tph/hpt Bernoulli choice, batch layout
//   OpenKE/openke/base/Base.cpp:161-197  sampling: pthread fan-out over row slices
//   OpenKE/openke/base/Corrupt.h:7-83    corrupt_head / corrupt_tail: exact-uniform draw outside the true set
//   OpenKE/openke/base/Random.h:11-29    per-thread LCG  (replaced by Philox4x32-10, see include/mre_b200.h)
// The reference walks rows sequentially per pthread because its LCG is a serial stream.  A counter-based
// generator makes every output slot independent: one GPU thread per slot (positive or negative), addressed as
// (row b, slot k), so writes are coalesced along b exactly in the reference's [B pos | neg blocks of B] layout.
// The filtered draw keeps the reference's order-statistic skip (no rejection loop => no divergence tail, and
// never a train triple), with the (entity, relation) run located by two lower_bounds on a packed key column.
#include <algorithm>

#include "common.h"
#include "device_utils.cuh"

namespace mre {

struct SamplerTables {
    const int64_t *tr_h, *tr_r, *tr_t;  // trainList columns
    const int64_t *hr_key;              // h*R + r, sorted; payload tr_t
    const int64_t *tr_key, *tr_val;     // t*R + r, sorted; payload heads
    const float *bern_prob;
    int64_t n_train, E, R;
};

// tmp-th entity NOT in the sorted run vals[ll..rr] (Corrupt.h:25-43 / 64-82)
__device__ __forceinline__ int64_t skip_draw(const int64_t *__restrict__ vals, int64_t ll, int64_t rr, int64_t E, uint64_t word,
                                             int64_t fallback) {
    const int64_t k = rr - ll + 1;
    if (E - k <= 0) return fallback;  // every entity completes a train triple: the reference would divide by zero
    const int64_t tmp = (int64_t)(word % (uint64_t)(E - k));
    if (tmp < __ldg(vals + ll)) return tmp;
    if (tmp > __ldg(vals + rr) - k) return tmp + k;
    int64_t lo = ll, hi = rr + 1;
    while (lo + 1 < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(vals + mid) - (mid - ll) - 1 < tmp) lo = mid; else hi = mid;
    }
    return tmp + (lo - ll) + 1;
}

__global__ void __launch_bounds__(256) sample_kernel(const SamplerTables T, uint32_t k0, uint32_t k1, uint32_t c2, uint32_t c3,
                                                     int64_t B, int64_t neg, int mode, int bern, int64_t *__restrict__ bh,
                                                     int64_t *__restrict__ bt, int64_t *__restrict__ br, float *__restrict__ by) {
    const int64_t n = B * (1 + neg);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = idx / B, b = idx - k * B;
        Philox4 x = philox4x32_10((uint32_t)b, 0u, c2, c3, k0, k1);
        const int64_t i = (int64_t)((((uint64_t)x.x[1] << 32) | x.x[0]) % (uint64_t)T.n_train);
        int64_t h = __ldg(T.tr_h + i), r = __ldg(T.tr_r + i), t = __ldg(T.tr_t + i);
        float y = 1.f;
        if (k > 0) {
            x = philox4x32_10((uint32_t)b, (uint32_t)k, c2, c3, k0, k1);
            const uint64_t word = ((uint64_t)x.x[2] << 32) | x.x[1];
            const float prob = bern ? __ldg(T.bern_prob + r) : 500.0f;
            const bool keep_head = mode == 0 ? ((float)(x.x[0] % 1000u) < prob) : (mode != -1);
            if (keep_head) {
                const int64_t key = h * T.R + r;
                const int64_t ll = lower_bound_i64(T.hr_key, 0, T.n_train, key);
                const int64_t rr = lower_bound_i64(T.hr_key, ll, T.n_train, key + 1) - 1;
                t = skip_draw(T.tr_t, ll, rr, T.E, word, t);
            } else {
                const int64_t key = t * T.R + r;
                const int64_t ll = lower_bound_i64(T.tr_key, 0, T.n_train, key);
                const int64_t rr = lower_bound_i64(T.tr_key, ll, T.n_train, key + 1) - 1;
                h = skip_draw(T.tr_val, ll, rr, T.E, word, h);
            }
            y = -1.f;
        }
        bh[idx] = h; bt[idx] = t; br[idx] = r; by[idx] = y;
    }
}

int sample(mre_ctx *ctx, const mre_index *ix, uint64_t seed, uint64_t step, uint32_t stream_id, int64_t B, int64_t neg,
           int32_t mode, int32_t bern, int64_t *h, int64_t *t, int64_t *r, float *y, cudaStream_t st) {
    MRE_CHECK_ARG(ix != nullptr, "index is NULL");
    MRE_CHECK_ARG(ix->device == ctx->device, "index is not on device %d (call mre_index_to_device)", ctx->device);
    MRE_CHECK_ARG(B > 0 && neg >= 0, "B must be positive and neg non-negative");
    MRE_CHECK_ARG(B <= 0xffffffffLL && neg < 0xffffffffLL, "B and neg must fit the 32-bit Philox counter words");
    MRE_CHECK_ARG(stream_id < 65536u, "stream_id must be < 65536");
    MRE_CHECK_ARG(mode >= -1 && mode <= 1, "mode must be -1, 0 or 1");
    MRE_CHECK_ARG(ix->n_train > 0, "the index holds no train triples");
    MRE_CHECK_ARG(h && t && r && y, "NULL output");
    SamplerTables T{ix->d_tr_h, ix->d_tr_r, ix->d_tr_t, ix->d_tr_hr_key, ix->d_tr_tr_key, ix->d_tr_tr_val, ix->d_bern_prob,
                    ix->n_train, ix->E, ix->R};
    const uint32_t c2 = (uint32_t)step, c3 = ((uint32_t)(step >> 32) & 0xffffu) | (stream_id << 16);
    const int64_t n = B * (1 + neg);
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8);
    sample_kernel<<<grid, 256, 0, st>>>(T, (uint32_t)seed, (uint32_t)(seed >> 32), c2, c3, B, neg, mode, bern, h, t, r, y);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

// ------------------------------------------------------------------------------------ type-constrained corruption
// corrupt(h, r), Corrupt.h:179-195: a tail drawn uniformly from the relation's tail-type list (type_constrain.txt), redrawn while
// (h, r, tail) is a known triple of ANY split (_find), and after 1000 failed draws the exact-uniform corrupt_head over the train
// set.  The reference draws with libc rand(); here attempt a of pair i takes Philox counter (i_lo, a | i_hi << 16, step words),
// key = seed: bit-reproducible, one thread per (h, r) pair, no shared state.  The type lists of the index are sorted and
// de-duplicated (the reference keeps repeated ids of the file, which then weigh more); an empty list goes straight to the fallback.
constexpr int TYPED_MAX_LOOP = 1000;
__global__ void __launch_bounds__(256) corrupt_typed_kernel(const SamplerTables T, const int64_t *__restrict__ all_key,
                                                            const int64_t *__restrict__ all_val, int64_t n_all,
                                                            const int64_t *__restrict__ tail_ptr, const int64_t *__restrict__ tail_idx,
                                                            uint32_t k0, uint32_t k1, uint32_t c2, uint32_t c3,
                                                            const int64_t *__restrict__ qh, const int64_t *__restrict__ qr, int64_t n,
                                                            int64_t *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t h = qh[i], r = qr[i];
        const int64_t ll = __ldg(tail_ptr + r), cnt = __ldg(tail_ptr + r + 1) - ll;
        const int64_t key = h * T.R + r;
        const int64_t lo = lower_bound_i64(all_key, 0, n_all, key), hi = lower_bound_i64(all_key, lo, n_all, key + 1);
        const uint32_t c0 = (uint32_t)i, hi16 = (uint32_t)((uint64_t)i >> 32) << 16;
        int64_t res = -1;
        for (int loop = 0; loop < TYPED_MAX_LOOP && cnt > 0; loop++) {
            const Philox4 x = philox4x32_10(c0, (uint32_t)loop | hi16, c2, c3, k0, k1);
            const int64_t t = __ldg(tail_idx + ll + (int64_t)((((uint64_t)x.x[1] << 32) | x.x[0]) % (uint64_t)cnt));
            if (!contains_i64(all_val, lo, hi, t)) { res = t; break; }
        }
        if (res < 0) {                                    // corrupt_head(0, h, r), Corrupt.h:7-44
            const Philox4 x = philox4x32_10(c0, (uint32_t)TYPED_MAX_LOOP | hi16, c2, c3, k0, k1);
            const int64_t tl = lower_bound_i64(T.hr_key, 0, T.n_train, key);
            const int64_t tr = lower_bound_i64(T.hr_key, tl, T.n_train, key + 1) - 1;
            res = tr >= tl ? skip_draw(T.tr_t, tl, tr, T.E, ((uint64_t)x.x[1] << 32) | x.x[0], h)
                           : (int64_t)((((uint64_t)x.x[1] << 32) | x.x[0]) % (uint64_t)T.E);    // no train tail to skip: any entity
        }
        out[i] = res;
    }
}

int corrupt_typed(mre_ctx *ctx, const mre_index *ix, uint64_t seed, uint64_t step, uint32_t stream_id, const int64_t *h,
                  const int64_t *r, int64_t n, int64_t *t_out, cudaStream_t st) {
    MRE_CHECK_ARG(ix != nullptr, "index is NULL");
    MRE_CHECK_ARG(ix->device == ctx->device, "index is not on device %d (call mre_index_to_device)", ctx->device);
    MRE_CHECK_ARG(ix->has_type && ix->d_type_ptr[1], "the index holds no type constraints on the device (load them before mre_index_to_device)");
    MRE_CHECK_ARG(n >= 0 && (n == 0 || (h && r && t_out)), "bad argument");
    MRE_CHECK_ARG(stream_id < 65536u, "stream_id must be < 65536");
    if (n == 0) return MRE_OK;
    SamplerTables T{ix->d_tr_h, ix->d_tr_r, ix->d_tr_t, ix->d_tr_hr_key, ix->d_tr_tr_key, ix->d_tr_tr_val, ix->d_bern_prob,
                    ix->n_train, ix->E, ix->R};
    const uint32_t c2 = (uint32_t)step, c3 = ((uint32_t)(step >> 32) & 0xffffu) | (stream_id << 16);
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8);
    corrupt_typed_kernel<<<grid, 256, 0, st>>>(T, ix->d_all_hr_key, ix->d_all_hr_val, ix->n_all, ix->d_type_ptr[1], ix->d_type_idx[1],
                                               (uint32_t)seed, (uint32_t)(seed >> 32), c2, c3, h, r, n, t_out);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

// ------------------------------------------------------------------------------------ subgraph sampler (paper side)
// module/NegativeSampling.py:114-140 (neg_sample_fn), :321-375 (__normal_batch, __corrupt_head, __corrupt_tail): per edge
// of a sampled subgraph, neg_ent corruptions drawn from the subgraph's node list (LOCAL ids), split into head- and
// tail-corruptions by neg_ent Bernoulli(prob) draws, heads filling slots 1..nh and tails slots nh+1..neg_ent; a drawn
// node is discarded when its GLOBAL id completes a train triple with the kept (entity, relation).  The reference does this
// with random.sample + np.in1d in a Python loop per edge and retries until enough survive; here one thread owns one
// output slot, re-derives the edge's head/tail split from the same neg_ent decision words, and runs a bounded rejection
// loop on its own Philox sub-stream (attempt number in the counter), then a deterministic scan of the node list, so the
// kernel always terminates (the reference spins forever when every node is a known answer).
constexpr int SUB_MAX_ATTEMPTS = 64;

__device__ __forceinline__ bool sub_known(const SamplerTables &T, const int64_t *__restrict__ keys, const int64_t *__restrict__ vals,
                                          int64_t fixed_global, int64_t r, int64_t cand_global) {
    if (fixed_global < 0 || fixed_global >= T.E || cand_global < 0) return false;
    const int64_t key = fixed_global * T.R + r;
    const int64_t ll = lower_bound_i64(keys, 0, T.n_train, key);
    const int64_t rr = lower_bound_i64(keys, ll, T.n_train, key + 1);
    return contains_i64(vals, ll, rr, cand_global);
}

__global__ void __launch_bounds__(256) subgraph_sample_kernel(const SamplerTables T, uint32_t k0, uint32_t k1, uint32_t c2, uint32_t c3,
                                                              const int64_t *__restrict__ eh, const int64_t *__restrict__ et,
                                                              const int64_t *__restrict__ er, int64_t n_edges,
                                                              const int64_t *__restrict__ nodes, int64_t n_nodes,
                                                              const int64_t *__restrict__ l2g, int64_t n_local, int64_t neg, int bern,
                                                              int filter, int32_t *__restrict__ oh, int32_t *__restrict__ ot,
                                                              int32_t *__restrict__ orel) {
    const int64_t n = n_edges * (1 + neg);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = idx / n_edges, b = idx - k * n_edges;
        int64_t h = __ldg(eh + b), t = __ldg(et + b);
        const int64_t r = __ldg(er + b);
        if (k > 0 && n_nodes > 0) {
            // head/tail split of this edge: neg decision words, uniform [0,1) from the top 24 bits
            float prob = 0.5f;
            if (bern && r >= 0 && r < T.R) prob = __ldg(T.bern_prob + r) * 1e-3f;
            int nh = 0;
            for (int64_t j = 1; j <= neg; j++) {
                const Philox4 x = philox4x32_10((uint32_t)b, (uint32_t)j, c2, c3, k0, k1);
                nh += ((float)(x.x[0] >> 8) * 5.9604644775390625e-8f < prob) ? 1 : 0;
            }
            const bool corrupt_head = k <= nh;
            const int64_t fixed_local = corrupt_head ? t : h;
            const int64_t fixed_global = (fixed_local >= 0 && fixed_local < n_local) ? __ldg(l2g + fixed_local) : -1;
            const int64_t *keys = corrupt_head ? T.tr_key : T.hr_key;
            const int64_t *vals = corrupt_head ? T.tr_val : T.tr_t;
            int64_t pick = -1, pos = 0;
            for (int a = 0; a < SUB_MAX_ATTEMPTS && pick < 0; a++) {
                const Philox4 x = philox4x32_10((uint32_t)b, (uint32_t)k | ((uint32_t)(a + 1) << 16), c2, c3, k0, k1);
                pos = (int64_t)((((uint64_t)x.x[1] << 32) | x.x[0]) % (uint64_t)n_nodes);
                const int64_t cand = __ldg(nodes + pos);
                const int64_t cg = (cand >= 0 && cand < n_local) ? __ldg(l2g + cand) : -1;
                if (!filter || !sub_known(T, keys, vals, fixed_global, r, cg)) pick = cand;
            }
            for (int64_t step = 1; step <= n_nodes && pick < 0; step++) {   // rare: the admissible set is tiny
                const int64_t cand = __ldg(nodes + (pos + step) % n_nodes);
                const int64_t cg = (cand >= 0 && cand < n_local) ? __ldg(l2g + cand) : -1;
                if (!sub_known(T, keys, vals, fixed_global, r, cg)) pick = cand;
            }
            if (pick >= 0) {
                if (corrupt_head) h = pick; else t = pick;
            }
        }
        oh[idx] = (int32_t)h; ot[idx] = (int32_t)t; orel[idx] = (int32_t)r;
    }
}

int sample_subgraph(mre_ctx *ctx, const mre_index *ix, uint64_t seed, uint64_t step, uint32_t stream_id, const int64_t *edge_h,
                    const int64_t *edge_t, const int64_t *edge_r, int64_t n_edges, const int64_t *node_list, int64_t n_nodes,
                    const int64_t *local_to_global, int64_t n_local, int64_t neg, int32_t bern, int32_t filter, int32_t *out_h,
                    int32_t *out_t, int32_t *out_r, cudaStream_t st) {
    MRE_CHECK_ARG(ix != nullptr, "index is NULL");
    MRE_CHECK_ARG(ix->device == ctx->device, "index is not on device %d (call mre_index_to_device)", ctx->device);
    MRE_CHECK_ARG(n_edges >= 0 && neg >= 0 && n_nodes >= 0 && n_local >= 0, "negative size");
    MRE_CHECK_ARG(n_edges <= 0xffffffffLL && neg < 65536, "n_edges must fit 32 bits and neg_ent 16 bits of the Philox counter");
    MRE_CHECK_ARG(stream_id < 65536u, "stream_id must be < 65536");
    if (n_edges == 0) return MRE_OK;
    MRE_CHECK_ARG(edge_h && edge_t && edge_r && out_h && out_t && out_r, "NULL edge / output array");
    MRE_CHECK_ARG(n_nodes == 0 || node_list, "node_list is NULL");
    MRE_CHECK_ARG(!filter || local_to_global, "filtering needs the local -> global id map");
    SamplerTables T{ix->d_tr_h, ix->d_tr_r, ix->d_tr_t, ix->d_tr_hr_key, ix->d_tr_tr_key, ix->d_tr_tr_val, ix->d_bern_prob,
                    ix->n_train, ix->E, ix->R};
    const uint32_t c2 = (uint32_t)step, c3 = ((uint32_t)(step >> 32) & 0xffffu) | (stream_id << 16);
    const int64_t n = n_edges * (1 + neg);
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8);
    subgraph_sample_kernel<<<grid, 256, 0, st>>>(T, (uint32_t)seed, (uint32_t)(seed >> 32), c2, c3, edge_h, edge_t, edge_r, n_edges,
                                                 node_list, n_nodes, local_to_global, n_local, neg, bern, filter, out_h, out_t, out_r);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

}  // namespace mre
