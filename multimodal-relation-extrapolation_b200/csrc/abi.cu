// extern "C" entry points of libmre_b200.so that dispatch to the kernels (see include/mre_b200.h), the
// host-buffer (end-to-end) variants, per-launch timing and the FP32 issue-rate probe.
#include <algorithm>

#include "common.h"

int mre_ctx::time_begin(cudaStream_t st) {
    if (!timing) return MRE_OK;
    if (ev_used + 2 > ev_pool.size()) {
        for (int i = 0; i < 2; i++) {
            cudaEvent_t e;
            MRE_CUDA(cudaEventCreate(&e));
            ev_pool.push_back(e);
        }
    }
    MRE_CUDA(cudaEventRecord(ev_pool[ev_used], st));
    return MRE_OK;
}
int mre_ctx::time_end(cudaStream_t st) {
    if (!timing) return MRE_OK;
    MRE_CUDA(cudaEventRecord(ev_pool[ev_used + 1], st));
    ev_used += 2;
    return MRE_OK;
}

int mre_ctx::allow_smem(const void *func, size_t bytes) {
    for (const void *f : smem_ready)
        if (f == func) return MRE_OK;
    MRE_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    smem_ready.push_back(func);
    return MRE_OK;
}

int mre_ctx::fork_aux(cudaStream_t st, cudaStream_t *aux_out) {
    if (!aux) {
        MRE_CUDA(cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking));
        MRE_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        MRE_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    }
    MRE_CUDA(cudaEventRecord(ev_fork, st));
    MRE_CUDA(cudaStreamWaitEvent(aux, ev_fork, 0));
    *aux_out = aux;
    return MRE_OK;
}
int mre_ctx::join_aux(cudaStream_t st) {
    MRE_CUDA(cudaEventRecord(ev_join, aux));
    MRE_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
    return MRE_OK;
}

namespace mre {

static int check_job(const mre_rank_job *job) {
    MRE_CHECK_ARG(job != nullptr, "job is NULL");
    MRE_CHECK_ARG(job->E > 0 && job->R > 0 && job->D > 0, "E, R, D must be positive");
    MRE_CHECK_ARG(job->Q >= 0, "Q must be non-negative");
    MRE_CHECK_ARG(job->E < (1LL << 31) && job->Q < (1LL << 31), "E and Q must fit int32 counts");
    MRE_CHECK_ARG(job->ent && job->rel, "ent / rel table is NULL");
    MRE_CHECK_ARG(job->scorer >= MRE_TRANSE && job->scorer <= MRE_ROTATE, "unknown scorer %d", job->scorer);
    MRE_CHECK_ARG(job->scorer != MRE_COMPLEX || (job->ent_im && job->rel_im), "ComplEx needs ent_im and rel_im");
    MRE_CHECK_ARG(job->filter >= MRE_FILTER_NONE && job->filter <= MRE_FILTER_CSR, "unknown filter %d", job->filter);
    MRE_CHECK_ARG(job->side == 0 || job->side == 1, "side must be 0 or 1");
    MRE_CHECK_ARG(job->Q == 0 || (job->q_h && job->q_t && job->q_r), "query arrays are NULL");
    MRE_CHECK_ARG(job->Q == 0 || job->counts, "counts is NULL");
    MRE_CHECK_ARG(((uintptr_t)job->ent & 15) == 0 && ((uintptr_t)job->rel & 15) == 0, "tables must be 16-byte aligned");
    return MRE_OK;
}

static int dispatch_rank(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, cudaStream_t st) {
    if (job->scorer == MRE_TRANSE) return rank_transe(ctx, ix, job, st);
    if (job->scorer == MRE_ROTATE) return rank_rotate(ctx, ix, job, st);
    return rank_bilinear(ctx, ix, job, st);
}

// -------------------------------------------------------------------------------- FP32 FADD issue-rate probe
constexpr int PROBE_ACC = 8;
__global__ void __launch_bounds__(256) fadd_probe_kernel(float *out, int iters, float c) {
    float a[PROBE_ACC];
#pragma unroll
    for (int i = 0; i < PROBE_ACC; i++) a[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < PROBE_ACC; i++) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PROBE_ACC; i++) s += a[i];
    if (s == 12345.678f) out[0] = s;  // never true in practice; keeps the chain live
}

// the packed form (add.f32x2 -> FADD2: two lane-ops per lane per instruction, two pipe cycles): what the TransE tile loop issues
__global__ void __launch_bounds__(256) fadd2_probe_kernel(float *out, int iters, float c) {
    unsigned long long a[PROBE_ACC];
#pragma unroll
    for (int i = 0; i < PROBE_ACC; i++) a[i] = (unsigned long long)(threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < PROBE_ACC; i++)
                asm volatile("{.reg .b64 t; mov.b64 t, {%1, %1}; add.f32x2 %0, %0, t;}" : "+l"(a[i]) : "f"(c));
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < PROBE_ACC; i++) s += a[i];
    if (s == 12345ull) out[0] = 1.f;  // never true in practice; keeps the chain live
}

// MUFU.SQRT stream (sqrt.approx.ftz): sixteen independent chains per thread, nothing else on the way -- the RotatE tile loop's bound
__global__ void __launch_bounds__(256) mufu_probe_kernel(float *out, int iters, float c) {
    float a[PROBE_ACC];
#pragma unroll
    for (int i = 0; i < PROBE_ACC; i++) a[i] = c + (float)(threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < PROBE_ACC; i++) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PROBE_ACC; i++) s += a[i];
    if (s == 12345.678f) out[0] = s;  // never true in practice; keeps the chain live
}

int probe_mufu_peak(mre_ctx *ctx, double *ops_per_s) {
    MRE_CHECK_ARG(ops_per_s != nullptr, "NULL output");
    MRE_TRY(ctx->misc.reserve(256));
    const int iters = 1024, blocks = ctx->sm_count * 8, threads = 256;
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        MRE_CUDA(cudaEventRecord(ctx->ev0, 0));
        mufu_probe_kernel<<<blocks, threads>>>(ctx->misc.as<float>(), iters, 1.5f);
        MRE_CUDA(cudaEventRecord(ctx->ev1, 0));
        MRE_CUDA(cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        MRE_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double ops = (double)blocks * threads * iters * 8.0 * PROBE_ACC;
        if (rep > 0) best = std::max(best, ops / (ms * 1e-3));
    }
    ctx->launches += 5;
    *ops_per_s = best;
    return MRE_OK;
}

int probe_fp32_peak(mre_ctx *ctx, double *lane_ops_per_s) {
    MRE_CHECK_ARG(lane_ops_per_s != nullptr, "NULL output");
    MRE_TRY(ctx->misc.reserve(256));
    const int iters = 4096, blocks = ctx->sm_count * 8, threads = 256;
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        MRE_CUDA(cudaEventRecord(ctx->ev0, 0));
        fadd_probe_kernel<<<blocks, threads>>>(ctx->misc.as<float>(), iters, 1e-7f);
        MRE_CUDA(cudaEventRecord(ctx->ev1, 0));
        MRE_CUDA(cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        MRE_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        double ops = (double)blocks * threads * iters * 8.0 * PROBE_ACC;
        if (rep > 0) best = std::max(best, ops / (ms * 1e-3));
    }
    for (int rep = 0; rep < 5; rep++) {   // the peak is the better of the scalar and the packed instruction streams
        MRE_CUDA(cudaEventRecord(ctx->ev0, 0));
        fadd2_probe_kernel<<<blocks, threads>>>(ctx->misc.as<float>(), iters, 1e-7f);
        MRE_CUDA(cudaEventRecord(ctx->ev1, 0));
        MRE_CUDA(cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        MRE_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        double ops = (double)blocks * threads * iters * 8.0 * PROBE_ACC * 2.0;
        if (rep > 0) best = std::max(best, ops / (ms * 1e-3));
    }
    ctx->launches += 10;
    *lane_ops_per_s = best;
    return MRE_OK;
}

}  // namespace mre

using namespace mre;

extern "C" {

int mre_rank(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, void *stream) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_TRY(check_job(job));
    MRE_CUDA(cudaSetDevice(ctx->device));
    return dispatch_rank(ctx, ix, job, (cudaStream_t)stream);
}

int mre_rank_host(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, void *stream) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_TRY(check_job(job));
    MRE_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t Q = job->Q;
    if (Q == 0) return MRE_OK;
    // device staging: [h | t | r] int64, counts int32 [4][Q], side bytes
    const size_t ids = (size_t)Q * sizeof(int64_t);
    const size_t cnt = (size_t)4 * Q * sizeof(int32_t);
    MRE_TRY(ctx->stage_dev.reserve(3 * ids + cnt + (size_t)Q));
    char *base = ctx->stage_dev.as<char>();
    int64_t *d_h = (int64_t *)base, *d_t = d_h + Q, *d_r = d_t + Q;
    int32_t *d_counts = (int32_t *)(base + 3 * ids);
    uint8_t *d_side = (uint8_t *)(base + 3 * ids + cnt);
    MRE_CUDA(cudaMemcpyAsync(d_h, job->q_h, ids, cudaMemcpyHostToDevice, st));
    MRE_CUDA(cudaMemcpyAsync(d_t, job->q_t, ids, cudaMemcpyHostToDevice, st));
    MRE_CUDA(cudaMemcpyAsync(d_r, job->q_r, ids, cudaMemcpyHostToDevice, st));
    if (job->q_side) MRE_CUDA(cudaMemcpyAsync(d_side, job->q_side, (size_t)Q, cudaMemcpyHostToDevice, st));
    mre_rank_job dev = *job;
    dev.q_h = d_h; dev.q_t = d_t; dev.q_r = d_r;
    dev.q_side = job->q_side ? d_side : nullptr;
    dev.counts = d_counts;
    MRE_TRY(dispatch_rank(ctx, ix, &dev, st));
    MRE_CUDA(cudaMemcpyAsync(job->counts, d_counts, cnt, cudaMemcpyDeviceToHost, st));
    MRE_CUDA(cudaStreamSynchronize(st));
    return MRE_OK;
}

int mre_predict(mre_ctx *ctx, const mre_rank_job *job, int64_t query, float *scores_out, void *stream) {
    MRE_CHECK_ARG(ctx != nullptr && scores_out != nullptr, "NULL argument");
    mre_rank_job j = *job;
    int32_t dummy = 0;
    if (!j.counts) j.counts = &dummy;  // unused by predict; keeps check_job happy
    MRE_TRY(check_job(&j));
    MRE_CHECK_ARG(query >= 0 && query < job->Q, "query %lld out of range", (long long)query);
    MRE_CHECK_ARG(job->scorer != MRE_ROTATE, "MRE_ROTATE is served by mre_rank only (RotatE.predict on explicit batches stays with the host mirror)");
    MRE_CUDA(cudaSetDevice(ctx->device));
    if (job->scorer == MRE_TRANSE) return predict_transe(ctx, job, query, scores_out, (cudaStream_t)stream);
    return predict_bilinear(ctx, job, query, scores_out, (cudaStream_t)stream);
}

int mre_bilinear_scores(mre_ctx *ctx, const mre_rank_job *job, float *scores_out, void *stream) {
    MRE_CHECK_ARG(ctx != nullptr && scores_out != nullptr, "NULL argument");
    mre_rank_job j = *job;
    int32_t dummy = 0;
    if (!j.counts) j.counts = &dummy;
    MRE_TRY(check_job(&j));
    MRE_CHECK_ARG(job->scorer == MRE_DISTMULT || job->scorer == MRE_COMPLEX, "mre_bilinear_scores is for the DistMult / ComplEx (tensor-core) path");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return bilinear_scores(ctx, job, scores_out, (cudaStream_t)stream);
}

int mre_metrics(mre_ctx *ctx, const int32_t *counts, const uint8_t *q_side, int32_t side, int64_t Q, int32_t rank_mode,
                int32_t raw, int64_t *sums_out, double *rr_out, int64_t *hist, int64_t hist_len, void *stream) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return metrics(ctx, counts, q_side, side, Q, rank_mode, raw, sums_out, rr_out, hist, hist_len, (cudaStream_t)stream);
}

int mre_sample(mre_ctx *ctx, const mre_index *ix, uint64_t seed, uint64_t step, uint32_t stream_id, int64_t B, int64_t neg,
               int32_t mode, int32_t bern, int64_t *h, int64_t *t, int64_t *r, float *y, void *stream) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return sample(ctx, ix, seed, step, stream_id, B, neg, mode, bern, h, t, r, y, (cudaStream_t)stream);
}

int mre_sample_host(mre_ctx *ctx, const mre_index *ix, uint64_t seed, uint64_t step, uint32_t stream_id, int64_t B,
                    int64_t neg, int32_t mode, int32_t bern, int64_t *h, int64_t *t, int64_t *r, float *y, void *stream) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CHECK_ARG(B > 0 && neg >= 0, "B must be positive and neg non-negative");
    MRE_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)B * (size_t)(1 + neg);
    MRE_TRY(ctx->stage_dev.reserve(n * (3 * sizeof(int64_t) + sizeof(float))));
    int64_t *dh = ctx->stage_dev.as<int64_t>(), *dt = dh + n, *dr = dt + n;
    float *dy = (float *)(dr + n);
    MRE_TRY(sample(ctx, ix, seed, step, stream_id, B, neg, mode, bern, dh, dt, dr, dy, st));
    MRE_CUDA(cudaMemcpyAsync(h, dh, n * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    MRE_CUDA(cudaMemcpyAsync(t, dt, n * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    MRE_CUDA(cudaMemcpyAsync(r, dr, n * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    MRE_CUDA(cudaMemcpyAsync(y, dy, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    MRE_CUDA(cudaStreamSynchronize(st));
    return MRE_OK;
}

int mre_sample_subgraph(mre_ctx *ctx, const mre_index *ix, uint64_t seed, uint64_t step, uint32_t stream_id, const int64_t *edge_h,
                        const int64_t *edge_t, const int64_t *edge_r, int64_t n_edges, const int64_t *node_list, int64_t n_nodes,
                        const int64_t *local_to_global, int64_t n_local, int64_t neg_ent, int32_t bern, int32_t filter,
                        int32_t *out_h, int32_t *out_t, int32_t *out_r, void *stream) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return sample_subgraph(ctx, ix, seed, step, stream_id, edge_h, edge_t, edge_r, n_edges, node_list, n_nodes, local_to_global,
                           n_local, neg_ent, bern, filter, out_h, out_t, out_r, (cudaStream_t)stream);
}

int mre_corrupt_typed(mre_ctx *ctx, const mre_index *ix, uint64_t seed, uint64_t step, uint32_t stream_id, const int64_t *h,
                      const int64_t *r, int64_t n, int64_t *t_out, void *stream) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return corrupt_typed(ctx, ix, seed, step, stream_id, h, r, n, t_out, (cudaStream_t)stream);
}

int mre_transe_margin_step(mre_ctx *ctx, const float *ent, const float *rel, int64_t E, int64_t R, int64_t D, const int64_t *h,
                           const int64_t *t, const int64_t *r, int64_t B, int64_t neg, float margin, int32_t p_norm,
                           int32_t normalize, float *grad_ent, float *grad_rel, float *loss_out, float *scores_out,
                           void *stream) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return transe_margin_step(ctx, ent, rel, E, R, D, h, t, r, B, neg, margin, p_norm, normalize, grad_ent, grad_rel, loss_out,
                              scores_out, (cudaStream_t)stream);
}

int mre_score_triples(mre_ctx *ctx, int32_t scorer, const float *ent, const float *ent_im, const float *rel, const float *rel_im,
                      int64_t D, const int64_t *h, const int64_t *t, const int64_t *r, int64_t n, int32_t p_norm, int32_t normalize,
                      float *score_out, void *stream) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return score_triples(ctx, scorer, ent, ent_im, rel, rel_im, D, h, t, r, n, p_norm, normalize, score_out, (cudaStream_t)stream);
}

int mre_transe_backward(mre_ctx *ctx, const float *ent, const float *rel, int64_t D, const int64_t *h, const int64_t *t,
                        const int64_t *r, int64_t n, int32_t p_norm, int32_t normalize, const float *score, const float *dscore,
                        float *grad_ent, float *grad_rel, void *stream) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return transe_backward(ctx, ent, rel, D, h, t, r, n, p_norm, normalize, score, dscore, grad_ent, grad_rel, (cudaStream_t)stream);
}

int mre_bilinear_backward(mre_ctx *ctx, int32_t scorer, const float *ent, const float *ent_im, const float *rel, const float *rel_im,
                          int64_t D, const int64_t *h, const int64_t *t, const int64_t *r, int64_t n, const float *dscore,
                          float *g_ent, float *g_ent_im, float *g_rel, float *g_rel_im, void *stream) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return bilinear_backward(ctx, scorer, ent, ent_im, rel, rel_im, D, h, t, r, n, dscore, g_ent, g_ent_im, g_rel, g_rel_im,
                             (cudaStream_t)stream);
}

int mre_ns_loss(mre_ctx *ctx, int32_t kind, const float *score, int64_t B, int64_t neg, float margin, int32_t adv,
                float adv_temperature, float *loss_out, float *dscore, void *stream) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return ns_loss(ctx, kind, score, B, neg, margin, adv, adv_temperature, loss_out, dscore, (cudaStream_t)stream);
}

int mre_ns_train_step(mre_ctx *ctx, int32_t scorer, const float *ent, const float *ent_im, const float *rel, const float *rel_im,
                      int64_t D, const int64_t *h, const int64_t *t, const int64_t *r, int64_t B, int64_t neg, int32_t loss_kind,
                      float margin, int32_t adv, float adv_temperature, int32_t p_norm, int32_t normalize,
                      float *g_ent, float *g_ent_im, float *g_rel, float *g_rel_im, float *loss_out, float *scores_out, void *stream) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return ns_train_step(ctx, scorer, ent, ent_im, rel, rel_im, D, h, t, r, B, neg, loss_kind, margin, adv, adv_temperature, p_norm,
                         normalize, g_ent, g_ent_im, g_rel, g_rel_im, loss_out, scores_out, (cudaStream_t)stream);
}

int mre_sgd_update(mre_ctx *ctx, float *w, float *g, int64_t n, float lr, void *stream) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return sgd_update(ctx, w, g, n, lr, (cudaStream_t)stream);
}

int mre_zsl_entity_features(mre_ctx *ctx, const mre_zsl_model *model, const int64_t *ent_symbol, const int64_t *conn, const float *deg,
                            int64_t n_ent, int32_t max_neighbor, float *A, float *B, void *stream) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return zsl_entity_features(ctx, model, ent_symbol, conn, deg, n_ent, max_neighbor, A, B, (cudaStream_t)stream);
}

int mre_zsl_rank(mre_ctx *ctx, const mre_zsl_model *model, const float *A, const float *B, int64_t n_ent, const int64_t *q_head,
                 const int64_t *q_rel, const int64_t *cand_ptr, const int64_t *cand_idx, int64_t T, int64_t P, const float *rel_vecs,
                 int64_t n_rel, int32_t n_vec, float *scores, int32_t *counts, void *stream) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return zsl_rank(ctx, model, A, B, n_ent, q_head, q_rel, cand_ptr, cand_idx, T, P, rel_vecs, n_rel, n_vec, scores, counts, (cudaStream_t)stream);
}

int mre_probe_fp32_peak(mre_ctx *ctx, double *lane_ops_per_s) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return probe_fp32_peak(ctx, lane_ops_per_s);
}

int mre_probe_mufu_peak(mre_ctx *ctx, double *ops_per_s) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return probe_mufu_peak(ctx, ops_per_s);
}

int mre_probe_tf32_peak(mre_ctx *ctx, double *flops_per_s) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return probe_tf32_peak(ctx, flops_per_s);
}

int mre_probe_bf16_peak(mre_ctx *ctx, double *flops_per_s) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    MRE_CUDA(cudaSetDevice(ctx->device));
    return probe_bf16_peak(ctx, flops_per_s);
}

int mre_ctx_option(mre_ctx *ctx, const char *key, int64_t value) {
    MRE_CHECK_ARG(ctx != nullptr && key != nullptr, "NULL argument");
    const std::string k(key);
    if (k == "bil_products") {
        MRE_CHECK_ARG(value == 1 || value == 3, "bil_products must be 1 (FP16 single product) or 3 (BF16 hi/lo split)");
        ctx->opt_bil_products = (int)value;
    } else if (k == "bil_pair") ctx->opt_bil_pair = value != 0;
    else if (k == "transe_ctas_per_sm") ctx->opt_transe_ctas = (int)std::max<int64_t>(0, value);
    else if (k == "zsl_fp32") ctx->opt_zsl_fp32 = value != 0;
    else if (k == "transe_lpt") ctx->opt_transe_lpt = value != 0;
    else {
        set_error("unknown option '%s'", key);
        return MRE_ERR_INVALID;
    }
    return MRE_OK;
}

int mre_ctx_stat(mre_ctx *ctx, const char *key, int64_t *value) {
    MRE_CHECK_ARG(ctx != nullptr && key != nullptr && value != nullptr, "NULL argument");
    MRE_CHECK_ARG(std::string(key) == "bil_rescored", "unknown statistic '%s'", key);
    *value = 0;
    if (!ctx->stats.p) return MRE_OK;
    MRE_CUDA(cudaSetDevice(ctx->device));
    unsigned long long v = 0;
    MRE_CUDA(cudaDeviceSynchronize());
    MRE_CUDA(cudaMemcpy(&v, ctx->stats.p, sizeof(v), cudaMemcpyDeviceToHost));
    MRE_CUDA(cudaMemset(ctx->stats.p, 0, sizeof(v)));
    *value = (int64_t)v;
    return MRE_OK;
}

int mre_ctx_timing(mre_ctx *ctx, int32_t enable) {
    MRE_CHECK_ARG(ctx != nullptr, "ctx is NULL");
    ctx->timing = enable != 0;
    return MRE_OK;
}

int mre_ctx_timing_read(mre_ctx *ctx, double *total_ms, int64_t *n_launches) {
    MRE_CHECK_ARG(ctx && total_ms && n_launches, "NULL argument");
    MRE_CUDA(cudaSetDevice(ctx->device));
    double total = 0;
    for (size_t i = 0; i + 1 < ctx->ev_used; i += 2) {
        MRE_CUDA(cudaEventSynchronize(ctx->ev_pool[i + 1]));
        float ms = 0;
        MRE_CUDA(cudaEventElapsedTime(&ms, ctx->ev_pool[i], ctx->ev_pool[i + 1]));
        total += ms;
    }
    *total_ms = total;
    *n_launches = (int64_t)(ctx->ev_used / 2);
    ctx->ev_used = 0;
    return MRE_OK;
}

}  // extern "C"
