// Shared declarations of libmre_b200.so (host side): error plumbing, the index and workspace structs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/mre_b200.h"

namespace mre {

void set_error(const char *fmt, ...);

#define MRE_CHECK_ARG(cond, ...)                  \
    do {                                          \
        if (!(cond)) {                            \
            mre::set_error(__VA_ARGS__);          \
            return MRE_ERR_INVALID;               \
        }                                         \
    } while (0)

#define MRE_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            mre::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return MRE_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define MRE_TRY(expr)            \
    do {                         \
        int _rc = (expr);        \
        if (_rc != MRE_OK) return _rc; \
    } while (0)

struct Triple {
    int64_t h, r, t;
};

// A device buffer that only ever grows; reallocation synchronises the device (rare: first calls only).
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <class T>
    T *as() const { return static_cast<T *>(p); }
};

struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <class T>
    T *as() const { return static_cast<T *>(p); }
};

}  // namespace mre

// Knowledge-graph index.  Host tables restate Reader.h; device tables are the packed-key CSR forms the
// kernels search.  Keys: head side key = h * R + r (sorted, with the tail as payload); tail side key = t * R + r.
struct mre_index {
    int64_t E = 0, R = 0;
    // ---- host, train side (Reader.h:53-160)
    std::vector<mre::Triple> train_head;  // de-duplicated, sorted (h,r,t) == trainList == trainHead
    std::vector<mre::Triple> train_tail;  // sorted (t,r,h)
    std::vector<float> left_mean, right_mean;  // tph, hpt
    std::vector<float> bern_prob;              // 1000*hpt/(hpt+tph) per relation, float32 (Base.cpp:113)
    // ---- host, test side (Reader.h:167-257)
    std::vector<mre::Triple> all_head;    // test + raw train + valid, sorted (h,r,t), duplicates kept (== tripleList)
    std::vector<mre::Triple> test, valid; // sorted (r,h,t)
    int64_t n_train_raw = 0;
    // ---- host, type constraints (Reader.h:267-317): per relation the sorted, de-duplicated head / tail candidate lists
    bool has_type = false;
    std::vector<int64_t> type_ptr[2];     // [side][R + 1], side 0 = head lists, 1 = tail lists
    std::vector<int64_t> type_idx[2];
    // ---- device (uploaded by mre_index_to_device)
    int device = -1;
    // filter tables over all splits, de-duplicated
    int64_t n_all = 0;
    int64_t *d_all_hr_key = nullptr, *d_all_hr_val = nullptr;  // key h*R+r sorted, val = t (sorted inside a run)
    int64_t *d_all_tr_key = nullptr, *d_all_tr_val = nullptr;  // key t*R+r sorted, val = h
    // sampler tables over de-duplicated train
    int64_t n_train = 0;
    int64_t *d_tr_h = nullptr, *d_tr_r = nullptr, *d_tr_t = nullptr;  // trainList columns, (h,r,t) order
    int64_t *d_tr_hr_key = nullptr;                                   // h*R+r of trainList (sorted); payload = d_tr_t
    int64_t *d_tr_tr_key = nullptr, *d_tr_tr_val = nullptr;           // (t,r,h) order: key t*R+r, val = h
    float *d_bern_prob = nullptr;
    int64_t *d_type_ptr[2] = {nullptr, nullptr}, *d_type_idx[2] = {nullptr, nullptr};   // type-constraint lists (when loaded)
};

struct mre_ctx {
    int device = 0;
    int sm_count = 0;
    std::atomic<int64_t> launches{0};
    // scratch (grow-only)
    mre::DevBuf ent_n, rel_n;        // normalised / padded tables
    mre::DevBuf ent_aux, ent_aux2;   // bilinear: tf32 hi/lo splits
    mre::DevBuf qvec, qvec2;         // per-query vectors (hi/lo for bilinear)
    mre::DevBuf thr;                 // per-query thresholds (2 floats) + true scores
    mre::DevBuf tiles;               // tile descriptors
    mre::DevBuf sched;               // cost-balanced work-item order of the TransE kernel (candidate groups)
    mre::DevBuf known_score, known_stamp, known_range;   // shared-run known-true pass (MRE_FILTER_INDEX): see rank_common.cuh KnownRuns
    unsigned int known_epoch = 0;
    size_t known_n = 0;
    mre::DevBuf counters;            // raw/corr counters, work counters
    mre::DevBuf misc;                // loss partials etc.
    mre::DevBuf misc2;               // known-true pair list of the tile filter
    mre::DevBuf met_scratch;         // mre_metrics: per-CTA float64 partials + ticket counter
    mre::DevBuf stats;               // device-side statistics counters (mre_ctx_stat)
    mre::DevBuf loss_acc;            // float64 loss accumulator + finished-block counter of ns_loss (self re-arming)
    mre::DevBuf stage_dev;           // device staging for *_host entry points
    mre::PinnedBuf stage_pin;        // pinned host staging
    // kernels whose dynamic shared-memory opt-in has been set on THIS context's device (function attributes are per device)
    std::vector<const void *> smem_ready;
    int allow_smem(const void *func, size_t bytes);
    // tunables (mre_ctx_option): BF16/FP16 MMAs per product of the bilinear path, CTA pairs on/off, persistent CTAs per SM of the
    // TransE kernel, FP32 fallback of the ZSL scorer -- developer A/B switches, read from the context, never from the environment
    int opt_bil_products = 3, opt_bil_pair = 1, opt_transe_ctas = 0, opt_zsl_fp32 = 0, opt_transe_lpt = 1;
    int zsl_const_slot = 0;          // this context's slot of the ZSL kernel's constant-bank vectors (zsl_rank.cu)
    unsigned int bil_epoch = 0;      // generation tag of the bilinear path's max-row-norm slot (no reset launch per call)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // optional per-launch timing of the dominant (rank) kernel: event pairs recorded on the launching stream
    bool timing = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    int time_begin(cudaStream_t st);
    int time_end(cudaStream_t st);
    // second stream for work that does not depend on the query vectors (the known-true tile filter): forked after the
    // caller's stream has the job's descriptors in flight, joined before the rank kernel
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int fork_aux(cudaStream_t st, cudaStream_t *aux_out);
    int join_aux(cudaStream_t st);
};

namespace mre {
int read_benchmark_dir(const std::string &dir, int64_t *E, int64_t *R, std::vector<Triple> &tr, std::vector<Triple> &va, std::vector<Triple> &te);   // index.cpp
// implemented in the .cu files
int rank_transe(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, cudaStream_t st);
int rank_bilinear(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, cudaStream_t st);
int rank_rotate(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, cudaStream_t st);
int predict_transe(mre_ctx *ctx, const mre_rank_job *job, int64_t query, float *scores_out, cudaStream_t st);
int predict_bilinear(mre_ctx *ctx, const mre_rank_job *job, int64_t query, float *scores_out, cudaStream_t st);
int bilinear_scores(mre_ctx *ctx, const mre_rank_job *job, float *scores_out, cudaStream_t st);
int metrics(mre_ctx *ctx, const int32_t *counts, const uint8_t *q_side, int32_t side, int64_t Q, int32_t rank_mode,
            int32_t raw, int64_t *sums_out, double *rr_out, int64_t *hist, int64_t hist_len, cudaStream_t st);
int sample(mre_ctx *ctx, const mre_index *ix, uint64_t seed, uint64_t step, uint32_t stream_id, int64_t B, int64_t neg,
           int32_t mode, int32_t bern, int64_t *h, int64_t *t, int64_t *r, float *y, cudaStream_t st);
int sample_subgraph(mre_ctx *ctx, const mre_index *ix, uint64_t seed, uint64_t step, uint32_t stream_id, const int64_t *edge_h,
                    const int64_t *edge_t, const int64_t *edge_r, int64_t n_edges, const int64_t *node_list, int64_t n_nodes,
                    const int64_t *local_to_global, int64_t n_local, int64_t neg, int32_t bern, int32_t filter, int32_t *out_h,
                    int32_t *out_t, int32_t *out_r, cudaStream_t st);
int corrupt_typed(mre_ctx *ctx, const mre_index *ix, uint64_t seed, uint64_t step, uint32_t stream_id, const int64_t *h,
                  const int64_t *r, int64_t n, int64_t *t_out, cudaStream_t st);
int transe_margin_step(mre_ctx *ctx, const float *ent, const float *rel, int64_t E, int64_t R, int64_t D,
                       const int64_t *h, const int64_t *t, const int64_t *r, int64_t B, int64_t neg, float margin,
                       int32_t p_norm, int32_t normalize, float *grad_ent, float *grad_rel, float *loss_out,
                       float *scores_out, cudaStream_t st);
int sgd_update(mre_ctx *ctx, float *w, float *g, int64_t n, float lr, cudaStream_t st);
int score_triples(mre_ctx *ctx, int scorer, const float *ent, const float *ent_im, const float *rel, const float *rel_im, int64_t D,
                  const int64_t *h, const int64_t *t, const int64_t *r, int64_t n, int32_t p_norm, int32_t normalize, float *score,
                  cudaStream_t st);
int transe_backward(mre_ctx *ctx, const float *ent, const float *rel, int64_t D, const int64_t *h, const int64_t *t, const int64_t *r,
                    int64_t n, int32_t p_norm, int32_t normalize, const float *score, const float *dscore, float *grad_ent,
                    float *grad_rel, cudaStream_t st);
int ns_loss(mre_ctx *ctx, int32_t kind, const float *score, int64_t B, int64_t neg, float margin, int32_t adv, float temperature,
            float *loss_out, float *dscore, cudaStream_t st);
int bilinear_backward(mre_ctx *ctx, int scorer, const float *ent, const float *ent_im, const float *rel, const float *rel_im, int64_t D,
                      const int64_t *h, const int64_t *t, const int64_t *r, int64_t n, const float *dscore, float *g_ent,
                      float *g_ent_im, float *g_rel, float *g_rel_im, cudaStream_t st);
int ns_train_step(mre_ctx *ctx, int32_t scorer, const float *ent, const float *ent_im, const float *rel, const float *rel_im,
                  int64_t D, const int64_t *h, const int64_t *t, const int64_t *r, int64_t B, int64_t neg, int32_t loss_kind,
                  float margin, int32_t adv, float temperature, int32_t p_norm, int32_t normalize, float *g_ent, float *g_ent_im,
                  float *g_rel, float *g_rel_im, float *loss_out, float *scores_out, cudaStream_t st);
int zsl_entity_features(mre_ctx *ctx, const mre_zsl_model *m, const int64_t *ent_symbol, const int64_t *conn, const float *deg,
                        int64_t n_ent, int32_t max_nb, float *A, float *B, cudaStream_t st);
int zsl_rank(mre_ctx *ctx, const mre_zsl_model *m, const float *A, const float *B, int64_t n_ent, const int64_t *q_head,
             const int64_t *q_rel, const int64_t *cand_ptr, const int64_t *cand_idx, int64_t T, int64_t P, const float *rel_vecs,
             int64_t n_rel, int32_t n_vec, float *scores, int32_t *counts, cudaStream_t st);
int probe_fp32_peak(mre_ctx *ctx, double *lane_ops_per_s);
int probe_mufu_peak(mre_ctx *ctx, double *ops_per_s);
int probe_tf32_peak(mre_ctx *ctx, double *flops_per_s);
int probe_bf16_peak(mre_ctx *ctx, double *flops_per_s);
}  // namespace mre
