/*
 * mre_b200.h -- C ABI of libmre_b200.so: the B200 (sm_100a) link-prediction scorer + filtered ranker,
 * Bernoulli negative sampler and TransE margin-loss step.
 *
 * This is the drop-in boundary for the reference's hot path.  The reference's own native boundary is
 * OpenKE's Base.so (all paths below are relative to /root/reference):
 *   OpenKE/openke/base/Setting.h:17-145   setInPath / setBern / setWorkThreads / get*Total
 *   OpenKE/openke/base/Reader.h:53-257    importTrainFiles / importTestFiles
 *   OpenKE/openke/base/Base.cpp:161-197   sampling(h,t,r,y,B,negRate,negRelRate,mode,filter,p,val_loss)
 *   OpenKE/openke/base/Test.h:22-390      initTest / getHeadBatch / getTailBatch / testHead / testTail /
 *                                         test_link_prediction / getTestLink{MRR,MR,Hit10,Hit3,Hit1}
 * That cut hands a host float[E] score vector per query to testHead/testTail, which is exactly what
 * forces the query x entity score matrix through host memory.  This ABI re-cuts the same path one level
 * up: the caller hands (h, r, t, side) query ids and the embedding tables, and gets integer rank counts
 * back; scores never leave the SM.  Every entry point names the reference interface it replaces.
 *
 * Conventions
 *   - plain C, no torch types; opaque handles; every function returns MRE_OK (0) or a negative code and
 *     records a message retrievable with mre_last_error() (thread-local).
 *   - "device" pointers are CUDA device pointers owned by the caller (e.g. torch storages); `stream` is
 *     a cudaStream_t passed as void* (NULL = legacy default stream).  Functions taking device pointers
 *     are asynchronous on `stream` unless stated otherwise.
 *   - ids are int64 (the reference's INT = long, Setting.h:3); scores float32 (REAL = float, :4).
 *   - side: 0 = head query (?, r, t) -- candidates replace the head (Test.h:65-127, mode "head_batch");
 *           1 = tail query (h, r, ?) -- candidates replace the tail (Test.h:130-192, mode "tail_batch").
 *   - score orientation: LOWER IS BETTER for every scorer, exactly as Model.predict returns
 *     (TransE.py:88-94 distance; DistMult.py:70-72 and ComplEx.py:60-61 negate the similarity).
 *   - there is NO CPU fallback: every compute entry point fails with MRE_ERR_CUDA without a B200.
 */
#ifndef MRE_B200_H
#define MRE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRE_ABI_VERSION 1

/* error codes */
#define MRE_OK               0
#define MRE_ERR_INVALID     -1   /* bad argument */
#define MRE_ERR_CUDA        -2   /* CUDA runtime error / no device */
#define MRE_ERR_NOMEM       -3
#define MRE_ERR_IO          -4   /* benchmark file missing / malformed */
#define MRE_ERR_UNSUPPORTED -5

/* scorers */
#define MRE_TRANSE   0   /* OpenKE/openke/module/model/TransE.py:46-60, module/NegativeSampling.py:142-157 */
#define MRE_DISTMULT 1   /* OpenKE/openke/module/model/DistMult.py:34-44 */
#define MRE_COMPLEX  2   /* OpenKE/openke/module/model/ComplEx.py:20-27 */
#define MRE_ROTATE   3   /* OpenKE/openke/module/model/RotatE.py:44-78: ent rows [re | im] (D = 2 dim), rel rows = phases (dim);
                            mre_rank only (forward on explicit triples stays with the host mirror) */

/* filter sources */
#define MRE_FILTER_NONE  0   /* raw ranks only */
#define MRE_FILTER_INDEX 1   /* train+valid+test membership from the index == _find, Corrupt.h:166-177 */
#define MRE_FILTER_CSR   2   /* caller-supplied per-query known-true lists (paper path: e1rel_e2, utils/gen_mode_candidates.py:30-34) */

/* rank conventions (the reference has three) */
#define MRE_RANK_STRICT      0   /* 1 + #(s_j <  s_true)                 Test.h:80-112 */
#define MRE_RANK_TIES_HALF   1   /* 1 + #(s_j <  s_true) + #(s_j == s_true)/2   main.py:245-250 */
#define MRE_RANK_PESSIMISTIC 2   /* 1 + #(s_j <= s_true), upper end of module/zsl_module.py:705-706's unpinned argsort */

/* negative-sampling losses (mre_ns_loss) */
#define MRE_LOSS_MARGIN   0   /* OpenKE/openke/module/loss/MarginLoss.py:24-28, module/loss.py:20-24 */
#define MRE_LOSS_SIGMOID  1   /* OpenKE/openke/module/loss/SigmoidLoss.py:22-26 */
#define MRE_LOSS_SOFTPLUS 2   /* OpenKE/openke/module/loss/SoftplusLoss.py:22-26 */

/* index totals */
#define MRE_TOTAL_ENTITY   0   /* getEntityTotal   Setting.h:107-110 */
#define MRE_TOTAL_RELATION 1   /* getRelationTotal */
#define MRE_TOTAL_TRAIN    2   /* getTrainTotal (after de-duplication, Reader.h:92-105) */
#define MRE_TOTAL_VALID    3   /* getValidTotal */
#define MRE_TOTAL_TEST     4   /* getTestTotal */
#define MRE_TOTAL_TRIPLE   5   /* getTripleTotal = test + raw train + valid */

#define MRE_SPLIT_TRAIN 0      /* de-duplicated, sorted (h,r,t)  == trainList, Reader.h:107 */
#define MRE_SPLIT_VALID 1      /* sorted (r,h,t), Reader.h:228 */
#define MRE_SPLIT_TEST  2      /* sorted (r,h,t), Reader.h:227 -- the order Tester iterates */

typedef struct mre_index mre_index;   /* knowledge-graph index (Reader.h tables), host + device copies */
typedef struct mre_ctx   mre_ctx;     /* per-device workspace: scratch buffers, pinned staging, SM count */

/* ---------------------------------------------------------------------------------------------- misc */
const char *mre_last_error(void);
int         mre_abi_version(void);
/* 0 when a CUDA device with compute capability 10.x is visible, MRE_ERR_CUDA otherwise */
int         mre_device_ok(int device);

/* --------------------------------------------------------------------------------------------- index */
/*
 * Build the index from id triples given as (h, t, r) column arrays in host memory (the column order of
 * OpenKE's *2id.txt files, OpenKE/README.md:126-141).  Restates importTrainFiles (Reader.h:53-160:
 * de-dup, (h,r,t) and (t,r,h) orders, tph/hpt) and importTestFiles (Reader.h:167-257: the all-splits
 * membership list, test/valid sorted by (r,h,t)).  Host-only; no device needed.
 */
int mre_index_create(int64_t E, int64_t R,
                     const int64_t *train_h, const int64_t *train_t, const int64_t *train_r, int64_t n_train,
                     const int64_t *valid_h, const int64_t *valid_t, const int64_t *valid_r, int64_t n_valid,
                     const int64_t *test_h, const int64_t *test_t, const int64_t *test_r, int64_t n_test,
                     mre_index **out);
/* Same, reading entity2id.txt / relation2id.txt / {train,valid,test}2id.txt under `in_path`
 * (setInPath + importTrainFiles + importTestFiles, Setting.h:17-27, Reader.h:53-257).
 * valid2id.txt / test2id.txt may be absent (training-only use). */
int mre_index_create_from_dir(const char *in_path, mre_index **out);
/*
 * mre_index_create + mre_index_to_device(device) with every sort of the build on the GPU (csrc/index_build.cu): the three
 * std::sort calls of importTrainFiles / importTestFiles (Reader.h:91-109, 201-227), the de-duplication (Reader.h:91-105) and the
 * freqRel / distinct-pair counters behind tph / hpt (Reader.h:142-159) run as LSD radix-sort passes, a flag / scan / compact and
 * an atomic histogram over int32 id columns; the sorted lists are copied back so that every host getter answers as after
 * mre_index_create, and the filter / sampler tables are written in place on `device`.  Same bits as the host build.
 * Limits: E, R and the total number of triples below 2^31 (MRE_ERR_INVALID otherwise: use mre_index_create).
 * build_ms (optional): device time of the build proper (first sort to last table written; the copy-back of the lists excluded).
 */
int mre_index_create_device(int device, int64_t E, int64_t R,
                            const int64_t *train_h, const int64_t *train_t, const int64_t *train_r, int64_t n_train,
                            const int64_t *valid_h, const int64_t *valid_t, const int64_t *valid_r, int64_t n_valid,
                            const int64_t *test_h, const int64_t *test_t, const int64_t *test_r, int64_t n_test,
                            mre_index **out, double *build_ms);
/* mre_index_create_from_dir with the build on the GPU: the files are parsed on the host, then as mre_index_create_device */
int mre_index_create_from_dir_device(const char *in_path, int device, mre_index **out, double *build_ms);
void mre_index_destroy(mre_index *ix);
/* upload the filter / sampler tables to `device` (idempotent) */
int mre_index_to_device(mre_index *ix, int device);
int64_t mre_index_total(const mre_index *ix, int which);
/* One column of the device-resident tables, copied to host_out (NULL: only the length is returned; < 0 on error).  which:
 * 0-3 the filter tables over all splits, de-duplicated -- key h R + r / payload t in (h,r,t) order, key t R + r / payload h in
 * (t,r,h) order (what _find bisects, Corrupt.h:166-177); 4-9 the sampler tables over de-duplicated train -- h, r, t and key
 * h R + r in (h,r,t) order, key t R + r / payload h in (t,r,h) order (trainHead / trainTail, Reader.h:107-140); int64 each;
 * 10 the Bernoulli threshold per relation (float32, Base.cpp:113).  For tests and debugging. */
int64_t mre_index_device_column(const mre_index *ix, int which, void *host_out);
int mre_index_get_split(const mre_index *ix, int split, int64_t *h, int64_t *t, int64_t *r);
/* left_mean = tph, right_mean = hpt per relation (Reader.h:142-159) */
int mre_index_get_means(const mre_index *ix, float *tph, float *hpt);
/*
 * Type constraints: importTypeFiles, Reader.h:267-317 (type_constrain.txt: per relation the admissible head ids, then the
 * admissible tail ids).  mre_index_create_from_dir loads the file when it is present; these entry points load it from
 * another path or install lists held in memory (prefix arrays [R+1] + ids; sorted and de-duplicated on the way in, as
 * the reference sorts them and its merge pointer counts a repeated id once, Test.h:89-90).  side 0 = head lists, 1 = tail.
 * A type-constrained rank (Test.h:88-98,153-163) is mre_rank with one candidate group per relation over these lists.
 */
int mre_index_load_type_constrain(mre_index *ix, const char *path);
int mre_index_set_type_constrain(mre_index *ix, const int64_t *head_ptr, const int64_t *head_idx,
                                 const int64_t *tail_ptr, const int64_t *tail_idx);
/* number of ids over all relations of `side`, or -1 when no constraints are loaded */
int64_t mre_index_type_total(const mre_index *ix, int side);
/* ptr: int64 [R+1], idx: int64 [mre_index_type_total] (may be NULL to fetch the prefix only) */
int mre_index_get_type_constrain(const mre_index *ix, int side, int64_t *ptr, int64_t *idx);
/* 1 when (h,r,t) is in train+valid+test: _find, Corrupt.h:166-177 (host) */
int mre_index_find(const mre_index *ix, int64_t h, int64_t t, int64_t r);

/* ----------------------------------------------------------------------------------------- workspace */
int  mre_ctx_create(int device, mre_ctx **out);
void mre_ctx_destroy(mre_ctx *ctx);
int  mre_ctx_sm_count(const mre_ctx *ctx);
/* number of kernels this library has launched through `ctx` since creation (bench.py gpu_launches) */
int64_t mre_ctx_launch_count(const mre_ctx *ctx);

/*
 * Tunables of one workspace (defaults in brackets): "bil_products" [3] = tensor-core products per FP32 product of the
 * DistMult / ComplEx path (3: BF16 hi/lo split, hi*hi + lo*hi + hi*lo; 1: one FP16 product with a wider -- still rigorous --
 * near-tie guard and more exact re-scores; the COUNTS are those of the sequential FP32 scorer either way);
 * "bil_pair" [1] = CTA pairs (cta_group::2) on that path; "transe_ctas_per_sm" [0 = built-in]; "zsl_fp32" [0] = run the ZSL
 * pair contraction on the FP32 CUDA-core kernels instead of 3xTF32 tcgen05; "transe_lpt" [1] = cost-balanced (longest item
 * first) work-item order of the TransE kernel on jobs of a few waves.  Unknown keys fail with MRE_ERR_INVALID.
 */
int mre_ctx_option(mre_ctx *ctx, const char *key, int64_t value);
/* Read-and-reset a device-side statistic (synchronises the device): "bil_rescored" = columns of the DistMult / ComplEx path that
 * fell inside the near-tie guard and were re-scored exactly since the last read (bench.py reports the fraction). */
int mre_ctx_stat(mre_ctx *ctx, const char *key, int64_t *value);

/* ------------------------------------------------------------------------------------------- ranking */
/*
 * One ranking job: Q queries, each ranked against a candidate set, fused score + compare + count.
 * Replaces, per query, Model.predict (TransE.py:88-94 / DistMult.py:70-72 / ComplEx.py:60-61) +
 * getHeadBatch/getTailBatch (Test.h:36-53) + testHead/testTail (Test.h:65-192); with candidate groups
 * it replaces main.evaluate's per-triple candidate loop (main.py:230-250).
 */
typedef struct mre_rank_job {
    /* embedding tables, device, row-major [rows, D], row stride = D floats */
    const float *ent;        /* [E, D]   ent_embeddings / ent_re_embeddings */
    const float *rel;        /* [R, D]   rel_embeddings / rel_re_embeddings */
    const float *ent_im;     /* [E, D]   ComplEx only (ent_im_embeddings) */
    const float *rel_im;     /* [R, D]   ComplEx only */
    int64_t E, R, D;
    int32_t scorer;          /* MRE_TRANSE | MRE_DISTMULT | MRE_COMPLEX | MRE_ROTATE */
    int32_t p_norm;          /* TransE: 1 or 2 (TransE.py:10,59) */
    int32_t normalize;       /* TransE norm_flag (TransE.py:47-50): L2-normalise h, r, t rows first */
    int32_t filter;          /* MRE_FILTER_* */
    /* queries, device int64 [Q]; q_side NULL => every query uses `side` */
    const int64_t *q_h, *q_t, *q_r;
    const uint8_t *q_side;
    int32_t side;
    int32_t n_groups;        /* 0 => one group: every query ranks against all E entities */
    int64_t Q;
    /* candidate groups (n_groups > 0): queries must be ordered by group.
     * group_qptr / group_cptr are HOST int64 [n_groups+1] prefix arrays; cand_idx is a DEVICE int64 array
     * of entity ids, each group's slice sorted ascending without duplicates. */
    const int64_t *group_qptr;
    const int64_t *group_cptr;
    const int64_t *cand_idx;
    /* MRE_FILTER_CSR: device int64 prefix [Q+1] and entity ids, each query's slice SORTED ASCENDING (repeated ids
     * count once); entries outside the query's candidate group are ignored; the true entity is always excluded
     * from the filtered counts. */
    const int64_t *filt_ptr;
    const int64_t *filt_idx;
    /* MRE_FILTER_CSR: filt_ptr[Q], the number of list entries (an upper bound will do; a smaller value drops entries).
     * With it the known-true correction runs flattened over the entries; 0 = unknown: one warp walks each query's list. */
    int64_t filt_nnz;
    /* output, device int32 [4][Q]: raw_lt, raw_eq, filt_lt, filt_eq
     *   raw_lt  = #{j in S_q : s_j <  s_true}          raw_eq  = #{j in S_q : s_j == s_true}
     *   filt_*  = the same over S_q minus known-true entities minus the true entity itself
     * OpenKE's l_s / l_filter_s (Test.h:80-87) are raw_lt / filt_lt. */
    int32_t *counts;
    /* MRE_ROTATE: rel_embedding_range / pi in float32 (RotatE.py:49: phase = r / (rel_embedding_range / pi)); 0 otherwise */
    float rotate_phase_div;
} mre_rank_job;

/* asynchronous on `stream`; `ix` may be NULL unless filter == MRE_FILTER_INDEX */
int mre_rank(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, void *stream);

/*
 * End-to-end variant for callers holding HOST query arrays (the reference's loaders yield numpy arrays,
 * Tester.py:62-68): q_h/q_t/q_r/q_side and counts in `job` are HOST pointers (cand_idx, filt_*, tables
 * stay device).  Copies queries in through pinned staging, ranks, copies counts back, synchronises.
 */
int mre_rank_host(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, void *stream);

/*
 * Materialised 1-vs-all scores for ONE query (Model.predict's float32[E] vector), for drop-in callers that
 * still want the score vector and for tests.  scores_out: device float32 [E].
 */
int mre_predict(mre_ctx *ctx, const mre_rank_job *job, int64_t query, float *scores_out, void *stream);

/*
 * Materialised tensor-core scores of a DistMult / ComplEx job (diagnostic; also serves callers that want many score
 * rows at once): scores_out is device float32 [Q, C] with C = E (all-entity jobs) or the total number of candidate
 * rows (grouped jobs; a query's own group occupies columns [group_cptr[g], group_cptr[g+1]), the rest is untouched).
 * Values are the raw similarities the tcgen05 path produces (predict() = their negation), i.e. BEFORE the exact FP32
 * re-score mre_rank applies to near-ties -- tests use it to bound the tensor-core vs FP32 discrepancy.
 */
int mre_bilinear_scores(mre_ctx *ctx, const mre_rank_job *job, float *scores_out, void *stream);

/*
 * Metric sums from counts (test_link_prediction, Test.h:232-277, without the printf table; the paper's
 * main.py:263-272 and zsl_module.py:707-745 summaries).  counts: device int32 [4][Q]; q_side as in the job.
 * sums_out: device int64 [2][8] per side s: {n, sum_rank, hits@1, hits@3, hits@5, hits@10, rr_fx, 0} for the
 * FILTERED rank, rr_fx = sum floor(2^32 / rank), the reciprocal-rank sum in 32.32 fixed point (an integer: shards of a query
 * set add up exactly, in any order and on any number of GPUs; MRR = rr_fx / 2^32 / n to within 2^-32), and
 * rr_out: device double [2] sum of 1/rank, reduced in a fixed order (deterministic on one device).
 * raw != 0 => use the raw counts instead of the filtered ones.  hist (nullable): device int64 [hist_len],
 * hist[k] += #queries with rank k (k clipped to hist_len-1): the integer form combined across GPUs by allreduce.
 */
int mre_metrics(mre_ctx *ctx, const int32_t *counts, const uint8_t *q_side, int32_t side, int64_t Q,
                int32_t rank_mode, int32_t raw, int64_t *sums_out, double *rr_out,
                int64_t *hist, int64_t hist_len, void *stream);

/* ------------------------------------------------------------------------------ projected translation models */
/*
 * TransH / TransD rank through the TransE kernel over PER-RELATION entity tables: the projection the reference applies to the
 * E rows of every 1-vs-all query (TransH._transfer, TransH.py:66-74: e - (e . w^) w^ with w^ = F.normalize(norm_vector[r]);
 * TransD._transfer, TransD.py:92-109: F.normalize(e + (e . e_p) r_p), dim_e == dim_r) depends on the relation only, so
 * mre_relation_project writes out[slot * E + e, :] once for the slot-th relation of `rels` (renormalize != 0 applies _calc's
 * F.normalize again, TransH.py:51-55).  A job then ranks with one candidate group per relation over this table.
 * ent [E, D]; ent_aux [E, D] = ent_transfer (TransD) or NULL; rel_aux [R, D] = norm_vector (TransH) / rel_transfer (TransD);
 * rels: device int64 [n_rel]; out: device float32 [n_rel * E, D]; D <= 512.
 */
#define MRE_PROJECT_TRANSH 0
#define MRE_PROJECT_TRANSD 1
int mre_relation_project(mre_ctx *ctx, int32_t kind, const float *ent, const float *ent_aux, const float *rel_aux,
                         const int64_t *rels, int64_t n_rel, int64_t E, int64_t D, int32_t renormalize, float *out, void *stream);

/* ------------------------------------------------------------------------------------------ sampling */
/*
 * One training batch of B positives + B*neg Bernoulli-corrupted negatives, layout [B pos | neg blocks of B]
 * (row b's k-th negative at b + (k+1)*B): sampling / getBatch, Base.cpp:78-197, with corrupt_head /
 * corrupt_tail (Corrupt.h:7-83) and the tph/hpt rule (Base.cpp:101-122).  The reference's per-thread LCG
 * (Random.h:11-29) is replaced by counter-based Philox4x32-10:
 *   key = (seed_lo, seed_hi); ctr = (row b, slot, step_lo, (step_hi & 0xffff) | stream_id << 16)
 *   slot 0: positive index = (x1:x0) % trainTotal;  slot k+1: keep_head = float(x0 % 1000) < prob_r, draw word (x2:x1)
 * mode: 0 normal (Bernoulli), -1 always corrupt head ("head_batch"), 1 always corrupt tail (Base.cpp:125-137).
 * Outputs are device arrays of length B*(1+neg): int64 h, t, r and float32 y (+1 / -1).
 */
int mre_sample(mre_ctx *ctx, const mre_index *ix, uint64_t seed, uint64_t step, uint32_t stream_id,
               int64_t B, int64_t neg, int32_t mode, int32_t bern,
               int64_t *h, int64_t *t, int64_t *r, float *y, void *stream);
/* same with HOST output arrays (what TrainDataLoader hands out); synchronous */
int mre_sample_host(mre_ctx *ctx, const mre_index *ix, uint64_t seed, uint64_t step, uint32_t stream_id,
                    int64_t B, int64_t neg, int32_t mode, int32_t bern,
                    int64_t *h, int64_t *t, int64_t *r, float *y, void *stream);

/*
 * The paper's subgraph sampler: NegativeSampling.neg_sample_fn + __normal_batch + __corrupt_head / __corrupt_tail
 * (module/NegativeSampling.py:114-140, 321-375).  Edges and node_list hold LOCAL node ids of one sampled subgraph;
 * local_to_global[local] is the entity id the train triples use.  Per edge, neg_ent decision draws split the negatives
 * into nh head-corruptions (slots 1..nh) and neg_ent - nh tail-corruptions (slots nh+1..neg_ent), each with probability
 * 0.5, or hpt/(hpt+tph) of the edge's relation when bern != 0 (the reference's bern branch reads tables it never fills,
 * :95-99,324).  A replacement is drawn uniformly from node_list; when filter != 0 (filter_flag) it is redrawn while its
 * global id completes a TRAIN triple with the kept entity and the relation (h_of_tr / t_of_hr, :360,373).
 *   key = (seed_lo, seed_hi); ctr = (edge b, slot | attempt << 16, step_lo, (step_hi & 0xffff) | stream_id << 16)
 *   attempt 0, slot j (1..neg_ent): decision j of the edge = (x0 >> 8) * 2^-24 < prob
 *   attempt a >= 1, slot k:         position (x1:x0) % n_nodes in node_list; at most 64 attempts, then the first
 *                                   admissible node scanning on from the last position; none => the edge's own id is kept
 * Outputs: device int32 arrays of length n_edges * (1 + neg_ent) in the reference's transposed layout (slot k of edge b at
 * k * n_edges + b; slot 0 = the edge itself): expand_edge_index[0], expand_edge_index[1], expand_edge_type (:138-140).
 */
int mre_sample_subgraph(mre_ctx *ctx, const mre_index *ix, uint64_t seed, uint64_t step, uint32_t stream_id,
                        const int64_t *edge_h, const int64_t *edge_t, const int64_t *edge_r, int64_t n_edges,
                        const int64_t *node_list, int64_t n_nodes, const int64_t *local_to_global, int64_t n_local,
                        int64_t neg_ent, int32_t bern, int32_t filter,
                        int32_t *out_h, int32_t *out_t, int32_t *out_r, void *stream);

/*
 * Type-constrained negative tails: corrupt(h, r), Corrupt.h:179-195 -- a tail drawn uniformly from the relation's tail-type
 * list (type_constrain.txt, importTypeFiles, Reader.h:267-317), redrawn while (h, r, tail) is a known triple of ANY split
 * (_find, Corrupt.h:166-177); after 1000 failed draws (or for an empty list) the exact-uniform corrupt_head over the train
 * set (Corrupt.h:7-44).  The reference draws with libc rand(); here
 *   key = (seed_lo, seed_hi); ctr = (pair i lo, attempt | (i hi) << 16, step_lo, (step_hi & 0xffff) | stream_id << 16)
 *   attempt 0..999: list position (x1:x0) % length;  attempt 1000: the draw word of the corrupt_head fallback.
 * The index's type lists are sorted and de-duplicated (the reference keeps repeated ids of the file, which then weigh
 * more).  h, r: device int64 [n]; t_out: device int64 [n].  The index must hold type constraints and be on ctx's device.
 */
int mre_corrupt_typed(mre_ctx *ctx, const mre_index *ix, uint64_t seed, uint64_t step, uint32_t stream_id,
                      const int64_t *h, const int64_t *r, int64_t n, int64_t *t_out, void *stream);

/* ------------------------------------------------------------------------------------------ ZSL scorer */
/*
 * The ZSL candidate scorer of ZSLmodule.eval (module/zsl_module.py:662-745): Extractor.forward (zsl_module.py:46-106, with
 * SupportEncoder, module/submodule.py:240-258) on every (head, candidate) pair of every test triple, cosine similarity
 * against the relation's generated vectors averaged over them (sklearn cosine_similarity(...).mean(axis=1), :699-701), and
 * the rank of the true candidate (index 0 of its list, :705-706).  All pointers are device pointers.  Tensor names follow the
 * reference's state_dict: symbol_emb.weight [n_symbols + 1, D] (last row = padding, zeros), gcn_w / fc1 / fc2 .weight [D/2, D]
 * and .bias [D/2], reshape_layer.weight [D, 2D] / .bias [D], support_encoder.proj1.weight [2D, D] / .bias [2D],
 * support_encoder.proj2.weight [D, 2D] / .bias [D], support_encoder.layer_norm.weight / .bias [D] (eps 1e-5).
 */
typedef struct mre_zsl_model {
    int64_t D;                 /* emb_dim, args.py:41 (200); a multiple of 8, at most 256 */
    const float *symbol_emb;
    const float *gcn_w, *gcn_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
    const float *reshape_w, *reshape_b;
    const float *proj1_w, *proj1_b, *proj2_w, *proj2_b;
    const float *ln_g, *ln_b;
    float ln_eps;
} mre_zsl_model;
/*
 * Per-entity halves of the pair encoder: reshape_layer([N_h | tanh fc1(h) | tanh fc2(c) | N_c]) = A[h] + B[c], where
 * N_e = tanh(sum_j gcn_w(symbol_emb[conn[e][j]]) / deg[e]) is neighbor_encoder (zsl_module.py:46-59; conn holds the
 * NEIGHBOUR symbol ids, connections[e, :, 1] of build_connection :239-268, padded with the pad id; deg = e1_degrees).
 * ent_symbol[e] = symbol id of entity e.  A, B: float32 [n_ent, D].
 */
int mre_zsl_entity_features(mre_ctx *ctx, const mre_zsl_model *model, const int64_t *ent_symbol, const int64_t *conn,
                            const float *deg, int64_t n_ent, int32_t max_neighbor, float *A, float *B, void *stream);
/*
 * Scores + ranks of T test triples.  Triple t has head entity q_head[t], relation slot q_rel[t] (row of rel_vecs
 * [n_rel, n_vec, D], the generator's test_sample vectors per relation, :657-660) and candidate entities
 * cand_idx[cand_ptr[t] .. cand_ptr[t+1]) with the TRUE tail first (test_candidates.json, :665-682); P = cand_ptr[T].
 * scores (nullable): float32 [P].  counts: int32 [4][T] in mre_metrics' layout -- rows 0 and 2 = #candidates scoring
 * HIGHER than the true one, rows 1 and 3 = #exact ties: rank = counts[0] + 1 (ties resolved for the true candidate) up to
 * counts[0] + counts[1] + 1 (MRE_RANK_PESSIMISTIC); the reference's argsort leaves exact ties unpinned.
 * A, B: the [n_ent, D] halves mre_zsl_entity_features wrote (n_ent rows: the hidden layer of the support encoder is
 * split per entity the same way, W1 A + b1 and W1 B, before the per-pair tensor-core contraction with W2).
 * Arithmetic: FP32 everywhere except that contraction, which runs as hi*hi + hi*lo + lo*hi of TF32-split FP32 operands
 * with FP32 accumulation (relative error 2^-21 per product; scores within 1e-7 of an all-FP32 evaluation, 5e-8 of the
 * reference's on tests/golden/golden_zsl.npz).  D up to 216 takes that path, wider models all-FP32 CUDA-core kernels.
 */
int mre_zsl_rank(mre_ctx *ctx, const mre_zsl_model *model, const float *A, const float *B, int64_t n_ent,
                 const int64_t *q_head, const int64_t *q_rel, const int64_t *cand_ptr, const int64_t *cand_idx, int64_t T,
                 int64_t P, const float *rel_vecs, int64_t n_rel, int32_t n_vec, float *scores, int32_t *counts, void *stream);

/* ------------------------------------------------------------------------------------------ training */
/*
 * Fused TransE margin-loss step, forward + backward, on one sampled batch (n = B*(1+neg) triples):
 *   score = || h^ + r^ - t^ ||_p  (TransE.py:46-74), p = score[:B], n[b,k] = score[B + k*B + b]
 *   (strategy/NegativeSampling.py:13-21), loss = mean max(p - n, -margin) + margin (MarginLoss.py:24-28).
 * Accumulates dLoss/d(ent) into grad_ent [E,D] and dLoss/d(rel) into grad_rel [R,D] (caller zeroes them),
 * writes the scalar loss to loss_out[0] and, if scores_out != NULL, the n scores.  All pointers device.
 */
int mre_transe_margin_step(mre_ctx *ctx, const float *ent, const float *rel, int64_t E, int64_t R, int64_t D,
                           const int64_t *h, const int64_t *t, const int64_t *r, int64_t B, int64_t neg,
                           float margin, int32_t p_norm, int32_t normalize,
                           float *grad_ent, float *grad_rel, float *loss_out, float *scores_out, void *stream);
/*
 * The same step in two halves, for callers that put their own loss between them (Model.forward, TransE.py:62-74,
 * DistMult.py:46-57, ComplEx.py:29-40; head_batch / tail_batch callers expand their index arrays to length n first):
 * mre_score_triples writes the raw model score of n explicit triples (TransE: the distance; DistMult / ComplEx: the
 * similarity, which predict() negates); mre_transe_backward scatters dLoss/d(ent), dLoss/d(rel) given dLoss/dscore.
 */
int mre_score_triples(mre_ctx *ctx, int32_t scorer, const float *ent, const float *ent_im, const float *rel, const float *rel_im,
                      int64_t D, const int64_t *h, const int64_t *t, const int64_t *r, int64_t n, int32_t p_norm,
                      int32_t normalize, float *score_out, void *stream);
int mre_transe_backward(mre_ctx *ctx, const float *ent, const float *rel, int64_t D, const int64_t *h, const int64_t *t,
                        const int64_t *r, int64_t n, int32_t p_norm, int32_t normalize, const float *score, const float *dscore,
                        float *grad_ent, float *grad_rel, void *stream);
/*
 * dLoss/d(tables) of the similarity models given dLoss/dscore of n explicit triples: autograd through DistMult._calc
 * (DistMult.py:34-44; s = sum h*r*t) and ComplEx._calc (ComplEx.py:20-27), accumulated with float atomics into dense gradient
 * tables of the tables' shapes (caller zeroes them; g_ent_im / g_rel_im only for ComplEx).
 */
int mre_bilinear_backward(mre_ctx *ctx, int32_t scorer, const float *ent, const float *ent_im, const float *rel, const float *rel_im,
                          int64_t D, const int64_t *h, const int64_t *t, const int64_t *r, int64_t n, const float *dscore,
                          float *g_ent, float *g_ent_im, float *g_rel, float *g_rel_im, void *stream);
/*
 * The losses of the negative-sampling strategy on its score layout (strategy/NegativeSampling.py:13-21: p_b = score[b],
 * n_bk = score[B + k*B + b]), value and gradient in one launch: kind = MRE_LOSS_* ; adv != 0 applies the detached
 * self-adversarial weights softmax_k(-T n_bk) (margin, MarginLoss.py:21-26) / softmax_k(+T n_bk) (sigmoid, softplus;
 * SigmoidLoss.py:19-24, SoftplusLoss.py:19-24) with T = adv_temperature.  loss_out: device float32 [1];
 * dscore (nullable): device float32 [B*(1+neg)] = dLoss/dscore.
 */
int mre_ns_loss(mre_ctx *ctx, int32_t kind, const float *score, int64_t B, int64_t neg, float margin, int32_t adv,
                float adv_temperature, float *loss_out, float *dscore, void *stream);
/*
 * One fused forward + loss + backward step of the strategy for ANY scorer and loss (Trainer.train_one_step, Trainer.py:43-54,
 * without the optimizer): mre_score_triples -> mre_ns_loss -> mre_transe_backward / mre_bilinear_backward, three launches.
 * mre_transe_margin_step is the (MRE_TRANSE, MRE_LOSS_MARGIN, adv = 0) case.
 */
int mre_ns_train_step(mre_ctx *ctx, int32_t scorer, const float *ent, const float *ent_im, const float *rel, const float *rel_im,
                      int64_t D, const int64_t *h, const int64_t *t, const int64_t *r, int64_t B, int64_t neg, int32_t loss_kind,
                      float margin, int32_t adv, float adv_temperature, int32_t p_norm, int32_t normalize,
                      float *g_ent, float *g_ent_im, float *g_rel, float *g_rel_im, float *loss_out, float *scores_out, void *stream);
/* w -= lr * g (torch.optim.SGD as Trainer.py:73-78 configures it), then g = 0; device arrays of n floats */
int mre_sgd_update(mre_ctx *ctx, float *w, float *g, int64_t n, float lr, void *stream);

/* ------------------------------------------------------------------------------- multi-GPU (one process per GPU) */
/*
 * The reference is one process on one GPU (Trainer.py:43-78; Test.h:232-277 sums the metrics in one address space).  The two
 * exchanges of the data-parallel path -- the gradient sum of a training step and the integer metric sums of a sharded
 * evaluation -- run here over NVLink PEER MEMORY, without any collective library:
 *   mre_peer_group_create   cudaMalloc of this rank's region [weights n | gradients n | exchange area | flags] + its CUDA IPC
 *                           handle (MRE_PEER_HANDLE_BYTES bytes).  The caller ships the handles between the processes by any
 *                           means (dist.py: torch.distributed.all_gather_object) and hands all of them, in rank order, to
 *   mre_peer_group_connect  which maps every other rank's region (cudaIpcOpenMemHandle).
 *   mre_peer_weights / mre_peer_grads   this rank's flat parameter and gradient buffers inside the region (n floats each):
 *                           the embedding tables and gradient tables of the training step live there back to back.
 *   mre_dp_sgd_step         ONE kernel per rank and step = reduce-scatter + SGD + all-gather: rank r sums slice r of every rank's
 *                           gradient buffer (P2P loads, rank order), applies w -= lr * sum, stores the new slice into every rank's
 *                           weight buffer (P2P stores), and zeroes its own gradient buffer once every rank is done with it.  Pass
 *                           lr / world for the mean of the ranks' mean losses.  All ranks must call it once per step, in step;
 *                           weights stay bit-identical on all ranks.  max_blocks: 0 = one block per SM (tests pass a small number).
 *   mre_peer_allreduce_i64  in-place sum over the ranks of a device int64 vector of at most 64 words (mre_metrics' sums_out).
 *   mre_peer_group_error    MRE_ERR_CUDA if an exchange ever timed out (a rank did not arrive within ~20 s; nothing hangs).
 * Flags carry the exchange number, which only grows: no resets, no host synchronisation between steps.
 */
typedef struct mre_peer_group mre_peer_group;
#define MRE_PEER_HANDLE_BYTES 64
#define MRE_PEER_MAX_RANKS 16
int mre_peer_group_create(mre_ctx *ctx, int32_t rank, int32_t world, int64_t n_floats, mre_peer_group **out,
                          unsigned char *handle_out /* [MRE_PEER_HANDLE_BYTES] */);
int mre_peer_group_connect(mre_peer_group *g, const unsigned char *handles /* [world][MRE_PEER_HANDLE_BYTES], rank order */);
/* the same for "ranks" that are all groups of THIS process on one device (tests): all[p] = rank p's group */
int mre_peer_group_connect_local(mre_peer_group *g, mre_peer_group *const *all);
void mre_peer_group_destroy(mre_peer_group *g);
float *mre_peer_weights(mre_peer_group *g);
float *mre_peer_grads(mre_peer_group *g);
int mre_dp_sgd_step(mre_ctx *ctx, mre_peer_group *g, float lr, int32_t max_blocks, void *stream);
int mre_peer_allreduce_i64(mre_ctx *ctx, mre_peer_group *g, int64_t *vec, int32_t count, void *stream);
int mre_peer_group_error(mre_peer_group *g);

/* ------------------------------------------------------------------------------------------- probes */
/* FP32 add-rate microbenchmark (the better of a scalar FADD and a packed FADD2 stream): lane-ops per second in *lane_ops_per_s (the TransE roofline
 * denominator) and the SM clock-independent instruction count used. Synchronous. */
int mre_probe_fp32_peak(mre_ctx *ctx, double *lane_ops_per_s);
/* MUFU.SQRT (sqrt.approx.ftz) rate microbenchmark: square roots per second in *ops_per_s -- the roofline denominator of the RotatE tile
 * kernel (RotatE.py:74-76: one complex modulus per query x entity x dimension). Synchronous. */
int mre_probe_mufu_peak(mre_ctx *ctx, double *ops_per_s);
/* tcgen05 dense MMA microbenchmarks, flops per second: kind::f16 with BF16 operands (what the DistMult / ComplEx kernel
 * issues -- its roofline denominator when MEASURED_PEAKS.json is absent) and kind::tf32 (for reference) */
int mre_probe_bf16_peak(mre_ctx *ctx, double *flops_per_s);
int mre_probe_tf32_peak(mre_ctx *ctx, double *flops_per_s);
/* Per-launch device timing of the dominant kernel (the fused score+rank kernel of mre_rank / mre_rank_host, or the
 * train-step kernel): while enabled, every such launch is bracketed by a CUDA event pair recorded on the launching
 * stream.  mre_ctx_timing_read synchronises, returns the summed duration and the number of launches since the
 * last read, and resets both (bench.py's roofline leg). */
int mre_ctx_timing(mre_ctx *ctx, int32_t enable);
int mre_ctx_timing_read(mre_ctx *ctx, double *total_ms, int64_t *n_launches);

#ifdef __cplusplus
}
#endif
#endif /* MRE_B200_H */
