// TransE L1/L2 fused score + rank for sm_100a.
//
// Reference path replaced (paths relative to /root/reference):
//   OpenKE/openke/module/model/TransE.py:46-60   _calc: optional F.normalize, h + (r - t) | (h + r) - t, ||.||_p
//   OpenKE/openke/module/model/TransE.py:88-94   predict -> host float32[E] per query
//   module/NegativeSampling.py:142-157,294-305   the paper's scorer/evaluate (p = 1, no normalisation)
//   OpenKE/openke/base/Test.h:65-192             testHead / testTail: E-long compare loop + _find per hit
//   main.py:245-250                              candidate-list rank with ties//2
//
// Design.  The query x entity score matrix is never written.  A pre-pass turns every query into one FP32
// vector  v_q  such that |v_q[d] - e_j[d]| is bit-identical to the reference's element (tail: v = h + r;
// head: v = -(r - t), because e + (r - t) == -(v - e) exactly), and computes s_true with the SAME sequential-d
// accumulation the tile kernel uses, so `s_j < s_true` is decided on identical bits for every j.
// The main kernel is persistent: each CTA walks 128-query x 128-entity work items; 128-row x 128-byte boxes of both
// operands stream into a 3-stage shared-memory ring through TMA (cp.async.bulk.tensor, 128-byte swizzle, mbarrier
// complete_tx; candidate lists are gathered into a dense table by a pre-pass), issued by whichever warp is last
// to release a stage; the eight warps hold an 8x8 register micro-tile per thread and read the ring with
// conflict-free 128-bit LDS.  The
// epilogue compares the 64 accumulators against the per-query thresholds, masks the known-true slots routed to this
// tile by tile_filter.cu (raw and filtered counts come from the SAME registers), reduces the counts with warp
// shuffles and adds them to the per-query counters.  The kernel is bound by the FP32 pipe: 2 lane-ops per (q, e, d).
#include <math.h>

#include <vector>

#include "common.h"
#include "rank_common.cuh"
#include "rank_host.h"
#include "tma_host.h"

namespace mre {

constexpr int CHUNK = 32;            // floats of D per pipeline stage (128 B per row)
constexpr int STAGES = 3;
constexpr int CONSUMER_WARPS = 8;
constexpr int RANK_THREADS = CONSUMER_WARPS * 32;
constexpr uint32_t STAGE_BYTES = (TILE_Q + TILE_E) * CHUNK * 4;  // two 16 KiB TMA boxes
constexpr size_t RANK_SMEM = 1024 + (size_t)STAGES * STAGE_BYTES + STAGES * sizeof(uint64_t) + 4 * sizeof(int) + CONSUMER_WARPS * 32 * sizeof(unsigned long long) + 64;

// ------------------------------------------------------------------------------------------ scalar scorer
// The one definition of a TransE accumulator: sequential over d, acc = acc + |v - e| (p = 1) or fma(u, u, acc) (p = 2).
// Query vectors are stored PAIR-SWAPPED (element d at position d ^ 1): a 128-bit load then puts v[d] in a register
// of the opposite even/odd bank to e[d], so the tile kernel's `v[d] - e[d]` never reads two same-bank registers.
template <int P>
__device__ __forceinline__ float transe_acc(const float *__restrict__ v, const float *__restrict__ e, int64_t D) {
    float acc = 0.f;
    const int n = (int)D;
#pragma unroll 8
    for (int d = 0; d < n; d += 4) {   // unrolled: the row fetches of 8 steps are in flight together (latency-bound callers)
        float4 a = __ldg(reinterpret_cast<const float4 *>(v + d));
        float4 b = __ldg(reinterpret_cast<const float4 *>(e + d));
        float u0 = a.y - b.x, u1 = a.x - b.y, u2 = a.w - b.z, u3 = a.z - b.w;
        if (P == 1) {
            acc = acc + fabsf(u0); acc = acc + fabsf(u1); acc = acc + fabsf(u2); acc = acc + fabsf(u3);
        } else {
            acc = fmaf(u0, u0, acc); acc = fmaf(u1, u1, acc); acc = fmaf(u2, u2, acc); acc = fmaf(u3, u3, acc);
        }
    }
    return acc;
}

// ------------------------------------------------------------------------------------------ pre-pass kernels
// F.normalize(x, 2, -1) = x / max(||x||_2, 1e-12) (TransE.py:47-50), written into a table whose rows are padded
// with zeros to Dp (a multiple of 4).  One thread per row, sequential fma: bit-identical to oracle/kge_oracle.c.
__global__ void normalize_rows_kernel(const float *__restrict__ x, int64_t n, int64_t D, int64_t Dp, int normalize,
                                      float *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *row = x + i * D;
    float nrm = 1.f;
    if (normalize) {
        float ss = 0.f;
        for (int64_t d = 0; d < D; d++) ss = fmaf(row[d], row[d], ss);
        nrm = __fsqrt_rn(ss);
        if (nrm < 1e-12f) nrm = 1e-12f;
    }
    float *o = out + i * Dp;
    for (int64_t d = 0; d < D; d++) o[d] = normalize ? __fdiv_rn(row[d], nrm) : row[d];
    for (int64_t d = D; d < Dp; d++) o[d] = 0.f;
}

// v_q = h + r (tail query) or -(r - t) (head query), stored pair-swapped; grid-stride over Q * D elements
__global__ void transe_qvec_kernel(const float *__restrict__ ent, const float *__restrict__ rel, int64_t D,
                                   const int64_t *__restrict__ q_h, const int64_t *__restrict__ q_t,
                                   const int64_t *__restrict__ q_r, const uint8_t *__restrict__ q_side, int side,
                                   int64_t Q, float *__restrict__ qvec) {
    int64_t total = Q * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t q = i / D, d = i - q * D;
        int s = q_side ? (int)q_side[q] : side;
        float rv = rel[q_r[q] * D + d];
        float v;
        if (s) v = ent[q_h[q] * D + d] + rv;
        else v = -(rv - ent[q_t[q] * D + d]);
        qvec[i ^ 1] = v;
    }
}

// thresholds from the true entity's accumulator.  p = 1: score = acc.  p = 2: score = sqrt(acc) and the reference
// compares the square roots, so lo = min{x : sqrt(x) >= s_true}, hi = min{x : sqrt(x) > s_true} (sqrt is monotone),
// which lets the tile kernel compare raw accumulators and still agree with sqrtf(acc_j) < sqrtf(acc_true) exactly.
template <int P>
__global__ void transe_threshold_kernel(const float *__restrict__ ent, int64_t D, const int64_t *__restrict__ q_h,
                                        const int64_t *__restrict__ q_t, const uint8_t *__restrict__ q_side, int side,
                                        int64_t Q, const float *__restrict__ qvec, float2 *__restrict__ thr) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    int s = q_side ? (int)q_side[q] : side;
    int64_t truth = s ? q_t[q] : q_h[q];
    float acc = transe_acc<P>(qvec + q * D, ent + truth * D, D);
    float lo = acc, hi = acc;
    if (acc >= 0.f && acc < INFINITY) {
        if (P == 1) {
            hi = __int_as_float(__float_as_int(acc) + 1);
        } else {
            float st = __fsqrt_rn(acc);
            float x = acc;
            while (x > 0.f) {
                float y = __int_as_float(__float_as_int(x) - 1);
                if (__fsqrt_rn(y) >= st) x = y; else break;
            }
            lo = x;
            x = acc;
            for (;;) {
                float y = __int_as_float(__float_as_int(x) + 1);
                if (y < INFINITY && __fsqrt_rn(y) <= st) x = y; else { hi = y; break; }
            }
        }
    }
    thr[q] = make_float2(lo, hi);
}

// Model.predict for one query: the materialised float32[E] score vector (tests / drop-in callers only)
template <int P>
__global__ void transe_predict_kernel(const float *__restrict__ ent, int64_t E, int64_t D, const float *__restrict__ qv,
                                      float *__restrict__ out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= E) return;
    float acc = transe_acc<P>(qv, ent + j * D, D);
    out[j] = P == 1 ? acc : __fsqrt_rn(acc);
}

// ------------------------------------------------------------------------------------------ main kernel
template <int P>
__device__ __forceinline__ float upd(float acc, float q, float e) {
    float u = q - e;
    return P == 1 ? acc + fabsf(u) : fmaf(u, u, acc);
}

// One pipeline chunk = CHUNK floats (128 B) of every row of one work item's two operands = two TMA boxes of
// 128 rows x 128 B, written into shared memory with the hardware 128-byte swizzle.  Issued by ONE thread.
__device__ __forceinline__ void issue_chunk(const RankParams &p, const CUtensorMap *tm_q, const CUtensorMap *tm_e, int64_t flat,
                                            int n_chunks, uint32_t ring_u32, uint32_t full0) {
    const int64_t item = blockIdx.x + (flat / n_chunks) * (int64_t)gridDim.x;
    if (item >= p.total_items) return;
    const int c = (int)(flat % n_chunks);
    int g, qt, et;
    decode_item(p, item, g, qt, et);
    const GroupDesc &gd = p.groups[g];
    const int qrow = (int)(gd.q0 + (int64_t)qt * TILE_Q);
    const int erow = (int)(gd.c0 + (int64_t)et * TILE_E);
    const int stage = (int)(flat % STAGES);
    const uint32_t full = full0 + 8 * stage;
    const uint32_t sq = ring_u32 + (uint32_t)stage * STAGE_BYTES;
    mbar_arrive_expect_tx(full, STAGE_BYTES);
    tma_load_2d(sq, tm_q, c * CHUNK, qrow, full);
    tma_load_2d(sq + TILE_Q * CHUNK * 4, tm_e, c * CHUNK, erow, full);
}

template <int P, bool NEED_EQ>
__global__ void __launch_bounds__(RANK_THREADS, 2)
transe_rank_kernel(const RankParams p, const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_e) {
    extern __shared__ unsigned char smem_raw[];
    // the 128-byte swizzle pattern is a function of the shared-memory address: tiles must start 1024-byte aligned
    const uint32_t ring_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *ring = smem_raw + (ring_u32 - smem_u32(smem_raw));
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring + (size_t)STAGES * STAGE_BYTES);
    int *done = reinterpret_cast<int *>(bars + STAGES);  // per-stage count of warps that finished reading the stage
    // per-warp known-true masks: bit (i * 8 + j) of lane l's word marks accumulator (i, j) of that thread
    unsigned long long *wmask = reinterpret_cast<unsigned long long *>(done + 4) + (threadIdx.x >> 5) * 32;
    const uint32_t full0 = smem_u32(bars);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_chunks = (int)((p.D + CHUNK - 1) / CHUNK);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_e);
        for (int s = 0; s < STAGES; s++) {
            mbar_init(full0 + 8 * s, 1);
            done[s] = 0;
        }
        fence_barrier_init();
        fence_proxy_async();
        for (int s = 0; s < STAGES; s++) issue_chunk(p, &tm_q, &tm_e, s, n_chunks, ring_u32, full0);
    }
    __syncthreads();

    // entity rows te + 16 j, query rows tq + 16 i.  Row r keeps its logical 16-byte chunk k at (k ^ (r & 7)), and
    // (te + 16 j) & 7 == te & 7: eight consecutive rows read eight distinct bank groups => conflict-free LDS.128.
    const int te = threadIdx.x & 15, tq = threadIdx.x >> 4;
    const int xe = te & 7, xq = tq & 7;
    int64_t it = 0;
    for (int64_t item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        int g, qt, et;
        decode_item(p, item, g, qt, et);
        const GroupDesc gd = p.groups[g];
        const int64_t qbase = gd.q0 + (int64_t)qt * TILE_Q;
        const int nq = (int)min((int64_t)TILE_Q, gd.q0 + gd.nq - qbase);
        const int ne = (int)min((int64_t)TILE_E, gd.nc - (int64_t)et * TILE_E);
        // this item's known-true pairs; the first 32 are fetched now so the epilogue does not wait on them
        const uint32_t pf0 = __ldg(p.tf_ptr + item), pf1 = __ldg(p.tf_ptr + item + 1);
        const uint32_t pair0 = pf0 + lane < pf1 ? __ldg(p.tf_pairs + pf0 + lane) : 0xffffffffu;

        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

        for (int c = 0; c < n_chunks; c++, it++) {
            const int stage = (int)(it % STAGES);
            mbar_wait(full0 + 8 * stage, (uint32_t)((it / STAGES) & 1));
            const unsigned char *sQ = ring + (size_t)stage * STAGE_BYTES + tq * (CHUNK * 4);
            const unsigned char *sE = ring + (size_t)stage * STAGE_BYTES + (TILE_Q + te) * (CHUNK * 4);
            const int nk4 = (int)min((int64_t)CHUNK, p.D - (int64_t)c * CHUNK) >> 2;
#pragma unroll 2
            for (int k4 = 0; k4 < nk4; k4++) {
                const int oe = (k4 ^ xe) << 4, oq = (k4 ^ xq) << 4;
                float4 ev[8];
#pragma unroll
                for (int j = 0; j < 8; j++) ev[j] = *reinterpret_cast<const float4 *>(sE + j * (16 * CHUNK * 4) + oe);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float4 qv = *reinterpret_cast<const float4 *>(sQ + i * (16 * CHUNK * 4) + oq);
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        float a = acc[i][j];
                        a = upd<P>(a, qv.y, ev[j].x);  // pair-swapped query layout: element d sits at d ^ 1
                        a = upd<P>(a, qv.x, ev[j].y);
                        a = upd<P>(a, qv.w, ev[j].z);
                        a = upd<P>(a, qv.z, ev[j].w);
                        acc[i][j] = a;
                    }
                }
            }
            // Release the stage.  The LAST warp to finish reading it refills it with the chunk STAGES ahead: no
            // dedicated producer warp, no empty-barrier spinning, and warps may drift up to STAGES-1 chunks apart.
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                const int old = atomicAdd(&done[stage], 1);
                if ((old & (CONSUMER_WARPS - 1)) == CONSUMER_WARPS - 1) {
                    __threadfence_block();
                    fence_proxy_async();
                    issue_chunk(p, &tm_q, &tm_e, it + STAGES, n_chunks, ring_u32, full0);
                }
            }
        }

        // ---- known-true mask of this warp's 32 threads: every pair routed to this item by tile_filter.cu
        unsigned long long known = 0ull;
        if (pf1 > pf0) {
            wmask[lane] = 0ull;
            __syncwarp();
            for (uint32_t k = pf0 + lane; k < pf1; k += 32) {
                const uint32_t pr = k < pf0 + 32 ? pair0 : __ldg(p.tf_pairs + k);
                const int row = (int)(pr >> 16), col = (int)(pr & 0xffffu);
                const int otq = row & 15;                    // owner thread: tq = row % 16, te = col % 16
                if ((otq >> 1) == warp)
                    atomicOr(&wmask[((otq & 1) << 4) | (col & 15)], 1ull << (((row >> 4) << 3) | (col >> 4)));
            }
            __syncwarp();
            known = wmask[lane];
        }
        // ---- epilogue: compare against the per-query thresholds, count raw and known hits, reduce over the 16 lanes
        // sharing a query row (four 8-bit counters packed in one word: each is at most 128 after the reduction)
        const bool any_known = __any_sync(0xffffffffu, known != 0ull);   // warp-uniform: most warps of most tiles hold none
        bool ev_ok[8];
#pragma unroll
        for (int j = 0; j < 8; j++) ev_ok[j] = (te + 16 * j) < ne;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int ql = tq + 16 * i;
            const bool q_ok = ql < nq;
            float2 th = make_float2(-INFINITY, -INFINITY);
            if (q_ok) th = __ldg(p.thr + qbase + ql);
            uint32_t packed = 0;
            if (any_known) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float s = acc[i][j];
                    const bool lt = ev_ok[j] && s < th.x;
                    const bool eq = NEED_EQ && ev_ok[j] && !lt && s < th.y;
                    const bool kn = (known >> (i * 8 + j)) & 1ull;
                    packed += (lt ? 1u : 0u) + (eq ? 0x100u : 0u) + ((lt && kn) ? 0x10000u : 0u) + ((eq && kn) ? 0x1000000u : 0u);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float s = acc[i][j];
                    const bool lt = ev_ok[j] && s < th.x;
                    const bool eq = NEED_EQ && ev_ok[j] && !lt && s < th.y;
                    packed += (lt ? 1u : 0u) + (eq ? 0x100u : 0u);
                }
            }
#pragma unroll
            for (int m = 1; m < 16; m <<= 1) packed += __shfl_xor_sync(0xffffffffu, packed, m);
            if (q_ok && te == 0 && packed) {
                const int64_t q = qbase + ql;
                const int n_lt = packed & 0xff, n_eq = (packed >> 8) & 0xff, k_lt = (packed >> 16) & 0xff, k_eq = packed >> 24;
                if (n_lt) atomicAdd(p.counts + q, n_lt);
                if (n_eq) atomicAdd(p.counts + p.Q + q, n_eq);
                if (n_lt - k_lt) atomicAdd(p.counts + 2 * p.Q + q, n_lt - k_lt);
                if (n_eq - k_eq) atomicAdd(p.counts + 3 * p.Q + q, n_eq - k_eq);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ host side
static int build_groups(mre_ctx *ctx, const mre_rank_job *job, int tile_q, int tile_e, std::vector<GroupDesc> &groups,
                        int64_t *total_items) {
    groups.clear();
    int64_t items = 0;
    if (job->n_groups <= 0) {
        GroupDesc g{};
        g.q0 = 0; g.nq = job->Q; g.c0 = 0; g.nc = job->E; g.item0 = 0;
        g.n_qt = (int32_t)((job->Q + tile_q - 1) / tile_q);
        g.n_et = (int32_t)((job->E + tile_e - 1) / tile_e);
        items = (int64_t)g.n_qt * g.n_et;
        groups.push_back(g);
    } else {
        MRE_CHECK_ARG(job->group_qptr && job->group_cptr && job->cand_idx, "candidate groups need group_qptr, group_cptr, cand_idx");
        MRE_CHECK_ARG(job->group_qptr[0] == 0 && job->group_qptr[job->n_groups] == job->Q, "group_qptr must span [0, Q]");
        for (int i = 0; i < job->n_groups; i++) {
            GroupDesc g{};
            g.q0 = job->group_qptr[i]; g.nq = job->group_qptr[i + 1] - g.q0;
            g.c0 = job->group_cptr[i]; g.nc = job->group_cptr[i + 1] - g.c0;
            MRE_CHECK_ARG(g.nq >= 0 && g.nc >= 0, "group %d has a negative size", i);
            if (g.nq == 0 || g.nc == 0) continue;
            g.item0 = items;
            g.n_qt = (int32_t)((g.nq + tile_q - 1) / tile_q);
            g.n_et = (int32_t)((g.nc + tile_e - 1) / tile_e);
            items += (int64_t)g.n_qt * g.n_et;
            groups.push_back(g);
        }
        if (groups.empty()) {  // nothing to count; keep one empty descriptor so lookups stay in range
            GroupDesc g{};
            groups.push_back(g);
        }
    }
    *total_items = items;
    (void)ctx;
    return MRE_OK;
}

int fill_rank_params(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, int tile_q, int tile_e, cudaStream_t st,
                     RankParams &p) {
    std::vector<GroupDesc> groups;
    int64_t items = 0;
    MRE_TRY(build_groups(ctx, job, tile_q, tile_e, groups, &items));
    MRE_TRY(ctx->tiles.reserve(groups.size() * sizeof(GroupDesc)));
    MRE_CUDA(cudaMemcpyAsync(ctx->tiles.p, groups.data(), groups.size() * sizeof(GroupDesc), cudaMemcpyHostToDevice, st));
    // the descriptor vector dies with this frame; the copy above is from pageable memory and therefore staged
    // synchronously by the runtime before cudaMemcpyAsync returns
    p.E = job->E; p.R = job->R;
    p.q_h = job->q_h; p.q_t = job->q_t; p.q_r = job->q_r; p.q_side = job->q_side; p.side = job->side; p.Q = job->Q;
    p.groups = ctx->tiles.as<GroupDesc>();
    p.n_groups = (int32_t)groups.size();
    p.all_entities = job->n_groups <= 0 ? 1 : 0;
    p.cand_idx = job->cand_idx;
    p.total_items = items;
    p.filter = job->filter;
    p.hr_key = p.hr_val = p.tr_key = p.tr_val = nullptr;
    p.n_all = 0;
    if (job->filter == MRE_FILTER_INDEX) {
        MRE_CHECK_ARG(ix != nullptr, "MRE_FILTER_INDEX needs an index");
        MRE_CHECK_ARG(ix->device == ctx->device, "index is not on device %d (call mre_index_to_device)", ctx->device);
        MRE_CHECK_ARG(ix->E == job->E && ix->R == job->R, "index E/R (%lld/%lld) differ from the job's (%lld/%lld)",
                      (long long)ix->E, (long long)ix->R, (long long)job->E, (long long)job->R);
        p.hr_key = ix->d_all_hr_key; p.hr_val = ix->d_all_hr_val; p.tr_key = ix->d_all_tr_key; p.tr_val = ix->d_all_tr_val;
        p.n_all = ix->n_all;
    } else if (job->filter == MRE_FILTER_CSR) {
        MRE_CHECK_ARG(job->filt_ptr && job->filt_idx, "MRE_FILTER_CSR needs filt_ptr and filt_idx");
    }
    p.filt_ptr = job->filt_ptr; p.filt_idx = job->filt_idx;
    p.counts = job->counts;
    return MRE_OK;
}

template <int P>
static int transe_prepass(mre_ctx *ctx, const mre_rank_job *job, cudaStream_t st, const float **ent_out, int64_t *Dp_out) {
    const int64_t D = job->D, Dp = (D + 3) & ~(int64_t)3;
    const float *ent = job->ent, *rel = job->rel;
    if (job->normalize || Dp != D) {
        MRE_TRY(ctx->ent_n.reserve((size_t)job->E * Dp * sizeof(float)));
        MRE_TRY(ctx->rel_n.reserve((size_t)job->R * Dp * sizeof(float)));
        normalize_rows_kernel<<<(unsigned)((job->E + 127) / 128), 128, 0, st>>>(job->ent, job->E, D, Dp, job->normalize, ctx->ent_n.as<float>());
        normalize_rows_kernel<<<(unsigned)((job->R + 127) / 128), 128, 0, st>>>(job->rel, job->R, D, Dp, job->normalize, ctx->rel_n.as<float>());
        ctx->launches += 2;
        ent = ctx->ent_n.as<float>();
        rel = ctx->rel_n.as<float>();
    }
    if (job->Q > 0) {
        MRE_TRY(ctx->qvec.reserve((size_t)job->Q * Dp * sizeof(float)));
        MRE_TRY(ctx->thr.reserve((size_t)job->Q * sizeof(float2)));
        transe_qvec_kernel<<<grid_for(job->Q * Dp, 256), 256, 0, st>>>(ent, rel, Dp, job->q_h, job->q_t, job->q_r, job->q_side,
                                                                       job->side, job->Q, ctx->qvec.as<float>());
        transe_threshold_kernel<P><<<(unsigned)((job->Q + 127) / 128), 128, 0, st>>>(
            ent, Dp, job->q_h, job->q_t, job->q_side, job->side, job->Q, ctx->qvec.as<float>(), ctx->thr.as<float2>());
        ctx->launches += 2;
    }
    *ent_out = ent;
    *Dp_out = Dp;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

template <int P, bool NEED_EQ>
static int launch_rank(mre_ctx *ctx, const RankParams &p, const CUtensorMap &tm_q, const CUtensorMap &tm_e, cudaStream_t st) {
    auto kern = transe_rank_kernel<P, NEED_EQ>;
    static bool configured = false;
    if (!configured) {
        MRE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RANK_SMEM));
        configured = true;
    }
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(p.total_items, (int64_t)ctx->sm_count * 2));
    kern<<<grid, RANK_THREADS, RANK_SMEM, st>>>(p, tm_q, tm_e);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

int rank_transe(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, cudaStream_t st) {
    MRE_CHECK_ARG(job->p_norm == 1 || job->p_norm == 2, "p_norm must be 1 or 2");
    const float *ent = nullptr;
    int64_t Dp = 0;
    if (job->p_norm == 1) MRE_TRY(transe_prepass<1>(ctx, job, st, &ent, &Dp));
    else MRE_TRY(transe_prepass<2>(ctx, job, st, &ent, &Dp));
    RankParams p{};
    MRE_TRY(fill_rank_params(ctx, ix, job, TILE_Q, TILE_E, st, p));
    p.ent = ent;
    p.D = Dp;
    p.qvec = ctx->qvec.as<float>();
    p.thr = ctx->thr.as<float2>();
    if (job->Q == 0) return MRE_OK;
    // the table the candidate tiles stream from: the entity table itself, or the gathered candidate rows
    const float *cand_table = ent;
    int64_t cand_rows = job->E;
    if (!p.all_entities) {
        cand_rows = job->group_cptr[job->n_groups];
        MRE_CHECK_ARG(cand_rows < (1LL << 31), "too many candidate rows");
        MRE_TRY(ctx->ent_aux.reserve((size_t)std::max<int64_t>(cand_rows, 1) * Dp * sizeof(float)));
        if (cand_rows > 0) {
            gather_rows_kernel<<<grid_for(cand_rows * (Dp >> 2), 256), 256, 0, st>>>(ent, Dp, job->cand_idx, cand_rows,
                                                                                     ctx->ent_aux.as<float>());
            ctx->launches += 1;
        }
        cand_table = ctx->ent_aux.as<float>();
    }
    init_counts_kernel<<<grid_for(4 * job->Q, 256), 256, 0, st>>>(job->counts, 4 * job->Q);
    ctx->launches += 1;
    MRE_TRY(build_tile_filter(ctx, job, p, TILE_Q, TILE_E, st));
    CUtensorMap tm_q, tm_e;
    MRE_TRY(make_tmap_f32_2d(&tm_q, p.qvec, job->Q, Dp, Dp, TILE_Q, CHUNK));
    MRE_TRY(make_tmap_f32_2d(&tm_e, cand_table, std::max<int64_t>(cand_rows, 1), Dp, Dp, TILE_E, CHUNK));
    MRE_TRY(ctx->time_begin(st));
    // the tie count is always on: one extra compare per score, in the epilogue only
    if (job->p_norm == 1) MRE_TRY((launch_rank<1, true>(ctx, p, tm_q, tm_e, st)));
    else MRE_TRY((launch_rank<2, true>(ctx, p, tm_q, tm_e, st)));
    MRE_TRY(ctx->time_end(st));
    return MRE_OK;
}

int predict_transe(mre_ctx *ctx, const mre_rank_job *job, int64_t query, float *scores_out, cudaStream_t st) {
    mre_rank_job one = *job;
    one.q_h = job->q_h + query; one.q_t = job->q_t + query; one.q_r = job->q_r + query;
    one.q_side = job->q_side ? job->q_side + query : nullptr;
    one.Q = 1;
    const float *ent = nullptr;
    int64_t Dp = 0;
    if (job->p_norm == 1) MRE_TRY(transe_prepass<1>(ctx, &one, st, &ent, &Dp));
    else MRE_TRY(transe_prepass<2>(ctx, &one, st, &ent, &Dp));
    unsigned grid = (unsigned)((job->E + 127) / 128);
    if (job->p_norm == 1) transe_predict_kernel<1><<<grid, 128, 0, st>>>(ent, job->E, Dp, ctx->qvec.as<float>(), scores_out);
    else transe_predict_kernel<2><<<grid, 128, 0, st>>>(ent, job->E, Dp, ctx->qvec.as<float>(), scores_out);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

}  // namespace mre
