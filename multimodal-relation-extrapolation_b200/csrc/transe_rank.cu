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
// The main kernel is persistent, ONE 16-warp CTA per SM: each CTA walks 256-query x 128-entity work items; boxes of both
// operands stream into a 4-stage shared-memory ring through TMA (cp.async.bulk.tensor, 128-byte swizzle, mbarrier
// complete_tx; candidate lists are gathered into a dense table by a pre-pass), issued by whichever warp is last
// to release a stage; the sixteen warps hold an 8x8 register micro-tile per thread and read the ring with
// conflict-free 128-bit LDS.  Why one big CTA: the SM sub-partition arbiter is greedy, so warps sharing a ring drift to
// the ring's limit and the fast ones sleep; with two 8-warp CTAs per SM each CTA's pace was set by its least-favoured warp
// on ANY sub-partition and the favoured warps of both CTAs idled together (measured: 0.826 of the FP32 add peak).  With all
// sixteen warps of the SM on one ring every sub-partition holds the same fixed work per chunk whatever the arbiter prefers,
// and early finishers run their epilogue while the others still compute (0.855).  The arithmetic is PACKED: query vectors are stored pair-interleaved ([slot / 2][d][slot & 1])
// so one 64-bit register pair holds element d of two adjacent queries, and each (subtract, add-|.|) step is one
// sub.f32x2 + one add.f32x2 (SASS FADD2, the entity value as a scalar-broadcast operand) for two (query, entity) pairs:
// half the issue slots of the scalar form, which leaves the FP32 pipe -- not instruction issue -- as the bound, with the
// LDS / barrier / address instructions issuing in the shadow of the two-cycle packed instructions.  Per (q, e) pair the
// accumulation is still sequential over d with the same roundings, so counts stay bit-identical to the oracle.  The
// epilogue compares the 64 accumulators against the per-query thresholds, reduces the counts with warp shuffles and adds them
// to the raw AND the filtered counters of the query; the known-true entities of each query (a handful) are scored by
// transe_known_kernel with the same accumulation order and taken back out of the filtered counters.  The kernel is bound by the FP32 pipe: 2 lane-ops per (q, e, d).
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "common.h"
#include "rank_common.cuh"
#include "rank_host.h"
#include "tma_host.h"

namespace mre {

constexpr int CHUNK = 32;            // floats of D per pipeline stage (128 B per row)
#ifndef MRE_TRANSE_TQ
#define MRE_TRANSE_TQ 256
#endif
#ifndef MRE_TRANSE_STAGES
#define MRE_TRANSE_STAGES 4
#endif
#ifndef MRE_TRANSE_CTAS
#define MRE_TRANSE_CTAS 1
#endif
constexpr int TQ = MRE_TRANSE_TQ;                  // queries per work item of this kernel
constexpr int STAGES = MRE_TRANSE_STAGES;
constexpr int CTAS_PER_SM = MRE_TRANSE_CTAS;
#ifndef MRE_RANK_WARPS
#define MRE_RANK_WARPS 16     // one 512-thread CTA per SM, four warps per SM sub-partition, all on ONE ring (see the kernel comment)
#endif
constexpr int CONSUMER_WARPS = MRE_RANK_WARPS;
constexpr int RANK_THREADS = CONSUMER_WARPS * 32;
constexpr int NTQ = RANK_THREADS / 16;              // threads along the query dimension of the tile
constexpr int NI = (TQ / 2) / NTQ;              // query PAIR-rows per thread (4 with 8 warps: an 8-query x 8-entity register tile)
constexpr uint32_t QBOX_BYTES = (TQ / 2) * 128;             // 64 query pair-rows x 128 B (16 d-values of two queries)
constexpr uint32_t STAGE_BYTES = (TQ + TILE_E) * CHUNK * 4;  // two 8 KiB query-pair boxes + one 16 KiB candidate box
constexpr size_t RANK_SMEM = 1024 + (size_t)STAGES * STAGE_BYTES + STAGES * sizeof(uint64_t) + 4 * sizeof(int) + (size_t)TQ * sizeof(float2) + 64;

// ------------------------------------------------------------------------------------------ scalar scorer
// The one definition of a TransE accumulator: sequential over d, acc = acc + |v - e| (p = 1) or fma(u, u, acc) (p = 2), with
// v_d = h_d + r_d (tail query) or -(r_d - t_d) (head query), so that |v_d - e_d| is bit-identical to the reference's element
// (|(h + r) - e| and |e + (r - t)|, TransE.py:55-58).  `a` is the fixed entity's row, `r` the relation's row.
template <int P>
__device__ __forceinline__ float transe_acc(const float *__restrict__ a, const float *__restrict__ r, int side,
                                            const float *__restrict__ e, int64_t D) {
    float acc = 0.f;
    const int n = (int)D;
#pragma unroll 16
    for (int d = 0; d < n; d += 4) {   // unrolled: the row fetches of 16 steps are in flight together (latency-bound callers)
        const float4 x = __ldg(reinterpret_cast<const float4 *>(a + d));
        const float4 y = __ldg(reinterpret_cast<const float4 *>(r + d));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(e + d));
        float v0, v1, v2, v3;
        if (side) { v0 = x.x + y.x; v1 = x.y + y.y; v2 = x.z + y.z; v3 = x.w + y.w; }
        else { v0 = -(y.x - x.x); v1 = -(y.y - x.y); v2 = -(y.z - x.z); v3 = -(y.w - x.w); }
        const float u0 = v0 - b.x, u1 = v1 - b.y, u2 = v2 - b.z, u3 = v3 - b.w;
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

// Everything per query in ONE launch:
//   * v_q = h + r (tail query) or -(r - t) (head query), written PAIR-INTERLEAVED by slot: a query's slot is its position in
//     its candidate group counted from the group's (even) first slot, element d of slot s lives at [s / 2][d][s & 1];
//     queries of dropped (empty) groups own no slot;
//   * the thresholds from the true entity's accumulator, which must carry the tile kernel's (and the oracle's) SEQUENTIAL order
//     over d.  A warp takes 8 queries at a time: for every 64-wide chunk of d the lanes first work ACROSS d (query by query:
//     coalesced row reads, v written out, u_d = v_d - e_true,d -- the same single rounding as transe_acc -- parked in shared
//     memory), then ALONG d (lane l < 8 accumulates query l's u_d in order): 8 sequential sums side by side instead of one
//     lane of a warp paying a whole issue slot per instruction.  p = 1: score = acc.  p = 2: score = sqrt(acc) and the
//     reference compares the square roots, so lo = min{x : sqrt(x) >= s_true}, hi = min{x : sqrt(x) > s_true} (sqrt is
//     monotone), which lets the tile kernel compare raw accumulators and still agree with sqrtf(acc_j) < sqrtf(acc_true) exactly;
//   * the zeroing of the query's four counters.
constexpr int TQ_WARPS = 8;
constexpr int TQ_G = 8;
constexpr int TQ_CH = 64;
template <int P>
__global__ void __launch_bounds__(TQ_WARPS * 32) transe_query_kernel(const RankParams p, const float *__restrict__ rel,
                                                                      float *__restrict__ qvec, float2 *__restrict__ thr) {
    __shared__ float sU[TQ_WARPS][TQ_G][TQ_CH + 1];      // +1: conflict-free both ways
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t D = p.D;
    const int64_t n_groups = (p.Q + TQ_G - 1) / TQ_G;
    for (int64_t grp = (int64_t)blockIdx.x * TQ_WARPS + warp; grp < n_groups; grp += (int64_t)gridDim.x * TQ_WARPS) {
        const int64_t q0 = grp * TQ_G;
        const int nq = (int)min((int64_t)TQ_G, p.Q - q0);
        // lane l < nq holds the descriptors of query q0 + l
        int my_s = 0, my_fixed = 0, my_truth = 0, my_r = 0;
        long long my_slot = -1;                                      // -1: the query owns no slot (its group was dropped)
        if (lane < nq) {
            const int64_t q = q0 + lane;
            const GroupDesc &gd = p.groups[p.n_groups > 1 ? group_of_query(p, q) : 0];
            if (q >= gd.q0 && q - gd.q0 < gd.nq) my_slot = gd.s0 + (q - gd.q0);
            my_s = query_side(p, q);
            my_fixed = (int)(my_s ? p.q_h[q] : p.q_t[q]);
            my_truth = (int)(my_s ? p.q_t[q] : p.q_h[q]);
            my_r = (int)p.q_r[q];
        }
        float acc = 0.f;
        for (int64_t c0 = 0; c0 < D; c0 += TQ_CH) {
            // every row element this chunk needs, for all the group's queries, is requested BEFORE anything is used: one memory
            // round trip per chunk instead of one per query (the pre-pass runs on cold caches: the chain of trips is its cost)
            constexpr int J = TQ_CH / 32;
            float xa[TQ_G][J], xr[TQ_G][J], xe[TQ_G][J];
            int sq[TQ_G];
            long long sl[TQ_G];
#pragma unroll
            for (int qi = 0; qi < TQ_G; qi++) {
                sq[qi] = __shfl_sync(0xffffffffu, my_s, qi);
                sl[qi] = __shfl_sync(0xffffffffu, my_slot, qi);
                const float *a = p.ent + (int64_t)__shfl_sync(0xffffffffu, my_fixed, qi) * D;
                const float *e = p.ent + (int64_t)__shfl_sync(0xffffffffu, my_truth, qi) * D;
                const float *rr = rel + (int64_t)__shfl_sync(0xffffffffu, my_r, qi) * D;
#pragma unroll
                for (int j = 0; j < J; j++) {
                    const int64_t d = c0 + lane + 32 * j;
                    const bool live = qi < nq && d < D;
                    xa[qi][j] = live ? a[d] : 0.f;
                    xr[qi][j] = live ? rr[d] : 0.f;
                    xe[qi][j] = live ? e[d] : 0.f;
                }
            }
#pragma unroll
            for (int qi = 0; qi < TQ_G; qi++) {
                if (qi >= nq) break;                                            // warp-uniform
                const int s = sq[qi];
                const long long slot = sl[qi];
                float *qrow = qvec + (slot >> 1) * (2 * D) + (slot & 1);
#pragma unroll
                for (int j = 0; j < J; j++) {
                    const int64_t d = c0 + lane + 32 * j;
                    float u = 0.f;
                    if (d < D) {
                        const float rv = xr[qi][j];
                        const float v = s ? xa[qi][j] + rv : -(rv - xa[qi][j]);
                        if (slot >= 0) qrow[2 * d] = v;
                        u = v - xe[qi][j];
                    }
                    sU[warp][qi][lane + 32 * j] = u;
                }
            }
            __syncwarp();
            if (lane < nq) {
                const int nd = (int)min((int64_t)TQ_CH, D - c0);
#pragma unroll 8
                for (int dd = 0; dd < nd; dd++) {
                    const float u = sU[warp][lane][dd];
                    acc = P == 1 ? acc + fabsf(u) : fmaf(u, u, acc);
                }
            }
            __syncwarp();
        }
        if (lane < nq) {
            const int64_t q = q0 + lane;
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
#pragma unroll
            for (int c = 0; c < 4; c++) p.counts[(int64_t)c * p.Q + q] = 0;
        }
    }
}

// Model.predict for one query: the materialised float32[E] score vector (tests / drop-in callers only)
template <int P>
__global__ void transe_predict_kernel(const float *__restrict__ ent, const float *__restrict__ rel, int64_t E, int64_t D,
                                      const int64_t *__restrict__ q_h, const int64_t *__restrict__ q_t,
                                      const int64_t *__restrict__ q_r, const uint8_t *__restrict__ q_side, int side,
                                      float *__restrict__ out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= E) return;
    const int s = q_side ? (int)q_side[0] : side;
    const int64_t fixed = s ? q_h[0] : q_t[0];
    float acc = transe_acc<P>(ent + fixed * D, rel + q_r[0] * D, s, ent + j * D, D);
    out[j] = P == 1 ? acc : __fsqrt_rn(acc);
}

// ------------------------------------------------------------------------------------------ main kernel
// The filtered counts' correction (rank_common.cuh: known_correction) with the TransE element: v_d = h_d + r_d (tail query) or
// -(r_d - t_d) (head query), term u = v_d - e_d (one rounding), fold acc + |u| (p = 1) or fma(u, u, acc) (p = 2) -- transe_acc's
// steps exactly -- and the tile kernel's threshold compare.
template <int P>
struct TranseKnownOp {
    const float *ent, *rel, *a, *r;
    const float2 *thr;
    int64_t D;
    int side;
    float2 th;
    __device__ __forceinline__ void query(int64_t q, int s, int64_t h, int64_t t, int64_t rr) {
        side = s;
        a = ent + (s ? h : t) * D;
        r = rel + rr * D;
    }
    __device__ __forceinline__ float vec(int d) const { return side ? __ldg(a + d) + __ldg(r + d) : -(__ldg(r + d) - __ldg(a + d)); }
    // flattened form: the query vector as transe_query_kernel wrote it (the same v), pair-interleaved by slot
    static constexpr int VSTRIDE = 2;
    const float *qvec;
    __device__ __forceinline__ const float *flat_query(int64_t q, int64_t slot, int, int64_t, int64_t, int64_t) {
        th = thr[q];
        return qvec + (slot >> 1) * (2 * D) + (slot & 1);
    }
    static constexpr bool DIRECT_ONLY = false;
    __device__ __forceinline__ float direct(int64_t x) const { return transe_acc<P>(a, r, side, ent + x * D, D); }
    __device__ __forceinline__ void thresholds(int64_t q) { th = thr[q]; }
    __device__ __forceinline__ bool truth_ties() const { return th.x < th.y; }       // what the tile kernel decides for acc_true
    __device__ __forceinline__ float term(float v, float e) const { return v - e; }
    __device__ __forceinline__ float fold(float acc, float u) const { return P == 1 ? acc + fabsf(u) : fmaf(u, u, acc); }
    __device__ __forceinline__ void classify(float acc, int &lt, int &eq) const {
        if (acc < th.x) lt++;
        else if (acc < th.y) eq++;
    }
};
template <int P>
__global__ void __launch_bounds__(KNOWN_WARPS * 32) transe_known_kernel(const RankParams p, const float *__restrict__ rel) {
    __shared__ float sT[KNOWN_WARPS][32][33];
    __shared__ int64_t sX[KNOWN_WARPS][32];
    TranseKnownOp<P> op{p.ent, rel, nullptr, nullptr, p.thr, p.D, 0, make_float2(0.f, 0.f), p.qvec};
    known_correction(p, op, sT[threadIdx.x >> 5], sX[threadIdx.x >> 5]);
}
template <int P>
__global__ void __launch_bounds__(KNOWN_WARPS * 32, 3) transe_known_score_kernel(const RankParams p, const float *__restrict__ rel, const KnownRuns kr) {
    __shared__ float sT[KNOWN_WARPS][32][33];
    __shared__ int64_t sX[KNOWN_WARPS][32];
    TranseKnownOp<P> op{p.ent, rel, nullptr, nullptr, p.thr, p.D, 0, make_float2(0.f, 0.f), p.qvec};
    known_score_runs(p, kr, op, sT[threadIdx.x >> 5], sX[threadIdx.x >> 5]);
}
template <int P>
__global__ void __launch_bounds__(KNOWN_WARPS * 32) transe_known_compare_kernel(const RankParams p, const KnownRuns kr) {
    TranseKnownOp<P> op{p.ent, nullptr, nullptr, nullptr, p.thr, p.D, 0, make_float2(0.f, 0.f), p.qvec};
    known_compare_runs(p, kr, op);
}
template <int P>
__global__ void __launch_bounds__(KNOWN_WARPS * 32) transe_known_flat_kernel(const RankParams p) {
    __shared__ float sT[KNOWN_WARPS][32][33];
    __shared__ int64_t sX[KNOWN_WARPS][32];
    __shared__ const float *sV[KNOWN_WARPS][32];
    TranseKnownOp<P> op{p.ent, nullptr, nullptr, nullptr, p.thr, p.D, 0, make_float2(0.f, 0.f), p.qvec};
    known_correction_flat(p, op, sT[threadIdx.x >> 5], sX[threadIdx.x >> 5], sV[threadIdx.x >> 5]);
}

// acc (two queries' accumulators against one entity) <- one more element d: u = q - e ; acc + |u|  (p = 1) | fma(u, u, acc)
// ptxas turns the {e, e} pair into a scalar-broadcast operand and folds the |.| into the add: FADD2 R, R.F32x2, -R.F32 ;
// FADD2 R, R.F32x2, |R|.F32x2  (FFMA2 for p = 2): two instructions for four lane-ops.
#ifndef MRE_VOLATILE
#define MRE_VOLATILE 1
#endif
#if MRE_VOLATILE
#define MRE_ASM asm volatile   // keeps the source order of the packed instructions: ptxas otherwise pulls dependent pairs together
#else
#define MRE_ASM asm
#endif
template <int P>
__device__ __forceinline__ unsigned long long sub2(unsigned long long q, float e) {
    unsigned long long u;
    MRE_ASM("{\n\t.reg .b64 t;\n\tmov.b64 t, {%2, %2};\n\tsub.f32x2 %0, %1, t;\n\t}" : "=l"(u) : "l"(q), "f"(e));
    return u;
}
template <int P>
__device__ __forceinline__ void acc2(unsigned long long &acc, unsigned long long u) {
    if (P == 1) {
        MRE_ASM("{\n\t.reg .b64 a;\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %1;\n\tabs.f32 lo, lo;\n\tabs.f32 hi, hi;\n\tmov.b64 a, {lo, hi};\n\t"
            "add.f32x2 %0, %0, a;\n\t}" : "+l"(acc) : "l"(u));
    } else {
        MRE_ASM("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc) : "l"(u));
    }
}
#ifndef MRE_K4_UNROLL
#define MRE_K4_UNROLL 4
#endif
constexpr int K4_UNROLL = MRE_K4_UNROLL;
#ifndef MRE_DIAG_NOWAIT
#define MRE_DIAG_NOWAIT 0
#endif
#ifndef MRE_DIAG_NOEPI
#define MRE_DIAG_NOEPI 0
#endif
#ifndef MRE_WAIT_SPIN
#define MRE_WAIT_SPIN 1
#endif
#ifndef MRE_GROUP
#define MRE_GROUP 4        // subtractions issued together before their dependent adds: hides the FADD2 latency inside ONE warp
#endif

// One pipeline chunk = CHUNK floats of d of every row of one work item's two operands: the interleaved query pairs as two
// boxes of 64 pair-rows x 128 B (16 d-values of two queries each), the candidates as one box of 128 rows x 128 B, all
// written into shared memory with the hardware 128-byte swizzle.  Issued by ONE thread.
__device__ __forceinline__ void issue_chunk(const RankParams &p, const CUtensorMap *tm_q, const CUtensorMap *tm_e, int64_t flat,
                                            int n_chunks, uint32_t ring_u32, uint32_t full0) {
    const int64_t item = cta_item(p, flat / n_chunks);
    if (item < 0) return;
    const int c = (int)(flat % n_chunks);
    int g, qt, et;
    decode_item(p, item, g, qt, et);
    const GroupDesc &gd = p.groups[g];
    const int prow = (int)((gd.s0 + (int64_t)qt * TQ) >> 1);
    const int erow = (int)(gd.c0 + (int64_t)et * TILE_E);
    const int stage = (int)(flat % STAGES);
    const uint32_t full = full0 + 8 * stage;
    const uint32_t sq = ring_u32 + (uint32_t)stage * STAGE_BYTES;
    mbar_arrive_expect_tx(full, STAGE_BYTES);
    tma_load_2d(sq, tm_q, c * (2 * CHUNK), prow, full);
    tma_load_2d(sq + QBOX_BYTES, tm_q, c * (2 * CHUNK) + CHUNK, prow, full);
    tma_load_2d(sq + 2 * QBOX_BYTES, tm_e, c * CHUNK, erow, full);
}

template <int P, bool NEED_EQ>
__global__ void __launch_bounds__(RANK_THREADS, CTAS_PER_SM)
transe_rank_kernel(const RankParams p, const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_e) {
    extern __shared__ unsigned char smem_raw[];
    // the 128-byte swizzle pattern is a function of the shared-memory address: tiles must start 1024-byte aligned
    const uint32_t ring_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *ring = smem_raw + (ring_u32 - smem_u32(smem_raw));
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring + (size_t)STAGES * STAGE_BYTES);
    int *done = reinterpret_cast<int *>(bars + STAGES);  // per-stage count of warps that finished reading the stage
    float2 *thr_s = reinterpret_cast<float2 *>(done + 4);   // [warps][4 NI rows]
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

    // entity rows te + 16 j (j < 8); query PAIR-rows tq + NTQ i (i < NI) = tile rows 2 (tq + NTQ i) and 2 (tq + NTQ i) + 1.
    // A row keeps its logical 16-byte chunk k at (k ^ (row & 7)), and (te + 16 j) & 7 == te & 7: the sixteen rows one
    // LDS.128 touches split into two conflict-free wavefronts; the two pair-rows a warp reads are broadcasts.
    const int te = threadIdx.x & 15, tq = threadIdx.x >> 4;
    const int xe = te & 7, xq = tq & 7;
    int64_t it = 0;
    for (int64_t k_item = 0;; k_item++) {
        const int64_t item = cta_item(p, k_item);
        if (item < 0) break;
        int g, qt, et;
        decode_item(p, item, g, qt, et);
        const GroupDesc gd = p.groups[g];
        const int64_t qbase = gd.q0 + (int64_t)qt * TQ;
        const int nq = (int)min((int64_t)TQ, gd.q0 + gd.nq - qbase);
        const int ne = (int)min((int64_t)TILE_E, gd.nc - (int64_t)et * TILE_E);
        const int ni_act = (nq + 2 * NTQ - 1) / (2 * NTQ);    // slabs of NTQ pair-rows (2 NTQ queries) holding at least one query
        // thresholds of this warp's 16 query rows -> shared memory (asynchronous copy: nobody waits for it before the epilogue);
        // slot k of the warp = row 2 ((2 warp + k / 2NI) + NTQ ((k % 2NI) / 2)) + k % 2; rows past the group's end get -inf (never counted)
        if (lane < 4 * NI) {
            const int k2 = lane % (2 * NI);
            const int ql = 2 * ((2 * warp + lane / (2 * NI)) + NTQ * (k2 >> 1)) + (k2 & 1);
            float2 *dst = thr_s + warp * (4 * NI) + lane;
            if (ql < nq) cp_async_8(smem_u32(dst), p.thr + qbase + ql);
            else *dst = make_float2(-INFINITY, -INFINITY);
        }
        cp_async_commit();

        unsigned long long acc[NI][8];   // acc[i][j] = (row 2 (tq + NTQ i), row 2 (tq + NTQ i) + 1) x entity te + 16 j
#pragma unroll
        for (int i = 0; i < NI; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[i][j] = 0ull;

        // Full tiles run the unguarded loop (ptxas hoists the loads of the next slab over the current one's arithmetic); a
        // partial query tile -- the last tile of a group, frequent with per-relation candidate groups -- runs a second copy
        // of the loop that skips its empty 64-query slabs with a uniform branch.
        auto tile_loop = [&](auto guard_tag) {
        constexpr bool GUARD = decltype(guard_tag)::value;
        for (int c = 0; c < n_chunks; c++, it++) {
            const int stage = (int)(it % STAGES);
#if MRE_DIAG_NOWAIT
            // diagnostic build only (wrong results): compute on whatever is in the ring
#elif MRE_WAIT_SPIN
            mbar_wait_spin(full0 + 8 * stage, (uint32_t)((it / STAGES) & 1));
#else
            mbar_wait(full0 + 8 * stage, (uint32_t)((it / STAGES) & 1));
#endif
            const unsigned char *sQ = ring + (size_t)stage * STAGE_BYTES + tq * 128;
            const unsigned char *sE = ring + (size_t)stage * STAGE_BYTES + 2 * QBOX_BYTES + te * 128;
            const int nk4 = (int)min((int64_t)CHUNK, p.D - (int64_t)c * CHUNK) >> 2;
#pragma unroll K4_UNROLL
            for (int k4 = 0; k4 < nk4; k4++) {
                const int oe = (k4 ^ xe) << 4;
                // four d-values of a query pair = 32 B = 16-byte chunks 2 (k4 & 3) and 2 (k4 & 3) + 1 of box k4 >> 2
                const int qb = (k4 >> 2) * (int)QBOX_BYTES;
                const int oq0 = qb + ((((k4 & 3) << 1) ^ xq) << 4), oq1 = qb + (((((k4 & 3) << 1) | 1) ^ xq) << 4);
                float4 ev[8];
#pragma unroll
                for (int j = 0; j < 8; j++) ev[j] = *reinterpret_cast<const float4 *>(sE + j * (16 * 128) + oe);
#pragma unroll
                for (int i = 0; i < NI; i++) {
                    if (GUARD && i >= ni_act) break;           // uniform: a partial query tile skips its empty 64-query slabs
                    const ulonglong2 qa = *reinterpret_cast<const ulonglong2 *>(sQ + i * (NTQ * 128) + oq0);
                    const ulonglong2 qc = *reinterpret_cast<const ulonglong2 *>(sQ + i * (NTQ * 128) + oq1);
                    // per accumulator the order is d, d+1, d+2, d+3 (sequential sum); across accumulators the work is
                    // grouped MRE_GROUP subtractions, then their MRE_GROUP adds
#pragma unroll
                    for (int dd = 0; dd < 4; dd++) {
                        const unsigned long long qd = dd == 0 ? qa.x : dd == 1 ? qa.y : dd == 2 ? qc.x : qc.y;
#pragma unroll
                        for (int j0 = 0; j0 < 8; j0 += MRE_GROUP) {
                            unsigned long long u[MRE_GROUP];
#pragma unroll
                            for (int j = 0; j < MRE_GROUP; j++) {
                                const float4 e4 = ev[j0 + j];
                                u[j] = sub2<P>(qd, dd == 0 ? e4.x : dd == 1 ? e4.y : dd == 2 ? e4.z : e4.w);
                            }
#pragma unroll
                            for (int j = 0; j < MRE_GROUP; j++) acc2<P>(acc[i][j0 + j], u[j]);
                        }
                    }
                }
            }
            // Release the stage.  The LAST warp to finish reading it refills it with the chunk STAGES ahead: no
            // dedicated producer warp, no empty-barrier spinning, and warps may drift up to STAGES-1 chunks apart.
            __syncwarp();
            if (lane == 0 && !MRE_DIAG_NOWAIT) {
                __threadfence_block();
                const int old = atomicAdd(&done[stage], 1);
                if ((old & (CONSUMER_WARPS - 1)) == CONSUMER_WARPS - 1) {
                    __threadfence_block();
                    fence_proxy_async();
                    issue_chunk(p, &tm_q, &tm_e, it + STAGES, n_chunks, ring_u32, full0);
                }
            }
        }

        };
        if (ni_act == NI) tile_loop(std::false_type{});
        else tile_loop(std::true_type{});

#if MRE_DIAG_NOEPI
        {   // diagnostic build only (wrong results): keep the accumulators live, skip the compare / count epilogue
            unsigned long long x = 0;
#pragma unroll
            for (int i = 0; i < NI; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) x ^= acc[i][j];
            if (x == 0x123456789abcdefull) atomicAdd(p.counts, 1);
            continue;
        }
#endif
        cp_async_wait_all();
        // ---- epilogue: compare against the per-query thresholds, count, reduce over the 16 lanes sharing a query row.
        // lt <=> s < th.x ; eq <=> th.x <= s < th.y, so eq = #(s < th.y) - #(s < th.x): two independent compare-and-count
        // chains per accumulator.  The thresholds of the warp's 16 rows were staged in shared memory at item start.
        // Known-true entities are NOT looked at here: transe_known_kernel scores each of them with the same arithmetic and takes
        // them out of the filtered counters, so raw and filtered counts receive the same additions from this kernel.
        __syncwarp();
        const float2 *th_w = thr_s + warp * (4 * NI) + (tq & 1) * (2 * NI);
        if (ne == TILE_E) {
            // fast path (every tile but a group's last): no candidate padding.  Two rows share one
            // word (four 8-bit counters, each at most 128 after the reduction), so 4 shuffles serve 2 rows.
#pragma unroll
            for (int i = 0; i < NI; i++) {
                const float2 t0 = th_w[2 * i], t1 = th_w[2 * i + 1];
                uint32_t a_x = 0, a_y = 0, b_x = 0, b_y = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float s0 = __uint_as_float((uint32_t)acc[i][j]), s1 = __uint_as_float((uint32_t)(acc[i][j] >> 32));
                    a_x += s0 < t0.x ? 1u : 0u; a_y += s0 < t0.y ? 1u : 0u;
                    b_x += s1 < t1.x ? 1u : 0u; b_y += s1 < t1.y ? 1u : 0u;
                }
                uint32_t packed = a_x | ((a_y - a_x) << 8) | (b_x << 16) | ((b_y - b_x) << 24);
#pragma unroll
                for (int m = 1; m < 16; m <<= 1) packed += __shfl_xor_sync(0xffffffffu, packed, m);
                if (te == 0 && packed) {
                    const int ql = 2 * (tq + NTQ * i);
                    const int64_t q = qbase + ql;
                    const int lt0 = packed & 0xff, eq0 = (packed >> 8) & 0xff, lt1 = (packed >> 16) & 0xff, eq1 = packed >> 24;
                    if (lt0) { atomicAdd(p.counts + q, lt0); atomicAdd(p.counts + 2 * p.Q + q, lt0); }
                    if (NEED_EQ && eq0) { atomicAdd(p.counts + p.Q + q, eq0); atomicAdd(p.counts + 3 * p.Q + q, eq0); }
                    if (lt1) { atomicAdd(p.counts + q + 1, lt1); atomicAdd(p.counts + 2 * p.Q + q + 1, lt1); }
                    if (NEED_EQ && eq1) { atomicAdd(p.counts + p.Q + q + 1, eq1); atomicAdd(p.counts + 3 * p.Q + q + 1, eq1); }
                }
            }
        } else {
            // candidate padding of a group's last tile: columns past the group's end never count
#pragma unroll
            for (int i2 = 0; i2 < 2 * NI; i2++) {             // i2 = 2 i + half (unrolled: register arrays need static indices)
                const int ql = 2 * (tq + NTQ * (i2 >> 1)) + (i2 & 1);
                const float2 th = th_w[i2];
                uint32_t packed = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const unsigned long long a2 = acc[i2 >> 1][j];
                    const float sc = __uint_as_float((i2 & 1) ? (uint32_t)(a2 >> 32) : (uint32_t)a2);
                    const bool ok = (te + 16 * j) < ne;
                    const bool lt = ok && sc < th.x;
                    const bool eq = NEED_EQ && ok && !lt && sc < th.y;
                    packed += (lt ? 1u : 0u) + (eq ? 0x100u : 0u);
                }
#pragma unroll
                for (int m = 1; m < 16; m <<= 1) packed += __shfl_xor_sync(0xffffffffu, packed, m);
                if (te == 0 && packed) {
                    const int64_t q = qbase + ql;
                    const int n_lt = packed & 0xff, n_eq = (packed >> 8) & 0xff;
                    if (n_lt) { atomicAdd(p.counts + q, n_lt); atomicAdd(p.counts + 2 * p.Q + q, n_lt); }
                    if (NEED_EQ && n_eq) { atomicAdd(p.counts + p.Q + q, n_eq); atomicAdd(p.counts + 3 * p.Q + q, n_eq); }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ host side
static int build_groups(mre_ctx *ctx, const mre_rank_job *job, int tile_q, int tile_e, std::vector<GroupDesc> &groups,
                        int64_t *total_items, int64_t *total_slots, int64_t *total_pitems) {
    groups.clear();
    int64_t items = 0, slots = 0, pitems = 0;
    if (job->n_groups <= 0) {
        GroupDesc g{};
        g.q0 = 0; g.nq = job->Q; g.c0 = 0; g.nc = job->E; g.item0 = 0;
        g.n_qt = (int32_t)((job->Q + tile_q - 1) / tile_q);
        g.n_et = (int32_t)((job->E + tile_e - 1) / tile_e);
        items = (int64_t)g.n_qt * g.n_et;
        pitems = (int64_t)((g.n_qt + 1) / 2) * g.n_et;
        slots = (job->Q + 1) & ~(int64_t)1;
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
            g.s0 = slots;
            slots += (g.nq + 1) & ~(int64_t)1;
            g.n_qt = (int32_t)((g.nq + tile_q - 1) / tile_q);
            g.n_et = (int32_t)((g.nc + tile_e - 1) / tile_e);
            items += (int64_t)g.n_qt * g.n_et;
            g.pitem0 = pitems;
            pitems += (int64_t)((g.n_qt + 1) / 2) * g.n_et;
            groups.push_back(g);
        }
        if (groups.empty()) {  // nothing to count; keep one empty descriptor so lookups stay in range
            GroupDesc g{};
            groups.push_back(g);
        }
    }
    *total_items = items;
    *total_slots = slots;
    *total_pitems = pitems;
    (void)ctx;
    return MRE_OK;
}

int fill_rank_params(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, int tile_q, int tile_e, cudaStream_t st,
                     RankParams &p, std::vector<GroupDesc> *groups_out) {
    std::vector<GroupDesc> groups;
    int64_t items = 0, slots = 0, pitems = 0;
    MRE_TRY(build_groups(ctx, job, tile_q, tile_e, groups, &items, &slots, &pitems));
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
    p.total_slots = slots;
    p.total_pitems = pitems;
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
    p.filt_nnz = job->filter == MRE_FILTER_CSR ? std::max<int64_t>(job->filt_nnz, 0) : 0;
    if (job->filter == MRE_FILTER_NONE) p.filt_ptr = p.filt_idx = nullptr;
    p.counts = job->counts;
    if (groups_out) groups_out->swap(groups);
    return MRE_OK;
}

// scratch of the shared-run known-true pass (MRE_FILTER_INDEX): score + stamp columns parallel to the index's payload columns,
// the per-query run ranges, and the job's epoch (stamps of earlier jobs are simply stale: no clearing between jobs)
int known_runs_scratch(mre_ctx *ctx, const RankParams &p, KnownRuns &kr) {
    const size_t n = (size_t)std::max<int64_t>(p.n_all, 1);
    const size_t need = 2 * n * sizeof(unsigned int);
    if (ctx->known_stamp.cap < need || ctx->known_n != n || ctx->known_epoch >= 0xfffffffeu) {
        // first use, another index on this context (the second column moves), or the epochs ran out: start from clean stamps
        MRE_TRY(ctx->known_stamp.reserve(need));
        MRE_CUDA(cudaMemset(ctx->known_stamp.p, 0, ctx->known_stamp.cap));
        ctx->known_epoch = 0;
        ctx->known_n = n;
    }
    ctx->known_epoch += 1;
    MRE_TRY(ctx->known_score.reserve(2 * n * sizeof(float)));
    MRE_TRY(ctx->known_range.reserve((size_t)std::max<int64_t>(p.Q, 1) * 2 * sizeof(int64_t)));
    kr.score0 = ctx->known_score.as<float>(); kr.score1 = kr.score0 + n;
    kr.stamp0 = ctx->known_stamp.as<unsigned int>(); kr.stamp1 = kr.stamp0 + n;
    kr.range = ctx->known_range.as<int64_t>();
    kr.epoch = ctx->known_epoch;
    return MRE_OK;
}

// Candidate groups give work items of UNEQUAL cost: the last query tile of a group holds 1 .. TQ queries and the kernel skips its
// empty 64-query slabs, and a job is only a few waves of items (FB15K-237-ZS: 680 items over 148 CTAs), so the round-robin order
// leaves whole CTAs idle through the last wave.  Longest-processing-time-first: items sorted by cost (stable), each dealt to the
// least-loaded CTA; a CTA walks its list in that order (expensive items first, the cheap partial tiles fill the tail).
static int build_item_schedule(mre_ctx *ctx, const std::vector<GroupDesc> &groups, int64_t total_items, int grid, cudaStream_t st,
                               RankParams &p) {
    struct Item { int32_t id; float cost; };
    std::vector<Item> items;
    items.reserve((size_t)total_items);
    for (const GroupDesc &g : groups) {
        for (int et = 0; et < g.n_et; et++)
            for (int qt = 0; qt < g.n_qt; qt++) {
                const int64_t nq = std::min<int64_t>(TQ, g.nq - (int64_t)qt * TQ);
                const int slabs = (int)((nq + 2 * NTQ - 1) / (2 * NTQ));
                // full tile: NI slabs at full speed; partial tile: the guarded loop, ~1.2x per slab; + the epilogue / ring turn-around
                const float cost = (slabs == NI ? (float)NI : 1.2f * (float)slabs) + 0.3f;
                items.push_back({(int32_t)(g.item0 + (int64_t)et * g.n_qt + qt), cost});
            }
    }
    std::stable_sort(items.begin(), items.end(), [](const Item &a, const Item &b) { return a.cost > b.cost; });
    std::vector<std::vector<int32_t>> lists((size_t)grid);
    std::vector<std::pair<float, int>> heap;      // (load, cta) min-heap
    heap.reserve((size_t)grid);
    for (int c = 0; c < grid; c++) heap.push_back({0.f, c});
    auto cmp = [](const std::pair<float, int> &a, const std::pair<float, int> &b) { return a.first > b.first || (a.first == b.first && a.second > b.second); };
    std::make_heap(heap.begin(), heap.end(), cmp);
    for (const Item &it : items) {
        std::pop_heap(heap.begin(), heap.end(), cmp);
        auto &top = heap.back();
        lists[(size_t)top.second].push_back(it.id);
        top.first += it.cost;
        std::push_heap(heap.begin(), heap.end(), cmp);
    }
    size_t rounds = 0;
    for (auto &l : lists) rounds = std::max(rounds, l.size());
    std::vector<int32_t> sched(rounds * (size_t)grid, -1);
    for (int c = 0; c < grid; c++)
        for (size_t k = 0; k < lists[(size_t)c].size(); k++) sched[k * (size_t)grid + (size_t)c] = lists[(size_t)c][k];
    MRE_TRY(ctx->sched.reserve(std::max<size_t>(sched.size(), 1) * sizeof(int32_t)));
    // pageable source: staged by the runtime before cudaMemcpyAsync returns
    MRE_CUDA(cudaMemcpyAsync(ctx->sched.p, sched.data(), sched.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    p.sched = ctx->sched.as<int32_t>();
    p.sched_rounds = (int32_t)rounds;
    return MRE_OK;
}

// normalised / zero-padded copies of the tables when the job asks for them (else the caller's tables are read in place)
static int transe_tables(mre_ctx *ctx, const mre_rank_job *job, cudaStream_t st, const float **ent_out, const float **rel_out,
                         int64_t *Dp_out) {
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
    *ent_out = ent;
    *rel_out = rel;
    *Dp_out = Dp;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

// per-query thresholds + the pair-interleaved query vectors + zeroed counters (p.ent, p.D, p.groups, p.total_slots, p.counts set)
template <int P>
static int transe_queries(mre_ctx *ctx, const RankParams &p, const float *rel, cudaStream_t st) {
    const size_t qbytes = (size_t)std::max<int64_t>(p.total_slots, 2) * p.D * sizeof(float);
    MRE_TRY(ctx->qvec.reserve(qbytes));
    MRE_TRY(ctx->thr.reserve((size_t)p.Q * sizeof(float2)));
    if (p.total_slots != p.Q) MRE_CUDA(cudaMemsetAsync(ctx->qvec.p, 0, qbytes, st));   // the odd halves no query owns
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((p.Q + TQ_G * TQ_WARPS - 1) / (TQ_G * TQ_WARPS), (int64_t)ctx->sm_count * 8));
    transe_query_kernel<P><<<grid, TQ_WARPS * 32, 0, st>>>(p, rel, ctx->qvec.as<float>(), ctx->thr.as<float2>());
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

static int transe_grid(const mre_ctx *ctx, int64_t total_items) {
    const int per_sm = ctx->opt_transe_ctas > 0 ? ctx->opt_transe_ctas : CTAS_PER_SM;
    return (int)std::max<int64_t>(1, std::min<int64_t>(total_items, (int64_t)ctx->sm_count * per_sm));
}

template <int P, bool NEED_EQ>
static int launch_rank(mre_ctx *ctx, const RankParams &p, const CUtensorMap &tm_q, const CUtensorMap &tm_e, cudaStream_t st) {
    auto kern = transe_rank_kernel<P, NEED_EQ>;
    MRE_TRY(ctx->allow_smem(reinterpret_cast<const void *>(kern), RANK_SMEM));     // per device: function attributes are
    const int grid = transe_grid(ctx, p.total_items);
    kern<<<grid, RANK_THREADS, RANK_SMEM, st>>>(p, tm_q, tm_e);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

int rank_transe(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, cudaStream_t st) {
    MRE_CHECK_ARG(job->p_norm == 1 || job->p_norm == 2, "p_norm must be 1 or 2");
    const float *ent = nullptr, *rel = nullptr;
    int64_t Dp = 0;
    MRE_TRY(transe_tables(ctx, job, st, &ent, &rel, &Dp));
    RankParams p{};
    std::vector<GroupDesc> groups;
    MRE_TRY(fill_rank_params(ctx, ix, job, TQ, TILE_E, st, p, &groups));
    p.ent = ent;
    p.D = Dp;
    if (job->Q == 0) return MRE_OK;
    const int rank_grid = transe_grid(ctx, p.total_items);
    // (jobs of many waves lose nothing to the last one and would need a long list: round-robin there)
    if (ctx->opt_transe_lpt && p.total_items > rank_grid && p.total_items <= 64LL * rank_grid)
        MRE_TRY(build_item_schedule(ctx, groups, p.total_items, rank_grid, st, p));
    // Known-true correction, part 1 (MRE_FILTER_INDEX): every run the job touches is scored once.  It needs the tables and the
    // index only, so it runs on the context's second stream beside the gather / query kernels and is joined before part 2.
    KnownRuns kr{};
    const unsigned qgrid = (unsigned)((p.Q + KNOWN_WARPS - 1) / KNOWN_WARPS);
    if (p.filter == MRE_FILTER_INDEX) {
        MRE_TRY(known_runs_scratch(ctx, p, kr));
        cudaStream_t aux = nullptr;
        MRE_TRY(ctx->fork_aux(st, &aux));
        if (job->p_norm == 1) transe_known_score_kernel<1><<<qgrid, KNOWN_WARPS * 32, 0, aux>>>(p, rel, kr);
        else transe_known_score_kernel<2><<<qgrid, KNOWN_WARPS * 32, 0, aux>>>(p, rel, kr);
        ctx->launches += 1;
        MRE_CUDA(cudaGetLastError());
    }
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
    if (job->p_norm == 1) MRE_TRY(transe_queries<1>(ctx, p, rel, st));
    else MRE_TRY(transe_queries<2>(ctx, p, rel, st));
    p.qvec = ctx->qvec.as<float>();
    p.thr = ctx->thr.as<float2>();
    // part 2: take the known-true entities out of the filtered counters (needs the thresholds and the zeroed counters)
    if (p.filter == MRE_FILTER_INDEX) {       // compare every query's thresholds against its run's stored scores
        MRE_TRY(ctx->join_aux(st));
        if (job->p_norm == 1) transe_known_compare_kernel<1><<<qgrid, KNOWN_WARPS * 32, 0, st>>>(p, kr);
        else transe_known_compare_kernel<2><<<qgrid, KNOWN_WARPS * 32, 0, st>>>(p, kr);
    } else if (known_is_flat(p)) {            // list prefix known on the host: one (query, entry) pair per lane
        const unsigned kgrid = (unsigned)((p.filt_nnz + p.Q + KNOWN_WARPS * 32 - 1) / (KNOWN_WARPS * 32));
        if (job->p_norm == 1) transe_known_flat_kernel<1><<<kgrid, KNOWN_WARPS * 32, 0, st>>>(p);
        else transe_known_flat_kernel<2><<<kgrid, KNOWN_WARPS * 32, 0, st>>>(p);
    } else {                                  // one warp walks each query's list
        if (job->p_norm == 1) transe_known_kernel<1><<<qgrid, KNOWN_WARPS * 32, 0, st>>>(p, rel);
        else transe_known_kernel<2><<<qgrid, KNOWN_WARPS * 32, 0, st>>>(p, rel);
    }
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    CUtensorMap tm_q, tm_e;
    // query vectors: [slots / 2 pair-rows][2 Dp floats], box = 64 pair-rows x 32 floats (16 d-values of two queries)
    MRE_TRY(make_tmap_f32_2d(&tm_q, p.qvec, std::max<int64_t>(p.total_slots / 2, 1), 2 * Dp, 2 * Dp, TQ / 2, CHUNK));
    MRE_TRY(make_tmap_f32_2d(&tm_e, cand_table, std::max<int64_t>(cand_rows, 1), Dp, Dp, TILE_E, CHUNK));
    MRE_TRY(ctx->time_begin(st));
    // the tie count is always on: one extra compare per score, in the epilogue only
    if (job->p_norm == 1) MRE_TRY((launch_rank<1, true>(ctx, p, tm_q, tm_e, st)));
    else MRE_TRY((launch_rank<2, true>(ctx, p, tm_q, tm_e, st)));
    MRE_TRY(ctx->time_end(st));
    return MRE_OK;
}

int predict_transe(mre_ctx *ctx, const mre_rank_job *job, int64_t query, float *scores_out, cudaStream_t st) {
    const float *ent = nullptr, *rel = nullptr;
    int64_t Dp = 0;
    MRE_TRY(transe_tables(ctx, job, st, &ent, &rel, &Dp));
    const uint8_t *qs = job->q_side ? job->q_side + query : nullptr;
    unsigned grid = (unsigned)((job->E + 127) / 128);
    if (job->p_norm == 1)
        transe_predict_kernel<1><<<grid, 128, 0, st>>>(ent, rel, job->E, Dp, job->q_h + query, job->q_t + query, job->q_r + query, qs,
                                                       job->side, scores_out);
    else
        transe_predict_kernel<2><<<grid, 128, 0, st>>>(ent, rel, job->E, Dp, job->q_h + query, job->q_t + query, job->q_r + query, qs,
                                                       job->side, scores_out);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

}  // namespace mre
