// DistMult / ComplEx fused contraction + rank on the 5th-generation tensor cores (tcgen05 + TMEM) for sm_100a.
//
// Reference path replaced (paths relative to /root/reference):
//   OpenKE/openke/module/model/DistMult.py:34-44,70-72   score = sum_d h*r*t ; predict = -score
//   OpenKE/openke/module/model/ComplEx.py:20-27,60-61    Re<h, r, conj t> over four real tables ; predict = -score
//   module/NegativeSampling.py:158-168                   the paper's distmult branch
//   OpenKE/openke/base/Test.h:65-192                     testHead / testTail compare loop + _find
//
// The 1-vs-all score of a query is one row of  (query vector)[Q, K] x (entity table)[E, K]^T :
//   DistMult  K = D  : tail query v = h o r ; head query v = r o t
//   ComplEx   K = 2D : entity row = [e_re | e_im]; tail v = [h_re r_re - h_im r_im | h_im r_re + h_re r_im];
//                      head v = [t_re r_re + t_im r_im | t_im r_re - t_re r_im]
// the one true dense contraction of the hot path, so it runs as a tcgen05 GEMM whose epilogue never writes scores:
// accumulator tiles (128 queries x 256 entities, FP32) live in TMEM, double-buffered; four epilogue warps read them
// back with tcgen05.ld (one query row per thread), compare against the query's true score and count (raw and filtered
// counters alike; bil_known_kernel scores each query's few known-true entities exactly and takes them back out of the
// filtered ones); columns closer to the true score than an error guard are re-scored in scalar FP32, which makes the COUNTS
// those of the sequential FP32 scorer.
// Precision: the reference is FP32, and the COUNTS must be those of the FP32 scorer.  The tensor cores only have to decide
// every column that is not a near-tie, so every operand is split x ~= hi + lo into two BF16 values (hi = rn(x),
// lo = rn(x - hi): 16 significant bits) and each product is issued as THREE kind::f16 BF16 MMAs  hi*hi + lo*hi + hi*lo
// with FP32 accumulation: relative error <= ~2^-16 of sum|v_d e_d|, at half the tensor-pipe time and half the operand
// bytes of the 3xTF32 form this kernel used first (that form ran into the L2 -> SM feed limit, ~42 B/clk/SM, at 71 % of
// the TF32 pipe).  A rigorous error guard then routes the (rare) columns within the guard of s_true to the exact scalar
// re-score.  Algorithmic flops are counted once (2*Q*E*K); the tensor pipe executes 3x that.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected thread),
// warps 2..9 = epilogue (TMEM lane quarter = warp % 4, two warps per quarter: one per 128-column half of the tile).  Shared memory: 2 stages x {A_hi, A_lo: 128 x 128 B;
// B_hi, B_lo: 256 x 128 B} (64 BF16 of K per row), 128-byte swizzle, K-major, fed by TMA tensor tiles.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.h"
#include "rank_common.cuh"
#include "rank_host.h"
#include "tma_host.h"

#ifndef MRE_DIAG_BIL_NOEPI
#define MRE_DIAG_BIL_NOEPI 0
#endif

namespace mre {

constexpr int BN = 256;                 // entities per tile (UMMA N)
constexpr int BM = 128;                 // queries per tile (UMMA M)
#ifndef MRE_BIL_BK
#define MRE_BIL_BK 64
#endif
constexpr int BK = MRE_BIL_BK;          // BF16 elements of K per stage: 64 = one 128-byte swizzle atom per row, 32 = a 64-byte atom (more, smaller stages)
constexpr int UK = 16;                  // elements of K per tcgen05.mma kind::f16
constexpr int TF_BK = 32, TF_UK = 8;    // the same for the kind::tf32 peak probe
constexpr uint32_t A_BYTES = BM * BK * 2;   // 16 KiB
constexpr uint32_t B_BYTES = BN * BK * 2;   // 32 KiB
// one CTA per tile: a stage holds A_hi, A_lo (128 rows) and B_hi, B_lo (256 rows) = 96 KiB, 2 stages.
// CTA pair (cta_group::2): M = 256 over two CTAs, each CTA stages its own 128 query rows and HALF of the candidate tile
// (128 rows) -- 64 KiB per stage, 3 stages -- and the pair's MMA reads both halves: a third less L2 -> SM traffic per flop.
// NPROD = 3: BF16 hi / lo operands, three MMAs per product.  NPROD = 1: ONE FP16 operand per side and one MMA per product --
// half the bytes per stage, so twice the stages in the same 192 KiB.
template <bool PAIR, int NPROD> struct StageCfg {
    static constexpr int STAGES = (PAIR ? 3 : 2) * (64 / BK) * (NPROD == 1 ? 2 : 1);
    static constexpr uint32_t B_HALF = PAIR ? B_BYTES / 2 : B_BYTES;
    static constexpr uint32_t BYTES = (NPROD == 1 ? 1 : 2) * (A_BYTES + B_HALF);
    static constexpr uint32_t B_OFF = (NPROD == 1 ? 1 : 2) * A_BYTES;      // first B operand inside a stage
};
constexpr int MAX_STAGES = 6 * (64 / BK);
__device__ __forceinline__ uint64_t umma_desc_op(uint32_t addr) { return BK == 64 ? umma_desc_k128(addr) : umma_desc_k64(addr); }
#ifndef MRE_EPI_WARPS
#define MRE_EPI_WARPS 4
#endif
constexpr int EPI_WARPS = MRE_EPI_WARPS;     // 4: one per TMEM lane quarter; 8: two per quarter, one per 128-column half (measured slower)
constexpr int RESCORE_WARPS = 4;             // one per epilogue warp of the first column slice
constexpr int BIL_THREADS = (2 + EPI_WARPS + RESCORE_WARPS) * 32;
constexpr int EPI_WARP0 = 2;
constexpr int PEND_CAP = 512;                // per epilogue warp: ring of near-ties handed to its re-score warp through shared memory
constexpr size_t BIL_SMEM = 1024 + (size_t)2 * (64 / BK) * (2 * A_BYTES + 2 * B_BYTES) + 48 * sizeof(uint64_t) + RESCORE_WARPS * (PEND_CAP * sizeof(uint2) + 16) + 64;   // both configurations: 192 KiB of stages
constexpr uint32_t TMEM_COLS = 512;     // two 256-column accumulator buffers

struct BilParams {
    RankParams r;            // r.ent = full-precision [E, K] table (scalar scorer), r.qvec = full-precision query vectors
    const float *delta;      // [Q] near-tie guard: |s_mma - s_true| <= delta => the column is re-scored in scalar FP32
    int64_t k8;              // K padded to a multiple of 8: row pitch (elements) of the BF16 hi / lo tables
    unsigned long long *rescored;   // running total of exact re-scores (mre_ctx_stat "bil_rescored")
    float *store;            // STORE mode only: [Q, store_ld] tensor-core similarities are written instead of counted
    int64_t store_ld;
};

// ------------------------------------------------------------------------------------------ pre-pass kernels
// x ~= hi + lo with hi = rn_bf16(x), lo = rn_bf16(x - hi)
__device__ __forceinline__ void bf16_split(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// entity table -> tensor-core operands [rows, K8] (ComplEx: [re | im]; zero padded) + optional full-precision copy [rows, Kp]
// + the largest row norm (atomicMax on (generation << 32 | bit pattern): positive floats order like their bits), ONE pass, one warp per row.
// NPROD = 3: BF16 hi / lo split.  NPROD = 1: hi = rn_fp16(x), lo untouched.
template <int NPROD>
__global__ void __launch_bounds__(256) bil_table_kernel(const float *__restrict__ re, const float *__restrict__ im, int64_t rows, int64_t D,
                                                        int64_t K, int64_t Kp, int64_t K8, float *__restrict__ full,
                                                        uint16_t *__restrict__ hi, uint16_t *__restrict__ lo,
                                                        unsigned long long *__restrict__ max_norm, unsigned int epoch) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    float best = 0.f;
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
        float ss = 0.f;
        for (int64_t d = lane; d < K8; d += 32) {
            float x = 0.f;
            if (d < K) x = d < D ? re[row * D + d] : im[row * D + (d - D)];
            if (full && d < Kp) full[row * Kp + d] = x;
            ss = fmaf(x, x, ss);
            if (NPROD == 3) {
                __nv_bfloat16 h, l;
                bf16_split(x, h, l);
                hi[row * K8 + d] = __bfloat16_as_ushort(h);
                lo[row * K8 + d] = __bfloat16_as_ushort(l);
            } else {
                hi[row * K8 + d] = __half_as_ushort(__float2half_rn(x));
            }
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);
        best = fmaxf(best, ss);
    }
    // the slot is tagged with the call's generation in its high word: a newer call always wins, so nothing resets it
    if (lane == 0 && best > 0.f)
        atomicMax(max_norm, ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(sqrtf(best) * 1.0001f));
}

// Everything per query in ONE launch: the query vector (full precision [Q, Kp] + tensor-core operands [Q, K8]), the threshold
// pair on the predict scale (p = -sim: lower is better), the near-tie guard, and the zeroing of the query's four counters.
// s_true must carry the roundings and the ORDER of the exact re-score (bil_dot: acc = acc + v_d * e_d, d ascending), and a
// sequential sum by one lane of a warp costs a whole issue slot per instruction.  So a warp takes 8 queries at a time: for
// every 128-wide chunk of d the lanes first work ACROSS d (query by query: coalesced row reads, the vector and its operands
// written out, the rounded products v_d * e_true,d parked in shared memory), then ALONG d (lane l < 8 adds query l's products
// to its accumulator in order) -- 8 sequential sums advance side by side, bit-identical to the scalar scorer.
// Guard (rigorous, relative to sum_d |v_d e_d| <= ||v|| * max_j ||e_j||): the tensor-core value differs from the sequential
// FP32 value by
//   NPROD = 3  the split: x - hi - lo <= 2^-18 |x| per operand and the dropped lo*lo <= 2^-18  ->  <= 3 * 2^-18 = 1.15e-5
//   NPROD = 1  one FP16 rounding per operand: |x - rn(x)| <= 2^-11 |x| + 2^-25 (subnormals) -> (2^-10 + 2^-22) of the scale
//              plus 2^-25 (||v||_1 + ||e||_1) <= 2^-25 sqrt(K) (||v|| + max||e||); operands beyond the FP16 range make the
//              guard infinite (every column of that query is re-scored exactly)
//   both       the FP32 accumulation of the K/16 (x3) MMAs and the scalar scorer's own K roundings  ->  < K * 1.2e-7
// Every column closer than the guard to s_true is re-scored with the scalar scorer, so the COUNTS are exactly those of the
// FP32 scorer for ANY table.
constexpr int QK_WARPS = 8;      // warps per block
constexpr int QK_G = 8;          // queries a warp works on together (lanes 0..7 run their sequential sums side by side)
constexpr int QK_CH = 64;        // elements of d per round: two per lane; a round requests all of its row elements up front
template <int NPROD>
__global__ void __launch_bounds__(QK_WARPS * 32) bil_query_kernel(int scorer, const float *__restrict__ ent, const float *__restrict__ ent_im,
        const float *__restrict__ rel, const float *__restrict__ rel_im, const float *__restrict__ ent_full, int64_t D, int64_t K, int64_t Kp,
        int64_t K8, const int64_t *__restrict__ q_h, const int64_t *__restrict__ q_t, const int64_t *__restrict__ q_r,
        const uint8_t *__restrict__ q_side, int side, int64_t Q, const unsigned long long *__restrict__ max_norm, unsigned int epoch,
        float *__restrict__ qv, uint16_t *__restrict__ qhi, uint16_t *__restrict__ qlo, float2 *__restrict__ thr, float *__restrict__ delta,
        int32_t *__restrict__ counts) {
    __shared__ float sP[QK_WARPS][QK_G][QK_CH + 1];     // rounded products v_d * e_true,d; +1: conflict-free both ways
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_groups = (Q + QK_G - 1) / QK_G;
    for (int64_t grp = (int64_t)blockIdx.x * QK_WARPS + warp; grp < n_groups; grp += (int64_t)gridDim.x * QK_WARPS) {
        const int64_t q0 = grp * QK_G;
        const int nq = (int)min((int64_t)QK_G, Q - q0);
        // lane l < nq holds the ids of query q0 + l
        int my_s = 0, my_fixed = 0, my_truth = 0, my_r = 0;
        if (lane < nq) {
            const int64_t q = q0 + lane;
            my_s = q_side ? (int)q_side[q] : side;
            my_fixed = (int)(my_s ? q_h[q] : q_t[q]);          // the entity that stays fixed in the query
            my_truth = (int)(my_s ? q_t[q] : q_h[q]);
            my_r = (int)q_r[q];
        }
        float acc = 0.f;                          // lane l: the sequential sum of query q0 + l
        float ss[QK_G], vmax[QK_G];               // per-lane partials of ||v||^2 and max |v| of the group's queries
#pragma unroll
        for (int qi = 0; qi < QK_G; qi++) { ss[qi] = 0.f; vmax[qi] = 0.f; }
        for (int64_t c0 = 0; c0 < K8; c0 += QK_CH) {
            // every row element this chunk needs, for all the group's queries, is requested BEFORE anything is used: one memory
            // round trip per chunk instead of one per query (the pre-pass runs on cold caches: the chain of trips is its cost)
            constexpr int J = QK_CH / 32;
            float x0[QK_G][J], x1[QK_G][J], x2[QK_G][J], x3[QK_G][J], xt[QK_G][J];
            int sq[QK_G];
            const int dbase = (int)c0 + lane, Di = (int)D, Ki = (int)K, Kpi = (int)Kp, K8i = (int)K8;     // K8 < 2^31: 32-bit column math
#pragma unroll
            for (int qi = 0; qi < QK_G; qi++) {
                sq[qi] = __shfl_sync(0xffffffffu, my_s, qi);
                const int64_t e = __shfl_sync(0xffffffffu, my_fixed, qi), truth = __shfl_sync(0xffffffffu, my_truth, qi);
                const int64_t r = __shfl_sync(0xffffffffu, my_r, qi);
                const float *pe = ent + e * D, *pr = rel + r * D, *pt = ent_full + truth * Kp;        // row bases: one 64-bit multiply each
                const float *pei = scorer != MRE_DISTMULT ? ent_im + e * D : pe, *pri = scorer != MRE_DISTMULT ? rel_im + r * D : pr;
#pragma unroll
                for (int j = 0; j < J; j++) {
                    const int d = dbase + 32 * j;
                    const bool live = qi < nq && d < Ki;
                    const int dd = d < Di ? d : d - Di;
                    x0[qi][j] = live ? pe[dd] : 0.f;
                    x1[qi][j] = live ? pr[dd] : 0.f;
                    if (scorer != MRE_DISTMULT) {
                        x2[qi][j] = live ? pei[dd] : 0.f;
                        x3[qi][j] = live ? pri[dd] : 0.f;
                    }
                    xt[qi][j] = (qi < nq && d < Kpi) ? pt[d] : 0.f;
                }
            }
#pragma unroll
            for (int qi = 0; qi < QK_G; qi++) {
                if (qi >= nq) break;                                            // warp-uniform
                const int s = sq[qi];
                const int64_t q = q0 + qi;
                float *oq = qv + q * Kp;
                uint16_t *ohi = qhi + q * K8, *olo = qlo + q * K8;
#pragma unroll
                for (int j = 0; j < J; j++) {
                    const int d = dbase + 32 * j;
                    float v = 0.f;
                    if (d < Ki) {
                        if (scorer == MRE_DISTMULT) {
                            v = s ? x0[qi][j] * x1[qi][j] : x1[qi][j] * x0[qi][j];
                        } else {
                            const float ere = x0[qi][j], eim = x2[qi][j], rre = x1[qi][j], rim = x3[qi][j];
                            if (s) v = d < Di ? ere * rre - eim * rim : eim * rre + ere * rim;
                            else v = d < Di ? ere * rre + eim * rim : eim * rre - ere * rim;
                        }
                    }
                    float prod = 0.f;
                    if (d < Kpi) {
                        oq[d] = v;
                        prod = v * xt[qi][j];
                    }
                    if (d < K8i) {
                        if (NPROD == 3) {
                            __nv_bfloat16 h, l;
                            bf16_split(v, h, l);
                            ohi[d] = __bfloat16_as_ushort(h);
                            olo[d] = __bfloat16_as_ushort(l);
                        } else {
                            ohi[d] = __half_as_ushort(__float2half_rn(v));
                        }
                    }
                    sP[warp][qi][lane + 32 * j] = prod;
                    ss[qi] = fmaf(v, v, ss[qi]);
                    vmax[qi] = fmaxf(vmax[qi], fabsf(v));
                }
            }
            __syncwarp();
            if (lane < nq) {                       // lane l adds query l's products in order: the scalar scorer's sum, 8 at a time
                const int nd = (int)min((int64_t)QK_CH, Kp - c0);
#pragma unroll 8
                for (int dd = 0; dd < nd; dd++) acc = acc + sP[warp][lane][dd];
            }
            __syncwarp();
        }
        // ||v||^2 and max |v| of every query of the group (any order will do: they only size the guard), handed to the query's lane
        float my_ss = 0.f, my_vmax = 0.f;
#pragma unroll
        for (int qi = 0; qi < QK_G; qi++) {
            float a = ss[qi], m = vmax[qi];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            }
            if (lane == qi) { my_ss = a; my_vmax = m; }
        }
        if (lane < nq) {
            const int64_t q = q0 + lane;
            const float pt = -acc;
            float hi = pt;
            if (pt == pt && fabsf(pt) < INFINITY) hi = nextafterf(pt, INFINITY);
            thr[q] = make_float2(pt, hi);
            const unsigned long long mn = *max_norm;
            const float vn = sqrtf(my_ss), en = (unsigned int)(mn >> 32) == epoch ? __uint_as_float((unsigned int)mn) : 0.f;
            float g;
            if (NPROD == 3) {
                g = (1.3e-5f + 1.2e-7f * (float)Kp) * vn * en;
            } else {
                g = (9.77e-4f + 1.2e-7f * (float)Kp) * vn * en + 3.1e-8f * sqrtf((float)Kp) * (vn + en);
                if (!(my_vmax < 6.0e4f) || !(en < 6.0e4f)) g = INFINITY;
            }
            delta[q] = g;
#pragma unroll
            for (int c = 0; c < 4; c++) counts[(int64_t)c * Q + q] = 0;
        }
    }
}

// sequential FP32 dot product: the value Model.predict's `-sum(...)` negates (mul and add are separate roundings)
// K is a multiple of 4 and both rows are 16-byte aligned: 128-bit loads, several in flight, same summation order.
__device__ __forceinline__ float bil_dot(const float *__restrict__ v, const float *__restrict__ e, int64_t K) {
    float acc = 0.f;
    const float4 *v4 = reinterpret_cast<const float4 *>(v), *e4 = reinterpret_cast<const float4 *>(e);
    const int n4 = (int)(K >> 2);
#pragma unroll 8
    for (int d = 0; d < n4; d++) {
        const float4 a = __ldg(v4 + d), b = __ldg(e4 + d);
        acc = acc + a.x * b.x;
        acc = acc + a.y * b.y;
        acc = acc + a.z * b.z;
        acc = acc + a.w * b.w;
    }
    return acc;
}

// The filtered counts' correction (rank_common.cuh: known_correction) with the bilinear element: term = v_d * e_d (its own
// rounding), fold acc + term -- bil_dot's steps exactly, the arithmetic the tile kernel's decisions are guaranteed to agree with
// (sign outside the guard, exact re-score inside).  p.qvec / p.ent are the full-precision [.., Kp] tables, p.D = Kp.
struct BilKnownOp {
    const float *ent;
    const float *qvec, *v;
    const float2 *thr;
    int64_t D;
    float st;
    __device__ __forceinline__ void query(int64_t q, int, int64_t, int64_t, int64_t) {
        v = qvec + q * D;
    }
    __device__ __forceinline__ float vec(int d) const { return __ldg(v + d); }
    static constexpr int VSTRIDE = 1;
    __device__ __forceinline__ const float *flat_query(int64_t q, int64_t, int, int64_t, int64_t, int64_t) {
        st = -thr[q].x;
        return qvec + q * D;
    }
    static constexpr bool DIRECT_ONLY = false;
    __device__ __forceinline__ float direct(int64_t x) const { return bil_dot(v, ent + x * D, D); }
    __device__ __forceinline__ void thresholds(int64_t q) { st = -thr[q].x; }
    __device__ __forceinline__ bool truth_ties() const { return st == st; }
    __device__ __forceinline__ float term(float a, float e) const { return a * e; }
    __device__ __forceinline__ float fold(float acc, float t) const { return acc + t; }
    __device__ __forceinline__ void classify(float acc, int &lt, int &eq) const {
        if (acc > st) lt++;
        else if (acc == st) eq++;
    }
};
__global__ void __launch_bounds__(KNOWN_WARPS * 32) bil_known_kernel(const RankParams p) {
    __shared__ float sT[KNOWN_WARPS][32][33];
    __shared__ int64_t sX[KNOWN_WARPS][32];
    BilKnownOp op{p.ent, p.qvec, nullptr, p.thr, p.D, 0.f};
    known_correction(p, op, sT[threadIdx.x >> 5], sX[threadIdx.x >> 5]);
}
// (Building the query vector from the embedding rows instead, so that this kernel could run beside bil_query_kernel on the second
// stream, was measured twice and does not pay: with a scalar short-run path the ComplEx step went 0.61 -> 0.85 ms, with the
// vector staged in shared memory 0.607 -> 0.616 ms -- the two kernels compete for the same SMs and L2 misses.)
#ifndef MRE_BIL_KNOWN_MINB
#define MRE_BIL_KNOWN_MINB 3
#endif
__global__ void __launch_bounds__(KNOWN_WARPS * 32, MRE_BIL_KNOWN_MINB) bil_known_score_kernel(const RankParams p, const KnownRuns kr) {
    __shared__ float sT[KNOWN_WARPS][32][33];
    __shared__ int64_t sX[KNOWN_WARPS][32];
    BilKnownOp op{p.ent, p.qvec, nullptr, p.thr, p.D, 0.f};
    known_score_runs(p, kr, op, sT[threadIdx.x >> 5], sX[threadIdx.x >> 5]);
}
__global__ void __launch_bounds__(KNOWN_WARPS * 32) bil_known_compare_kernel(const RankParams p, const KnownRuns kr) {
    BilKnownOp op{p.ent, p.qvec, nullptr, p.thr, p.D, 0.f};
    known_compare_runs(p, kr, op);
}
__global__ void __launch_bounds__(KNOWN_WARPS * 32) bil_known_flat_kernel(const RankParams p) {
    __shared__ float sT[KNOWN_WARPS][32][33];
    __shared__ int64_t sX[KNOWN_WARPS][32];
    __shared__ const float *sV[KNOWN_WARPS][32];
    BilKnownOp op{p.ent, p.qvec, nullptr, p.thr, p.D, 0.f};
    known_correction_flat(p, op, sT[threadIdx.x >> 5], sX[threadIdx.x >> 5], sV[threadIdx.x >> 5]);
}

__global__ void bil_predict_kernel(const float *__restrict__ ent, int64_t E, int64_t K, const float *__restrict__ qv,
                                   float *__restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < E) out[j] = -bil_dot(qv, ent + j * K, K);
}

// ------------------------------------------------------------------------------------------ main kernel
// pair work item -> (group, query-tile PAIR, candidate tile); pairs vary fastest (as decode_item does with query tiles)
__device__ __forceinline__ void decode_pitem(const RankParams &p, int64_t item, int &g, int &qp, int &et) {
    int lo = 0, hi = p.n_groups;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (p.groups[mid].pitem0 <= item) lo = mid; else hi = mid;
    }
    g = lo;
    const int n_qp = (p.groups[g].n_qt + 1) >> 1;
    const int64_t local = item - p.groups[g].pitem0;
    qp = (int)(local % n_qp);
    et = (int)(local / n_qp);
}

template <bool STORE, bool PAIR, int NPROD>
__global__ void __launch_bounds__(BIL_THREADS, 1)
bilinear_rank_kernel(const BilParams bp, const __grid_constant__ CUtensorMap tm_ahi, const __grid_constant__ CUtensorMap tm_alo,
                     const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo) {
    using Cfg = StageCfg<PAIR, NPROD>;
    constexpr int B_STAGES = Cfg::STAGES;
    constexpr uint32_t BSTAGE_BYTES = Cfg::BYTES, B_HALF = Cfg::B_HALF, B_OFF = Cfg::B_OFF;
    const RankParams &p = bp.r;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t ring_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *ring = smem_raw + (ring_u32 - smem_u32(smem_raw));
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring + (size_t)B_STAGES * BSTAGE_BYTES);
    // bars: full[3], empty[3], tmem_full[2], tmem_empty[2]; then the TMEM base address word
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + MAX_STAGES), tfull0 = smem_u32(bars + 2 * MAX_STAGES),
                   tempty0 = smem_u32(bars + 2 * MAX_STAGES + 2);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_STAGES + 4);
    uint2 *pend_all = reinterpret_cast<uint2 *>(bars + 48);   // per epilogue warp: [PEND_CAP] ring of near-ties
    volatile uint32_t *pend_ctl = reinterpret_cast<volatile uint32_t *>(pend_all + RESCORE_WARPS * PEND_CAP);   // per ring: published, consumed, done
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_kb = (int)((bp.k8 + BK - 1) / BK);
    // work distribution: a CTA walks tiles on its own, a CTA pair walks two-query-tile items together
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int64_t worker = PAIR ? (blockIdx.x >> 1) : blockIdx.x, n_workers = PAIR ? (gridDim.x >> 1) : gridDim.x;
    const int64_t n_items = PAIR ? p.total_pitems : p.total_items;

    if (threadIdx.x == 0) {
        for (int s = 0; s < B_STAGES; s++) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int k = 0; k < RESCORE_WARPS * 4; k++) pend_ctl[k] = 0u;
        for (int s = 0; s < 2; s++) {
            mbar_init(tfull0 + 8 * s, 1);
            mbar_init(tempty0 + 8 * s, PAIR ? 2 * EPI_WARPS : EPI_WARPS);     // the pair's leader collects the epilogue warps of both CTAs
        }
        fence_barrier_init();
        tma_prefetch_desc(&tm_ahi); tma_prefetch_desc(&tm_alo); tma_prefetch_desc(&tm_bhi); tma_prefetch_desc(&tm_blo);
    }
    if (warp == 1) {
        if (PAIR) tmem_alloc_pair(smem_u32(tmem_slot), TMEM_COLS);
        else tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();      // both CTAs' barriers are initialised before either signals the other's
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================= TMA producer =================================================
        if (lane == 0) {
            uint32_t it = 0;
            const uint32_t full_leader0 = PAIR ? mapa_shared(full0, 0) : full0;   // TMA bytes of both CTAs are counted by the leader
            for (int64_t item = worker; item < n_items; item += n_workers) {
                int g, qt, et;
                if (PAIR) {
                    decode_pitem(p, item, g, qt, et);
                    qt = 2 * qt + (int)rank;
                } else {
                    decode_item(p, item, g, qt, et);
                }
                const GroupDesc &gd = p.groups[g];
                const int qrow = (int)(gd.q0 + (int64_t)qt * BM);
                const int erow = (int)(gd.c0 + (int64_t)et * BN) + (PAIR ? (int)rank * (BN / 2) : 0);
                for (int kb = 0; kb < n_kb; kb++, it++) {
                    const int s = it % B_STAGES;
                    mbar_wait(empty0 + 8 * s, ((it / B_STAGES) & 1) ^ 1);
                    const uint32_t base = ring_u32 + (uint32_t)s * BSTAGE_BYTES;
                    // query tiles are re-read for every candidate tile of the sweep: keep them in L2 (evict-last) so the
                    // streaming candidate tiles (each shared by the CTAs of one round, then dead) cannot push them out
                    if (PAIR) {
                        const uint32_t full = full_leader0 + 8 * s;
                        if (rank == 0) mbar_arrive_expect_tx(full0 + 8 * s, 2 * BSTAGE_BYTES);
                        tma_load_2d_pair(base, &tm_ahi, kb * BK, qrow, full, L2_EVICT_LAST);
                        if (NPROD == 3) tma_load_2d_pair(base + A_BYTES, &tm_alo, kb * BK, qrow, full, L2_EVICT_LAST);
                        tma_load_2d_pair(base + B_OFF, &tm_bhi, kb * BK, erow, full, L2_EVICT_NORMAL);
                        if (NPROD == 3) tma_load_2d_pair(base + B_OFF + B_HALF, &tm_blo, kb * BK, erow, full, L2_EVICT_NORMAL);
                    } else {
                        const uint32_t full = full0 + 8 * s;
                        mbar_arrive_expect_tx(full, BSTAGE_BYTES);
                        tma_load_2d_hint(base, &tm_ahi, kb * BK, qrow, full, L2_EVICT_LAST);
                        if (NPROD == 3) tma_load_2d_hint(base + A_BYTES, &tm_alo, kb * BK, qrow, full, L2_EVICT_LAST);
                        tma_load_2d_hint(base + B_OFF, &tm_bhi, kb * BK, erow, full, L2_EVICT_NORMAL);
                        if (NPROD == 3) tma_load_2d_hint(base + B_OFF + B_HALF, &tm_blo, kb * BK, erow, full, L2_EVICT_NORMAL);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================================= MMA issuer ===================================================
        if (lane == 0 && rank == 0) {      // in a pair only the leader CTA issues: one instruction drives both SMs' tensor cores
            constexpr uint32_t idesc = NPROD == 3 ? umma_idesc_bf16(PAIR ? 2 * BM : BM, BN) : umma_idesc_f16(PAIR ? 2 * BM : BM, BN);
            uint32_t it = 0, tile = 0;
            for (int64_t item = worker; item < n_items; item += n_workers, tile++) {
                const uint32_t buf = tile & 1;
                mbar_wait(tempty0 + 8 * buf, ((tile >> 1) & 1) ^ 1);   // epilogue has drained this accumulator buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BN;
                for (int kb = 0; kb < n_kb; kb++, it++) {
                    const int s = it % B_STAGES;
                    mbar_wait(full0 + 8 * s, (it / B_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t base = ring_u32 + (uint32_t)s * BSTAGE_BYTES;
                    const int n_ks = (int)min((int64_t)(BK / UK), (bp.k8 - (int64_t)kb * BK + UK - 1) / UK);
                    for (int ks = 0; ks < n_ks; ks++) {
                        const uint32_t koff = ks * UK * 2;   // bytes along K inside the 128-byte swizzle atom
                        const uint64_t ahi = umma_desc_op(base + koff), bhi = umma_desc_op(base + B_OFF + koff);
                        if (PAIR) umma_bf16_pair(d_tmem, ahi, bhi, idesc, (kb | ks) != 0);       // kind::f16: BF16 or FP16 per idesc
                        else umma_bf16(d_tmem, ahi, bhi, idesc, (kb | ks) != 0);
                        if (NPROD == 3) {
                            const uint64_t alo = umma_desc_op(base + A_BYTES + koff), blo = umma_desc_op(base + B_OFF + B_HALF + koff);
                            if (PAIR) {
                                umma_bf16_pair(d_tmem, alo, bhi, idesc, 1);
                                umma_bf16_pair(d_tmem, ahi, blo, idesc, 1);
                            } else {
                                umma_bf16(d_tmem, alo, bhi, idesc, 1);
                                umma_bf16(d_tmem, ahi, blo, idesc, 1);
                            }
                        }
                    }
                    // shared-memory stage reusable (in both CTAs of a pair) once these MMAs retire
                    if (PAIR) umma_commit_pair(empty0 + 8 * s, 3);
                    else umma_commit(empty0 + 8 * s);
                }
                // accumulator tile complete (each CTA of a pair holds its own 128 rows)
                if (PAIR) umma_commit_pair(tfull0 + 8 * buf, 3);
                else umma_commit(tfull0 + 8 * buf);
            }
        }
        __syncwarp();
    } else if (warp < EPI_WARP0 + EPI_WARPS) {
        // ================================================= epilogue warps ===============================================
        const int quarter = warp & 3;                   // TMEM lanes [32 * quarter, 32 * quarter + 32)
        const int chalf = (warp - EPI_WARP0) >> 2;      // which column slice of the accumulator tile this warp compares
        const int row = quarter * 32 + lane;            // query row of the tile owned by this thread
        // Near-ties (columns within the guard of s_true) are handed, as (query, entity), to this warp's RE-SCORE
        // warp through a small shared-memory ring; the exact scalar re-score (two row gathers per entry: pure latency) thus
        // never sits on the epilogue's critical path.  published / consumed are monotonic counters.
        static_assert(EPI_WARPS == RESCORE_WARPS, "one re-score warp per epilogue warp");
        const uint32_t tempty_leader0 = PAIR ? mapa_shared(tempty0, 0) : tempty0;
        const int ring_id = warp - EPI_WARP0;
        uint2 *pend = pend_all + ring_id * PEND_CAP;
        volatile uint32_t *ctl = pend_ctl + ring_id * 4;
        uint32_t n_pub = 0;                             // warp-uniform copy of ctl[0]
        auto rescore = [&](uint2 it2) {
            const int64_t q2 = it2.x, ent_id = it2.y;
            const float s2 = bil_dot(p.qvec + q2 * p.D, p.ent + ent_id * p.D, p.D);
            const float st = -__ldg(&p.thr[q2].x);
            if (s2 > st) { atomicAdd(p.counts + q2, 1); atomicAdd(p.counts + 2 * p.Q + q2, 1); }
            if (s2 == st) { atomicAdd(p.counts + p.Q + q2, 1); atomicAdd(p.counts + 3 * p.Q + q2, 1); }
        };
        // Per-tile metadata (tile geometry, this row's true score and guard) is
        // fetched ONE TILE AHEAD: a lone epilogue warp per SM sub-partition cannot hide dependent global-load latency,
        // and the epilogue of tile t must finish before the MMA warp may start tile t + 2.
        struct TileMeta {
            int64_t qbase, crow0;
            int nq, ne;
            float sim_true, guard;
        };
        const GroupDesc gd0 = p.groups[0];
        auto load_meta = [&](int64_t item) {
            TileMeta m;
            int g = 0, qt, et;
            GroupDesc gd = gd0;
            if (PAIR) {
                if (p.n_groups > 1) {
                    decode_pitem(p, item, g, qt, et);
                    gd = p.groups[g];
                } else {
                    const int n_qp = (gd0.n_qt + 1) >> 1;
                    qt = (int)(item % n_qp);
                    et = (int)(item / n_qp);
                }
                qt = 2 * qt + (int)rank;           // this CTA's query tile of the pair (past the group's end when n_qt is odd)
            } else if (p.n_groups > 1) {
                decode_item(p, item, g, qt, et);
                gd = p.groups[g];
            } else {
                qt = (int)(item % gd0.n_qt);
                et = (int)(item / gd0.n_qt);
            }
            m.qbase = gd.q0 + (int64_t)qt * BM;
            m.crow0 = gd.c0 + (int64_t)et * BN;
            m.nq = (int)min((int64_t)BM, gd.q0 + gd.nq - m.qbase);
            m.ne = (int)min((int64_t)BN, gd.nc - (int64_t)et * BN);
            m.sim_true = INFINITY;
            m.guard = -1.f;
            if (row < m.nq) {
                m.sim_true = -__ldg(&p.thr[m.qbase + row].x);
                m.guard = __ldg(bp.delta + m.qbase + row);
            }
            return m;
        };
        uint32_t tile = 0;
        TileMeta nxt{};
        if (worker < n_items) nxt = load_meta(worker);
        for (int64_t item = worker; item < n_items; item += n_workers, tile++) {
            const TileMeta cur = nxt;
            if (item + n_workers < n_items) nxt = load_meta(item + n_workers);
            const int64_t qbase = cur.qbase;
            const int nq = cur.nq, ne = cur.ne;
            const bool q_ok = row < nq;
            const float sim_true = cur.sim_true, guard = cur.guard;
            const uint32_t buf = tile & 1;
            mbar_wait(tfull0 + 8 * buf, (tile >> 1) & 1);
            tc_fence_after();
            int r_lt = 0;      // columns of this thread's query row that beat the true entity (raw and filtered counters alike:
                               // bil_known_kernel takes the known-true entities back out of the filtered ones)
            const uint32_t taddr = tmem_base + buf * BN + ((uint32_t)(quarter * 32) << 16);
            // better <=> larger similarity (predict = -sim).  s > thr_hi: counted; s < thr_lo: not; in between: near-tie
            const float thr_hi = sim_true + guard, thr_lo = sim_true - guard;
            // One chunk = 32 accumulator columns of this thread's row.  The two comparisons are collected as BIT MASKS, so
            // the counts and the near-tie set are a handful of popc / and instructions and the
            // cold paths below are rolled loops: the kernel stays a few thousand instructions (it was 22 K when every cold
            // path was unrolled 32x, which thrashed the instruction cache under the MMA-issuing warp).
            auto process = [&](const uint32_t (&v)[32], int c0) {
                const int left = ne - c0;                       // warp-uniform
                if (STORE) {                                    // diagnostic / materialised-score mode
                    if (q_ok) {
                        float *o = bp.store + (qbase + row) * bp.store_ld + cur.crow0 + c0;
#pragma unroll
                        for (int c = 0; c < 32; c++)
                            if (c < left) o[c] = __uint_as_float(v[c]);
                    }
                    return;
                }
                // bit c of gtm: column c0 + c beats thr_hi; bit c of ltm: it falls short of thr_lo.  Both are SIGN bits:
                // x > y <=> (y - x) < 0 exactly (a difference of distinct floats never rounds to zero), so each compare is one
                // FADD and one funnel shift that collects the sign -- no predicates, four independent 8-column chains per mask.
                uint32_t g4[4] = {0u, 0u, 0u, 0u}, l4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int k = 0; k < 4; k++) {
#pragma unroll
                    for (int c = 7; c >= 0; c--) {              // descending: column 8 k + c ends up at bit c of the byte
                        const float sc = __uint_as_float(v[8 * k + c]);
                        g4[k] = __funnelshift_l(__float_as_uint(thr_hi - sc), g4[k], 1);
                        l4[k] = __funnelshift_l(__float_as_uint(sc - thr_lo), l4[k], 1);
                    }
                }
                const uint32_t valid = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);   // padding of the last candidate tile
                const uint32_t gtm = (g4[0] | (g4[1] << 8) | (g4[2] << 16) | (g4[3] << 24)) & valid;
                const uint32_t gem = ~(l4[0] | (l4[1] << 8) | (l4[2] << 16) | (l4[3] << 24)) & valid;
                r_lt += __popc(gtm);
                uint32_t near = gem & ~gtm;                     // rare: park the near-ties for the exact re-score
                if (__any_sync(0xffffffffu, near != 0u)) {
                    // slots by a warp prefix sum of the per-lane counts: no atomics
                    const int mine = __popc(near);
                    int incl = mine;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int up = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += up;
                    }
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
                    if (total <= PEND_CAP) {
                        while (n_pub + total - ctl[1] > (uint32_t)PEND_CAP) __nanosleep(64);   // ring full: the re-score warp is draining it
                        uint32_t slot = n_pub + incl - mine;
                        while (near) {
                            const int c = __ffs(near) - 1;
                            near &= near - 1;
                            const int64_t crow = cur.crow0 + c0 + c;
                            const int64_t ent_id = p.all_entities ? crow : __ldg(p.cand_idx + crow);
                            pend[(slot++) & (PEND_CAP - 1)] = make_uint2((uint32_t)(qbase + row), (uint32_t)ent_id);
                        }
                        n_pub += total;
                        __syncwarp();
                        __threadfence_block();                  // entries before the counter
                        if (lane == 0) ctl[0] = n_pub;
                    } else {                                    // a chunk with more near-ties than the ring holds (mass ties): in place
                        while (near) {
                            const int c = __ffs(near) - 1;
                            near &= near - 1;
                            const int64_t crow = cur.crow0 + c0 + c;
                            const int64_t ent_id = p.all_entities ? crow : __ldg(p.cand_idx + crow);
                            rescore(make_uint2((uint32_t)(qbase + row), (uint32_t)ent_id));
                        }
                    }
                    __syncwarp();
                }
            };
            // two register buffers: the TMEM load of the next 32 columns is in flight while this chunk is compared
            uint32_t va[32], vb[32];
            constexpr int CSLICE = BN / (EPI_WARPS / 4);
            const int cbeg = chalf * CSLICE, cend = min(ne, cbeg + CSLICE);   // this warp's slice of the tile's columns
#if MRE_DIAG_BIL_NOEPI
            if (false)      // diagnostic build only (wrong results): hand the accumulator buffer straight back
#endif
#if MRE_DIAG_BIL_NOEPI == 2      // diagnostic (wrong results): TMEM loads only
            for (int c0 = cbeg; c0 < cend; c0 += 32) { tmem_ld_32x32(taddr + c0, va); tmem_ld_wait(); }
            if (va[0] == 0x12345678u && va[31] == 0x9abcdef0u) r_lt++;
#elif MRE_DIAG_BIL_NOEPI == 3    // diagnostic (wrong results): compares only, no TMEM loads
#pragma unroll
            for (int c = 0; c < 32; c++) { va[c] = (uint32_t)(row * c + tile); vb[c] = va[c] ^ 0x3f800000u; }
            for (int c0 = cbeg; c0 < cend; c0 += 64) { process(va, c0); process(vb, c0 + 32); }
#else
            if (cbeg < cend) tmem_ld_32x32(taddr + cbeg, va);
#pragma unroll 1
            for (int c0 = cbeg; c0 < (MRE_DIAG_BIL_NOEPI ? 0 : cend); c0 += 64) {          // cend is warp-uniform: the collective loads stay aligned
                tmem_ld_wait();
                if (c0 + 32 < cend) tmem_ld_32x32(taddr + c0 + 32, vb);
                process(va, c0);
                if (c0 + 32 < cend) {
                    tmem_ld_wait();
                    if (c0 + 64 < cend) tmem_ld_32x32(taddr + c0 + 64, va);
                    process(vb, c0 + 32);
                }
            }
#endif
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {     // relaxed: the TMEM reads are complete (wait::ld); nothing in global memory is ordered by this barrier
                if (PAIR) mbar_arrive_cluster_relaxed(tempty_leader0 + 8 * buf);
                else mbar_arrive_relaxed(tempty0 + 8 * buf);
            }
            if (q_ok) {
                const int64_t q = qbase + row;
                if (r_lt) { atomicAdd(p.counts + q, r_lt); atomicAdd(p.counts + 2 * p.Q + q, r_lt); }
            }
        }
        __syncwarp();
        __threadfence_block();
        if (lane == 0) ctl[2] = 1u;                       // no more entries will be published
    } else {
        // ================================================= re-score warps ===============================================
        // NPROD = 3 (near-ties are rare, ~5e-4 of the columns): entries are re-scored as soon as they are published, one per lane,
        // spread over the kernel's whole run.  NPROD = 1 (~1e-2 of the columns): drained in batches so that all 32 lanes work,
        // every lane interleaving four sequential dot products (a single FP32 chain issues one add every ~4 cycles).
        constexpr uint32_t BATCH = NPROD == 3 ? 1u : (uint32_t)(PEND_CAP / 4);
        const int ring_id = warp - EPI_WARP0 - EPI_WARPS;
        const uint2 *pend = pend_all + ring_id * PEND_CAP;
        volatile uint32_t *ctl = pend_ctl + ring_id * 4;
        uint32_t cons = 0;
        const int n4 = (int)(p.D >> 2);
        for (;;) {
            const uint32_t done = ctl[2];               // read BEFORE the counter: done => the counter is final
            const uint32_t pub = ctl[0];
            if (pub - cons < BATCH && !(done && pub != cons)) {
                if (done && pub == cons) break;
                __nanosleep(NPROD == 3 ? 200 : 100);
                continue;
            }
            __threadfence_block();                      // the counter before the entries
            const uint32_t n_new = pub - cons;          // <= PEND_CAP
            if (NPROD == 3) {
                for (uint32_t k = lane; k < n_new; k += 32) {
                    const uint2 it2 = pend[(cons + k) & (PEND_CAP - 1)];
                    const int64_t q2 = it2.x, ent_id = it2.y;
                    const float s2 = bil_dot(p.qvec + q2 * p.D, p.ent + ent_id * p.D, p.D);
                    const float st = -__ldg(&p.thr[q2].x);
                    if (s2 > st) { atomicAdd(p.counts + q2, 1); atomicAdd(p.counts + 2 * p.Q + q2, 1); }
                    if (s2 == st) { atomicAdd(p.counts + p.Q + q2, 1); atomicAdd(p.counts + 3 * p.Q + q2, 1); }
                }
            } else
            for (uint32_t k0 = 0; k0 < n_new; k0 += 128) {
                int64_t q2[4], ent_id[4];
                bool live[4];
                const float4 *v4[4], *e4[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t k = k0 + 32 * u + lane;
                    live[u] = k < n_new;
                    const uint2 it2 = live[u] ? pend[(cons + k) & (PEND_CAP - 1)] : make_uint2(0u, 0u);
                    q2[u] = it2.x; ent_id[u] = it2.y;
                    v4[u] = reinterpret_cast<const float4 *>(p.qvec + q2[u] * p.D);
                    e4[u] = reinterpret_cast<const float4 *>(p.ent + ent_id[u] * p.D);
                }
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
                for (int d = 0; d < n4; d++) {
                    float4 a[4], b[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) { a[u] = __ldg(v4[u] + d); b[u] = __ldg(e4[u] + d); }
#pragma unroll
                    for (int u = 0; u < 4; u++) acc[u] = acc[u] + a[u].x * b[u].x;
#pragma unroll
                    for (int u = 0; u < 4; u++) acc[u] = acc[u] + a[u].y * b[u].y;
#pragma unroll
                    for (int u = 0; u < 4; u++) acc[u] = acc[u] + a[u].z * b[u].z;
#pragma unroll
                    for (int u = 0; u < 4; u++) acc[u] = acc[u] + a[u].w * b[u].w;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (!live[u]) continue;
                    const float st = -__ldg(&p.thr[q2[u]].x);
                    if (acc[u] > st) { atomicAdd(p.counts + q2[u], 1); atomicAdd(p.counts + 2 * p.Q + q2[u], 1); }
                    if (acc[u] == st) { atomicAdd(p.counts + p.Q + q2[u], 1); atomicAdd(p.counts + 3 * p.Q + q2[u], 1); }
                }
            }
            __syncwarp();
            cons = pub;
            if (lane == 0) {
                ctl[1] = cons;                          // the ring slots may be reused
                atomicAdd(bp.rescored, (unsigned long long)n_new);
            }
        }
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all();      // neither CTA may retire while its partner can still touch its barriers / shared memory
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
        else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------ dense MMA peak probes
constexpr uint32_t PROBE_A = BM * 128, PROBE_B = BN * 128;   // one 128-byte-swizzled k-block of each operand
template <bool BF16>
__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int iters) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *al = smem_raw + (base - smem_u32(smem_raw));
    uint64_t *bar = reinterpret_cast<uint64_t *>(al + PROBE_A + PROBE_B);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
    for (uint32_t i = threadIdx.x; i < (PROBE_A + PROBE_B) / 4; i += blockDim.x) reinterpret_cast<float *>(al)[i] = 0.f;
    if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); fence_barrier_init(); }
    fence_proxy_async();
    if (threadIdx.x < 32) tmem_alloc(smem_u32(slot), 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = BF16 ? umma_idesc_bf16(BM, BN) : umma_idesc_tf32(BM, BN);
        for (int i = 0; i < iters; i++) {
            const uint32_t koff = (i & 3) * 32;   // one k-step = 32 bytes of the 128-byte swizzle atom in both kinds
            if (BF16) umma_bf16(tmem, umma_desc_k128(base + koff), umma_desc_k128(base + PROBE_A + koff), idesc, i != 0);
            else umma_tf32(tmem, umma_desc_k128(base + koff), umma_desc_k128(base + PROBE_A + koff), idesc, i != 0);
        }
        umma_commit(smem_u32(bar));
        mbar_wait(smem_u32(bar), 0);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

template <bool BF16>
static int probe_mma_peak(mre_ctx *ctx, double *flops_per_s) {
    MRE_CHECK_ARG(flops_per_s != nullptr, "NULL output");
    const size_t smem = 1024 + PROBE_A + PROBE_B + 64;
    MRE_CUDA(cudaFuncSetAttribute(mma_probe_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 8192;
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        MRE_CUDA(cudaEventRecord(ctx->ev0, 0));
        mma_probe_kernel<BF16><<<ctx->sm_count, 128, smem>>>(iters);
        MRE_CUDA(cudaEventRecord(ctx->ev1, 0));
        MRE_CUDA(cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        MRE_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double flops = (double)ctx->sm_count * iters * 2.0 * BM * BN * (BF16 ? UK : TF_UK);
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3));
    }
    ctx->launches += 4;
    MRE_CUDA(cudaGetLastError());
    *flops_per_s = best;
    return MRE_OK;
}

int probe_tf32_peak(mre_ctx *ctx, double *flops_per_s) { return probe_mma_peak<false>(ctx, flops_per_s); }
int probe_bf16_peak(mre_ctx *ctx, double *flops_per_s) { return probe_mma_peak<true>(ctx, flops_per_s); }

// ------------------------------------------------------------------------------------------ host side
struct BilScratch {
    const float *ent_full;   // [E, Kp] full precision (the scalar scorer's table)
    const uint16_t *ent_hi, *ent_lo;   // [E, K8] BF16 hi / lo (NPROD = 3) or FP16 (NPROD = 1; lo unused)
    const uint16_t *q_hi, *q_lo;       // [Q, K8]
    float2 *thr;
    float *delta;
    int64_t K, Kp, K8;
};

// the two pre-pass launches: table operands + largest row norm, then everything per query (vector, operands, threshold, guard,
// counter zeroing).  `counts` may be NULL for callers that only want the query vectors (predict).
static int bil_prepass(mre_ctx *ctx, const mre_rank_job *job, int nprod, int32_t *counts, cudaStream_t st, BilScratch &sc) {
    const int64_t D = job->D, Q = std::max<int64_t>(job->Q, 1);
    sc.K = job->scorer == MRE_COMPLEX ? 2 * D : D;
    sc.Kp = (sc.K + 3) & ~(int64_t)3;
    sc.K8 = (sc.K + 7) & ~(int64_t)7;
    const size_t half = ((size_t)job->E * sc.K8 * sizeof(uint16_t) + 255) & ~(size_t)255;
    // layout of ctx->ent_n: [hi | lo | full (only when a repacked full-precision copy is needed)]
    const bool need_full = job->scorer == MRE_COMPLEX || sc.Kp != D;
    MRE_TRY(ctx->ent_n.reserve(2 * half + (need_full ? (size_t)job->E * sc.Kp * sizeof(float) : 0)));
    char *base = ctx->ent_n.as<char>();
    uint16_t *hi = reinterpret_cast<uint16_t *>(base), *lo = reinterpret_cast<uint16_t *>(base + half);
    float *full = need_full ? reinterpret_cast<float *>(base + 2 * half) : nullptr;
    const size_t qhalf = ((size_t)Q * sc.K8 * sizeof(uint16_t) + 255) & ~(size_t)255;
    MRE_TRY(ctx->qvec.reserve((size_t)Q * sc.Kp * sizeof(float)));
    MRE_TRY(ctx->qvec2.reserve(2 * qhalf));
    MRE_TRY(ctx->thr.reserve((size_t)Q * (sizeof(float2) + sizeof(float)) + (counts ? 0 : (size_t)4 * Q * sizeof(int32_t)) + 16));
    sc.thr = ctx->thr.as<float2>();
    sc.delta = reinterpret_cast<float *>(sc.thr + Q);
    if (!counts) counts = reinterpret_cast<int32_t *>(sc.delta + Q);
    if (!ctx->stats.p) {
        MRE_TRY(ctx->stats.reserve(64));
        MRE_CUDA(cudaMemset(ctx->stats.p, 0, 64));
    }
    unsigned long long *max_norm = ctx->stats.as<unsigned long long>() + 1;     // slot 1; slot 0 counts the exact re-scores
    if (++ctx->bil_epoch == 0xffffffffu) {            // generation tags are exhausted once in 4 billion calls
        MRE_CUDA(cudaMemsetAsync(max_norm, 0, sizeof(*max_norm), st));
        ctx->bil_epoch = 1;
    }
    const unsigned int epoch = ctx->bil_epoch;
    sc.ent_full = need_full ? full : job->ent;
    sc.ent_hi = hi; sc.ent_lo = lo;
    sc.q_hi = ctx->qvec2.as<uint16_t>();
    sc.q_lo = reinterpret_cast<const uint16_t *>(ctx->qvec2.as<char>() + qhalf);
    const int tgrid = (int)std::max<int64_t>(1, std::min<int64_t>((job->E + 7) / 8, (int64_t)ctx->sm_count * 8));
    if (nprod == 3) bil_table_kernel<3><<<tgrid, 256, 0, st>>>(job->ent, job->ent_im, job->E, D, sc.K, sc.Kp, sc.K8, full, hi, lo, max_norm, epoch);
    else bil_table_kernel<1><<<tgrid, 256, 0, st>>>(job->ent, job->ent_im, job->E, D, sc.K, sc.Kp, sc.K8, full, hi, lo, max_norm, epoch);
    ctx->launches += 1;
    if (job->Q > 0) {
        const size_t smem = 0;
        const int qgrid = (int)std::max<int64_t>(1, std::min<int64_t>((job->Q + QK_G * QK_WARPS - 1) / (QK_G * QK_WARPS), (int64_t)ctx->sm_count * 8));
        uint16_t *qhi = const_cast<uint16_t *>(sc.q_hi), *qlo = const_cast<uint16_t *>(sc.q_lo);
        if (nprod == 3)
            bil_query_kernel<3><<<qgrid, QK_WARPS * 32, smem, st>>>(job->scorer, job->ent, job->ent_im, job->rel, job->rel_im, sc.ent_full, D, sc.K,
                                                                  sc.Kp, sc.K8, job->q_h, job->q_t, job->q_r, job->q_side, job->side, job->Q,
                                                                  max_norm, epoch, ctx->qvec.as<float>(), qhi, qlo, sc.thr, sc.delta, counts);
        else
            bil_query_kernel<1><<<qgrid, QK_WARPS * 32, smem, st>>>(job->scorer, job->ent, job->ent_im, job->rel, job->rel_im, sc.ent_full, D, sc.K,
                                                                  sc.Kp, sc.K8, job->q_h, job->q_t, job->q_r, job->q_side, job->side, job->Q,
                                                                  max_norm, epoch, ctx->qvec.as<float>(), qhi, qlo, sc.thr, sc.delta, counts);
        ctx->launches += 1;
    }
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

template <bool STORE, bool PAIR, int NPROD>
static int launch_bilinear(mre_ctx *ctx, int grid, const BilParams &bp, const CUtensorMap &tm_ahi, const CUtensorMap &tm_alo,
                           const CUtensorMap &tm_bhi, const CUtensorMap &tm_blo, cudaStream_t st) {
    auto kern = bilinear_rank_kernel<STORE, PAIR, NPROD>;
    MRE_TRY(ctx->allow_smem(reinterpret_cast<const void *>(kern), BIL_SMEM));
    if (PAIR) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(BIL_THREADS);
        cfg.dynamicSmemBytes = BIL_SMEM;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MRE_CUDA(cudaLaunchKernelEx(&cfg, kern, bp, tm_ahi, tm_alo, tm_bhi, tm_blo));
    } else {
        kern<<<grid, BIL_THREADS, BIL_SMEM, st>>>(bp, tm_ahi, tm_alo, tm_bhi, tm_blo);
    }
    return MRE_OK;
}

static int run_bilinear(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, float *store, cudaStream_t st) {
    BilScratch sc{};
    BilParams bp{};
    RankParams &p = bp.r;
    const int nprod = ctx->opt_bil_products == 1 ? 1 : 3;
    MRE_TRY(fill_rank_params(ctx, ix, job, BM, BN, st, p));
    const bool shared_runs = !store && job->Q > 0 && p.filter == MRE_FILTER_INDEX;
    const unsigned qgrid = (unsigned)((std::max<int64_t>(p.Q, 1) + KNOWN_WARPS - 1) / KNOWN_WARPS);
    MRE_TRY(bil_prepass(ctx, job, nprod, job->counts, st, sc));
    p.ent = sc.ent_full;
    p.D = sc.Kp;
    bp.k8 = sc.K8;
    p.qvec = ctx->qvec.as<float>();
    if (job->Q == 0) return MRE_OK;
    p.thr = sc.thr;
    bp.delta = sc.delta;
    if (!store) {   // known-true correction of the filtered counters (needs the query vectors, thresholds and zeroed counters)
        if (shared_runs) {                       // every run the job touches scored once, then every query's thresholds against them
            KnownRuns kr{};
            MRE_TRY(known_runs_scratch(ctx, p, kr));
            bil_known_score_kernel<<<qgrid, KNOWN_WARPS * 32, 0, st>>>(p, kr);
            bil_known_compare_kernel<<<qgrid, KNOWN_WARPS * 32, 0, st>>>(p, kr);
            ctx->launches += 1;
        } else if (known_is_flat(p))
            bil_known_flat_kernel<<<(unsigned)((p.filt_nnz + p.Q + KNOWN_WARPS * 32 - 1) / (KNOWN_WARPS * 32)), KNOWN_WARPS * 32, 0, st>>>(p);
        else
            bil_known_kernel<<<(unsigned)((p.Q + KNOWN_WARPS - 1) / KNOWN_WARPS), KNOWN_WARPS * 32, 0, st>>>(p);
        ctx->launches += 1;
        MRE_CUDA(cudaGetLastError());
    }
    // candidate tables the B tiles stream from
    const uint16_t *b_hi = sc.ent_hi, *b_lo = sc.ent_lo;
    int64_t cand_rows = job->E;
    if (!p.all_entities) {
        cand_rows = job->group_cptr[job->n_groups];
        MRE_CHECK_ARG(cand_rows < (1LL << 31), "too many candidate rows");
        const size_t cb = ((size_t)std::max<int64_t>(cand_rows, 1) * sc.K8 * sizeof(uint16_t) + 255) & ~(size_t)255;
        MRE_TRY(ctx->ent_aux.reserve(2 * cb));
        uint16_t *g_hi = ctx->ent_aux.as<uint16_t>(), *g_lo = reinterpret_cast<uint16_t *>(ctx->ent_aux.as<char>() + cb);
        if (cand_rows > 0) {
            // a 16-bit row of K8 elements is K8 / 2 floats (a multiple of 4): the float4 row gather serves it unchanged
            const int64_t row_f = sc.K8 >> 1;
            gather_rows_kernel<<<grid_for(cand_rows * (row_f >> 2), 256), 256, 0, st>>>(reinterpret_cast<const float *>(sc.ent_hi), row_f,
                                                                                        job->cand_idx, cand_rows, reinterpret_cast<float *>(g_hi));
            ctx->launches += 1;
            if (nprod == 3) {
                gather_rows_kernel<<<grid_for(cand_rows * (row_f >> 2), 256), 256, 0, st>>>(reinterpret_cast<const float *>(sc.ent_lo), row_f,
                                                                                            job->cand_idx, cand_rows, reinterpret_cast<float *>(g_lo));
                ctx->launches += 1;
            }
        }
        b_hi = g_hi;
        b_lo = g_lo;
    }
    // CTA pairs (cta_group::2) whenever there is more than one query tile; a lone CTA per tile otherwise
    const bool pair = p.total_items > p.total_pitems && ctx->sm_count >= 2 && ctx->opt_bil_pair;
    CUtensorMap tm_ahi, tm_alo, tm_bhi, tm_blo;
    MRE_TRY(make_tmap_bf16_2d(&tm_ahi, sc.q_hi, job->Q, sc.K8, sc.K8, BM, BK));
    MRE_TRY(make_tmap_bf16_2d(&tm_alo, sc.q_lo, job->Q, sc.K8, sc.K8, BM, BK));
    MRE_TRY(make_tmap_bf16_2d(&tm_bhi, b_hi, std::max<int64_t>(cand_rows, 1), sc.K8, sc.K8, pair ? BN / 2 : BN, BK));
    MRE_TRY(make_tmap_bf16_2d(&tm_blo, b_lo, std::max<int64_t>(cand_rows, 1), sc.K8, sc.K8, pair ? BN / 2 : BN, BK));
    const int grid = pair ? 2 * (int)std::max<int64_t>(1, std::min<int64_t>(p.total_pitems, ctx->sm_count / 2))
                          : (int)std::max<int64_t>(1, std::min<int64_t>(p.total_items, ctx->sm_count));
    bp.store = store;
    bp.store_ld = cand_rows;
    bp.rescored = ctx->stats.as<unsigned long long>();
    MRE_TRY(ctx->time_begin(st));
    int rc;
    if (nprod == 3) {
        if (pair) rc = store ? launch_bilinear<true, true, 3>(ctx, grid, bp, tm_ahi, tm_alo, tm_bhi, tm_blo, st)
                             : launch_bilinear<false, true, 3>(ctx, grid, bp, tm_ahi, tm_alo, tm_bhi, tm_blo, st);
        else rc = store ? launch_bilinear<true, false, 3>(ctx, grid, bp, tm_ahi, tm_alo, tm_bhi, tm_blo, st)
                        : launch_bilinear<false, false, 3>(ctx, grid, bp, tm_ahi, tm_alo, tm_bhi, tm_blo, st);
    } else {
        if (pair) rc = store ? launch_bilinear<true, true, 1>(ctx, grid, bp, tm_ahi, tm_alo, tm_bhi, tm_blo, st)
                             : launch_bilinear<false, true, 1>(ctx, grid, bp, tm_ahi, tm_alo, tm_bhi, tm_blo, st);
        else rc = store ? launch_bilinear<true, false, 1>(ctx, grid, bp, tm_ahi, tm_alo, tm_bhi, tm_blo, st)
                        : launch_bilinear<false, false, 1>(ctx, grid, bp, tm_ahi, tm_alo, tm_bhi, tm_blo, st);
    }
    MRE_TRY(rc);
    MRE_TRY(ctx->time_end(st));
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

int rank_bilinear(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, cudaStream_t st) {
    return run_bilinear(ctx, ix, job, nullptr, st);
}

int bilinear_scores(mre_ctx *ctx, const mre_rank_job *job, float *scores_out, cudaStream_t st) {
    MRE_CHECK_ARG(scores_out != nullptr, "scores_out is NULL");
    mre_rank_job j = *job;
    j.filter = MRE_FILTER_NONE;
    MRE_TRY(ctx->misc.reserve(256 + (size_t)4 * std::max<int64_t>(job->Q, 1) * sizeof(int32_t)));
    j.counts = reinterpret_cast<int32_t *>(ctx->misc.as<char>() + 256);   // zeroed, otherwise unused in STORE mode
    return run_bilinear(ctx, nullptr, &j, scores_out, st);
}

int predict_bilinear(mre_ctx *ctx, const mre_rank_job *job, int64_t query, float *scores_out, cudaStream_t st) {
    mre_rank_job one = *job;
    one.q_h = job->q_h + query; one.q_t = job->q_t + query; one.q_r = job->q_r + query;
    one.q_side = job->q_side ? job->q_side + query : nullptr;
    one.Q = 1;
    BilScratch sc{};
    MRE_TRY(bil_prepass(ctx, &one, 3, nullptr, st, sc));
    bil_predict_kernel<<<(unsigned)((job->E + 127) / 128), 128, 0, st>>>(sc.ent_full, job->E, sc.Kp, ctx->qvec.as<float>(), scores_out);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

}  // namespace mre
