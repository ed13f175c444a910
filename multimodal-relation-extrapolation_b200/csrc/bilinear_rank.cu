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
// back with tcgen05.ld (one query row per thread), compare against the query's true score and count, taking the
// known-true columns routed to the tile by tile_filter.cu back out of the filtered count; columns closer to the true
// score than an error guard are re-scored in scalar FP32, which makes the COUNTS those of the sequential FP32 scorer.
// Precision: the reference is FP32.  kind::tf32 keeps 11 significant bits, which would move ranks well outside the
// 1e-5 tie band, so every operand is split x ~= hi + lo (hi = rn_tf32(x), lo = rn_tf32(x - hi): 22 significant bits,
// both exactly representable so the MMA's operand truncation loses nothing) and each product is issued as THREE TF32
// MMAs  hi*hi + lo*hi + hi*lo  (the dropped terms are <= 2^-22 relative and unbiased).
// Algorithmic flops are counted once (2*Q*E*K); the tensor pipe executes 3x that.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected thread),
// warps 2..5 = epilogue (TMEM lane quarter = warp % 4).  Shared memory: 2 stages x {A_hi, A_lo: 128 x 128 B;
// B_hi, B_lo: 256 x 128 B}, 128-byte swizzle, K-major, fed by TMA tensor tiles.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.h"
#include "rank_common.cuh"
#include "rank_host.h"
#include "tma_host.h"

namespace mre {

constexpr int BN = 256;                 // entities per tile (UMMA N)
constexpr int BM = 128;                 // queries per tile (UMMA M)
constexpr int BK = 32;                  // floats of K per stage = one 128-byte swizzle atom
constexpr int UK = 8;                   // floats of K per tcgen05.mma kind::tf32
constexpr int B_STAGES = 2;
constexpr uint32_t A_BYTES = BM * BK * 4;   // 16 KiB
constexpr uint32_t B_BYTES = BN * BK * 4;   // 32 KiB
constexpr uint32_t BSTAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
constexpr int BIL_THREADS = 192;
constexpr int EPI_WARP0 = 2;
constexpr int MASK_STRIDE = 9;               // words per row of the per-warp known-true mask (8 + 1 pad: conflict-free)
constexpr size_t BIL_SMEM = 1024 + (size_t)B_STAGES * BSTAGE_BYTES + 16 * sizeof(uint64_t) + 4 * 32 * MASK_STRIDE * 4 + 64;
constexpr uint32_t TMEM_COLS = 512;     // two 256-column accumulator buffers

struct BilParams {
    RankParams r;            // r.ent = full-precision [E, K] table (scalar scorer), r.qvec = full-precision query vectors
    const float *delta;      // [Q] near-tie guard: |s_mma - s_true| <= delta => the column is re-scored in scalar FP32
    uint2 *tie_queue;        // [gridDim.x][tie_cap] (query, entity) pairs awaiting the exact re-score
    uint32_t tie_cap;        // queue entries per CTA
    float *store;            // STORE mode only: [Q, store_ld] tensor-core similarities are written instead of counted
    int64_t store_ld;
};

// ------------------------------------------------------------------------------------------ pre-pass kernels
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// entity table -> [rows, Kp] full-precision copy (ComplEx: [re | im]; zero padded) + TF32 hi / lo splits
__global__ void bil_split_table_kernel(const float *__restrict__ re, const float *__restrict__ im, int64_t rows, int64_t D,
                                       int64_t K, int64_t Kp, float *__restrict__ full, float *__restrict__ hi,
                                       float *__restrict__ lo) {
    const int64_t total = rows * Kp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / Kp, d = i - row * Kp;
        float x = 0.f;
        if (d < K) x = d < D ? re[row * D + d] : im[row * D + (d - D)];
        const float h = tf32_rn(x);
        if (full) full[i] = x;
        hi[i] = h;
        lo[i] = tf32_rn(x - h);   // exactly representable: the MMA's operand truncation then loses nothing
    }
}

// per-query vector (see the header comment), full precision + hi / lo
__global__ void bil_qvec_kernel(int scorer, const float *__restrict__ ent, const float *__restrict__ ent_im,
                                const float *__restrict__ rel, const float *__restrict__ rel_im, int64_t D, int64_t K, int64_t Kp,
                                const int64_t *__restrict__ q_h, const int64_t *__restrict__ q_t, const int64_t *__restrict__ q_r,
                                const uint8_t *__restrict__ q_side, int side, int64_t Q, float *__restrict__ qv,
                                float *__restrict__ qhi, float *__restrict__ qlo) {
    const int64_t total = Q * Kp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = i / Kp, d = i - q * Kp;
        const int s = q_side ? (int)q_side[q] : side;
        const int64_t e = s ? q_h[q] : q_t[q];   // the entity that stays fixed in the query
        const int64_t r = q_r[q];
        float v = 0.f;
        if (d < K) {
            if (scorer == MRE_DISTMULT) {
                v = s ? ent[e * D + d] * rel[r * D + d] : rel[r * D + d] * ent[e * D + d];
            } else {
                const int64_t dd = d < D ? d : d - D;
                const float ere = ent[e * D + dd], eim = ent_im[e * D + dd], rre = rel[r * D + dd], rim = rel_im[r * D + dd];
                if (s) v = d < D ? ere * rre - eim * rim : eim * rre + ere * rim;
                else v = d < D ? ere * rre + eim * rim : eim * rre - ere * rim;
            }
        }
        const float h = tf32_rn(v);
        qv[i] = v;
        qhi[i] = h;
        qlo[i] = tf32_rn(v - h);
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

// largest row norm of the entity table (positive floats order like their bit patterns)
__global__ void bil_max_rownorm_kernel(const float *__restrict__ ent, int64_t E, int64_t K, unsigned int *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    float best = 0.f;
    for (int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < E; j += warps) {
        float ss = 0.f;
        for (int64_t d = lane; d < K; d += 32) ss = fmaf(ent[j * K + d], ent[j * K + d], ss);
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, m);
        best = fmaxf(best, ss);
    }
    if (lane == 0) atomicMax(out, __float_as_uint(sqrtf(best) * 1.0001f));
}

// per query: threshold pair on the predict scale (p = -sim: lower is better) and the near-tie guard.
// Guard: the tensor-core value differs from the sequential FP32 value by the split error (<= 3 * 2^-22), the
// accumulator's rounding over <= 3K/8 MMAs and the scalar sum's own rounding, all relative to
// sum_d |v_d e_d| <= ||v|| * max_j ||e_j||.  The all-errors-aligned worst case is ~2^-15 (K = 256); rounding errors
// do not align, and the largest discrepancy measured on any test table is < 2^-22 (tests/test_bilinear_gpu.py asserts
// it stays 4x under the guard), so the guard is 2^-18 (scaled linearly beyond K = 256).  Every column closer than the
// guard to s_true is re-scored with the scalar scorer, so the COUNTS are exactly those of the FP32 scorer whenever the
// discrepancy is below the guard; if it ever were not, only columns within 2^-18 relative of s_true could flip --
// well inside the 1e-5 tie band the reference itself cannot resolve.
__global__ void bil_threshold_kernel(const RankParams p, const unsigned int *__restrict__ max_norm, float2 *__restrict__ thr,
                                     float *__restrict__ delta) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= p.Q) return;
    const int s = p.q_side ? (int)p.q_side[q] : p.side;
    const int64_t truth = s ? p.q_t[q] : p.q_h[q];
    const float *v = p.qvec + q * p.D;
    const float pt = -bil_dot(v, p.ent + truth * p.D, p.D);
    float hi = pt;
    if (pt == pt && fabsf(pt) < INFINITY) hi = nextafterf(pt, INFINITY);
    thr[q] = make_float2(pt, hi);
    float ss = 0.f;
    for (int64_t d = 0; d < p.D; d++) ss = fmaf(v[d], v[d], ss);
    const float scale = p.D > 256 ? (float)p.D / 256.f : 1.f;
    delta[q] = 3.814697265625e-06f * scale * sqrtf(ss) * __uint_as_float(*max_norm);   // 2^-18
}

__global__ void bil_predict_kernel(const float *__restrict__ ent, int64_t E, int64_t K, const float *__restrict__ qv,
                                   float *__restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < E) out[j] = -bil_dot(qv, ent + j * K, K);
}

// ------------------------------------------------------------------------------------------ main kernel
template <bool STORE>
__global__ void __launch_bounds__(BIL_THREADS, 1)
bilinear_rank_kernel(const BilParams bp, const __grid_constant__ CUtensorMap tm_ahi, const __grid_constant__ CUtensorMap tm_alo,
                     const __grid_constant__ CUtensorMap tm_bhi, const __grid_constant__ CUtensorMap tm_blo) {
    const RankParams &p = bp.r;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t ring_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *ring = smem_raw + (ring_u32 - smem_u32(smem_raw));
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring + (size_t)B_STAGES * BSTAGE_BYTES);
    // bars: full[2], empty[2], tmem_full[2], tmem_empty[2]; then the TMEM base address word
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + 2), tfull0 = smem_u32(bars + 4), tempty0 = smem_u32(bars + 6);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);
    uint32_t *tie_count = reinterpret_cast<uint32_t *>(bars + 9);      // near-ties queued by this CTA
    uint32_t *mask_all = reinterpret_cast<uint32_t *>(bars + 16);      // per epilogue warp: [32 rows][MASK_STRIDE] known-true bits
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_kb = (int)((p.D + BK - 1) / BK);

    if (threadIdx.x == 0) {
        for (int s = 0; s < B_STAGES; s++) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
            mbar_init(tfull0 + 8 * s, 1);
            mbar_init(tempty0 + 8 * s, 4);
        }
        *tie_count = 0;
        fence_barrier_init();
        tma_prefetch_desc(&tm_ahi); tma_prefetch_desc(&tm_alo); tma_prefetch_desc(&tm_bhi); tma_prefetch_desc(&tm_blo);
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================= TMA producer =================================================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                int g, qt, et;
                decode_item(p, item, g, qt, et);
                const GroupDesc &gd = p.groups[g];
                const int qrow = (int)(gd.q0 + (int64_t)qt * BM);
                const int erow = (int)(gd.c0 + (int64_t)et * BN);
                for (int kb = 0; kb < n_kb; kb++, it++) {
                    const int s = it % B_STAGES;
                    mbar_wait(empty0 + 8 * s, ((it / B_STAGES) & 1) ^ 1);
                    const uint32_t full = full0 + 8 * s;
                    const uint32_t base = ring_u32 + (uint32_t)s * BSTAGE_BYTES;
                    mbar_arrive_expect_tx(full, BSTAGE_BYTES);
                    // query tiles are re-read for every candidate tile of the sweep: keep them in L2 (evict-last) so the
                    // streaming candidate tiles (each shared by the ~128 CTAs of one round, then dead) cannot push them out
                    tma_load_2d_hint(base, &tm_ahi, kb * BK, qrow, full, L2_EVICT_LAST);
                    tma_load_2d_hint(base + A_BYTES, &tm_alo, kb * BK, qrow, full, L2_EVICT_LAST);
                    tma_load_2d_hint(base + 2 * A_BYTES, &tm_bhi, kb * BK, erow, full, L2_EVICT_NORMAL);
                    tma_load_2d_hint(base + 2 * A_BYTES + B_BYTES, &tm_blo, kb * BK, erow, full, L2_EVICT_NORMAL);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================================= MMA issuer ===================================================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(BM, BN);
            uint32_t it = 0, tile = 0;
            for (int64_t item = blockIdx.x; item < p.total_items; item += gridDim.x, tile++) {
                const uint32_t buf = tile & 1;
                mbar_wait(tempty0 + 8 * buf, ((tile >> 1) & 1) ^ 1);   // epilogue has drained this accumulator buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BN;
                for (int kb = 0; kb < n_kb; kb++, it++) {
                    const int s = it % B_STAGES;
                    mbar_wait(full0 + 8 * s, (it / B_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t base = ring_u32 + (uint32_t)s * BSTAGE_BYTES;
                    const int n_ks = (int)min((int64_t)(BK / UK), (p.D - (int64_t)kb * BK + UK - 1) / UK);
                    for (int ks = 0; ks < n_ks; ks++) {
                        const uint32_t koff = ks * UK * 4;   // bytes along K inside the 128-byte swizzle atom
                        const uint64_t ahi = umma_desc_k128(base + koff), alo = umma_desc_k128(base + A_BYTES + koff);
                        const uint64_t bhi = umma_desc_k128(base + 2 * A_BYTES + koff);
                        const uint64_t blo = umma_desc_k128(base + 2 * A_BYTES + B_BYTES + koff);
                        umma_tf32(d_tmem, ahi, bhi, idesc, (kb | ks) != 0);
                        umma_tf32(d_tmem, alo, bhi, idesc, 1);
                        umma_tf32(d_tmem, ahi, blo, idesc, 1);
                    }
                    umma_commit(empty0 + 8 * s);        // shared-memory stage reusable once these MMAs retire
                }
                umma_commit(tfull0 + 8 * buf);          // accumulator tile complete
            }
        }
        __syncwarp();
    } else {
        // ================================================= epilogue warps ===============================================
        const int quarter = warp & 3;                   // TMEM lanes [32 * quarter, 32 * quarter + 32)
        const int row = quarter * 32 + lane;            // query row of the tile owned by this thread
        uint2 *my_queue = bp.tie_queue + (size_t)blockIdx.x * bp.tie_cap;
        uint32_t *mask = mask_all + (warp - EPI_WARP0) * 32 * MASK_STRIDE;
        // Per-tile metadata (tile geometry, this row's true score and guard, the tile's known-true pair range) is
        // fetched ONE TILE AHEAD: a lone epilogue warp per SM sub-partition cannot hide dependent global-load latency,
        // and the epilogue of tile t must finish before the MMA warp may start tile t + 2.
        struct TileMeta {
            int64_t qbase, crow0;
            int nq, ne;
            float sim_true, guard;
            uint32_t pf0, pf1;
        };
        const GroupDesc gd0 = p.groups[0];
        auto load_meta = [&](int64_t item) {
            TileMeta m;
            int g = 0, qt, et;
            GroupDesc gd = gd0;
            if (p.n_groups > 1) {
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
            m.pf0 = __ldg(p.tf_ptr + item);
            m.pf1 = __ldg(p.tf_ptr + item + 1);
            return m;
        };
        uint32_t tile = 0;
        TileMeta nxt{};
        if ((int64_t)blockIdx.x < p.total_items) nxt = load_meta(blockIdx.x);
        for (int64_t item = blockIdx.x; item < p.total_items; item += gridDim.x, tile++) {
            const TileMeta cur = nxt;
            if (item + gridDim.x < p.total_items) nxt = load_meta(item + gridDim.x);
            const int64_t qbase = cur.qbase;
            const int nq = cur.nq, ne = cur.ne;
            const bool q_ok = row < nq;
            const float sim_true = cur.sim_true, guard = cur.guard;
            // known-true mask of this warp's 32 rows: the pairs tile_filter.cu routed to this item (the truth included)
            const uint32_t pf0 = cur.pf0, pf1 = cur.pf1;
            const bool has_mask = pf1 > pf0;
            if (has_mask) {
                for (int k = lane; k < 32 * MASK_STRIDE; k += 32) mask[k] = 0u;
                __syncwarp();
                for (uint32_t k = pf0 + lane; k < pf1; k += 32) {
                    const uint32_t pr = __ldg(p.tf_pairs + k);
                    const int prow = (int)(pr >> 16), pcol = (int)(pr & 0xffffu);
                    if ((prow >> 5) == quarter) atomicOr(&mask[(prow & 31) * MASK_STRIDE + (pcol >> 5)], 1u << (pcol & 31));
                }
                __syncwarp();
            }
            const uint32_t buf = tile & 1;
            mbar_wait(tfull0 + 8 * buf, (tile >> 1) & 1);
            tc_fence_after();
            int r_lt = 0, f_lt = 0, r_eq = 0, f_eq = 0;      // raw / filtered counts of this thread's query row
            const uint32_t taddr = tmem_base + buf * BN + ((uint32_t)(quarter * 32) << 16);
            // better <=> larger similarity (predict = -sim).  s > thr_hi: counted; s < thr_lo: not; in between: near-tie
            const float thr_hi = sim_true + guard, thr_lo = sim_true - guard;
            // One chunk = 32 accumulator columns of this thread's row.  The two comparisons are collected as BIT MASKS, so
            // the counts, the known-true subtraction and the near-tie set are a handful of popc / and instructions and the
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
                uint32_t gtm = 0u, gem = 0u;                    // bit c: column c0 + c beats thr_hi / reaches thr_lo
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    const float sc = __uint_as_float(v[c]);
                    gtm |= sc > thr_hi ? (1u << c) : 0u;
                    gem |= sc >= thr_lo ? (1u << c) : 0u;
                }
                const uint32_t valid = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);   // padding of the last candidate tile
                gtm &= valid;
                gem &= valid;
                const uint32_t mw = has_mask ? mask[lane * MASK_STRIDE + (c0 >> 5)] : 0u;   // known-true columns of this chunk
                r_lt += __popc(gtm);
                f_lt += __popc(gtm & ~mw);
                uint32_t near = gem & ~gtm;                     // rare: queue the near-ties for the exact re-score
                while (near) {
                    const int c = __ffs(near) - 1;
                    near &= near - 1;
                    const uint32_t kn = (mw >> c) & 1u;
                    const int64_t crow = cur.crow0 + c0 + c;
                    const int64_t ent_id = p.all_entities ? crow : __ldg(p.cand_idx + crow);
                    const uint32_t slot = atomicAdd(tie_count, 1u);
                    if (slot < bp.tie_cap) {
                        my_queue[slot] = make_uint2((uint32_t)(qbase + row), (uint32_t)ent_id | (kn << 31));
                    } else {                                    // queue full (pathological ties): re-score in place
                        const float s2 = bil_dot(p.qvec + (qbase + row) * p.D, p.ent + ent_id * p.D, p.D);
                        r_lt += s2 > sim_true ? 1 : 0;
                        r_eq += s2 == sim_true ? 1 : 0;
                        f_lt += (!kn && s2 > sim_true) ? 1 : 0;
                        f_eq += (!kn && s2 == sim_true) ? 1 : 0;
                    }
                }
            };
            // two register buffers: the TMEM load of the next 32 columns is in flight while this chunk is compared
            uint32_t va[32], vb[32];
            tmem_ld_32x32(taddr, va);
#pragma unroll 1
            for (int c0 = 0; c0 < ne; c0 += 64) {               // ne is warp-uniform: the collective loads stay aligned
                tmem_ld_wait();
                if (c0 + 32 < ne) tmem_ld_32x32(taddr + c0 + 32, vb);
                process(va, c0);
                if (c0 + 32 < ne) {
                    tmem_ld_wait();
                    if (c0 + 64 < ne) tmem_ld_32x32(taddr + c0 + 64, va);
                    process(vb, c0 + 32);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
            if (q_ok) {
                const int64_t q = qbase + row;
                if (r_lt) atomicAdd(p.counts + q, r_lt);
                if (r_eq) atomicAdd(p.counts + p.Q + q, r_eq);
                if (f_lt) atomicAdd(p.counts + 2 * p.Q + q, f_lt);
                if (f_eq) atomicAdd(p.counts + 3 * p.Q + q, f_eq);
            }
        }
        // exact FP32 re-score of this CTA's queued near-ties: off the tile loop's critical path, one per thread, so the
        // row fetches of ~128 items are in flight together.  Known-true entities only enter the RAW counts.
        asm volatile("bar.sync 1, 128;" ::: "memory");          // the four epilogue warps have finished pushing
        if (!STORE) {
            const uint32_t n_tie = min(*tie_count, bp.tie_cap);
            for (uint32_t i = (warp - EPI_WARP0) * 32 + lane; i < n_tie; i += 128) {
                const uint2 it2 = my_queue[i];
                const int64_t q2 = it2.x, ent_id = it2.y & 0x7fffffffu;
                const bool kn = (it2.y >> 31) != 0u;
                const float s2 = bil_dot(p.qvec + q2 * p.D, p.ent + ent_id * p.D, p.D);
                const float st = -__ldg(&p.thr[q2].x);
                if (s2 > st) { atomicAdd(p.counts + q2, 1); if (!kn) atomicAdd(p.counts + 2 * p.Q + q2, 1); }
                if (s2 == st) { atomicAdd(p.counts + p.Q + q2, 1); if (!kn) atomicAdd(p.counts + 3 * p.Q + q2, 1); }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------ TF32 MMA peak probe
__global__ void __launch_bounds__(128, 1) tf32_probe_kernel(int iters) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *al = smem_raw + (base - smem_u32(smem_raw));
    uint64_t *bar = reinterpret_cast<uint64_t *>(al + A_BYTES + B_BYTES);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
    for (uint32_t i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) reinterpret_cast<float *>(al)[i] = 0.f;
    if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); fence_barrier_init(); }
    fence_proxy_async();
    if (threadIdx.x < 32) tmem_alloc(smem_u32(slot), 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = umma_idesc_tf32(BM, BN);
        for (int i = 0; i < iters; i++) {
            const uint32_t koff = (i & 3) * UK * 4;
            umma_tf32(tmem, umma_desc_k128(base + koff), umma_desc_k128(base + A_BYTES + koff), idesc, i != 0);
        }
        umma_commit(smem_u32(bar));
        mbar_wait(smem_u32(bar), 0);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int probe_tf32_peak(mre_ctx *ctx, double *flops_per_s) {
    MRE_CHECK_ARG(flops_per_s != nullptr, "NULL output");
    const size_t smem = 1024 + A_BYTES + B_BYTES + 64;
    MRE_CUDA(cudaFuncSetAttribute(tf32_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 8192;
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        MRE_CUDA(cudaEventRecord(ctx->ev0, 0));
        tf32_probe_kernel<<<ctx->sm_count, 128, smem>>>(iters);
        MRE_CUDA(cudaEventRecord(ctx->ev1, 0));
        MRE_CUDA(cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        MRE_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double flops = (double)ctx->sm_count * iters * 2.0 * BM * BN * UK;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3));
    }
    ctx->launches += 4;
    MRE_CUDA(cudaGetLastError());
    *flops_per_s = best;
    return MRE_OK;
}

// ------------------------------------------------------------------------------------------ host side
struct BilScratch {
    const float *ent_full;   // [E, Kp]
    const float *ent_hi, *ent_lo;
    int64_t K, Kp;
};

static int bil_prepass(mre_ctx *ctx, const mre_rank_job *job, cudaStream_t st, BilScratch &sc) {
    const int64_t D = job->D;
    sc.K = job->scorer == MRE_COMPLEX ? 2 * D : D;
    sc.Kp = (sc.K + 3) & ~(int64_t)3;
    const size_t tbl = (size_t)job->E * sc.Kp * sizeof(float);
    // layout of ctx->ent_n: [hi | lo | full (only when a repacked full-precision copy is needed)]
    const bool need_full = job->scorer == MRE_COMPLEX || sc.Kp != D;
    MRE_TRY(ctx->ent_n.reserve(tbl * (need_full ? 3 : 2)));
    float *hi = ctx->ent_n.as<float>(), *lo = hi + (size_t)job->E * sc.Kp;
    float *full = need_full ? lo + (size_t)job->E * sc.Kp : nullptr;
    bil_split_table_kernel<<<grid_for(job->E * sc.Kp, 256), 256, 0, st>>>(job->ent, job->ent_im, job->E, D, sc.K, sc.Kp, full, hi, lo);
    ctx->launches += 1;
    sc.ent_full = need_full ? full : job->ent;
    sc.ent_hi = hi;
    sc.ent_lo = lo;
    if (job->Q > 0) {
        const size_t qb = (size_t)job->Q * sc.Kp * sizeof(float);
        MRE_TRY(ctx->qvec.reserve(qb));
        MRE_TRY(ctx->qvec2.reserve(2 * qb));
        bil_qvec_kernel<<<grid_for(job->Q * sc.Kp, 256), 256, 0, st>>>(job->scorer, job->ent, job->ent_im, job->rel, job->rel_im, D, sc.K,
                                                                     sc.Kp, job->q_h, job->q_t, job->q_r, job->q_side, job->side, job->Q,
                                                                     ctx->qvec.as<float>(), ctx->qvec2.as<float>(),
                                                                     ctx->qvec2.as<float>() + (size_t)job->Q * sc.Kp);
        ctx->launches += 1;
    }
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

static int run_bilinear(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, float *store, cudaStream_t st) {
    BilScratch sc{};
    MRE_TRY(bil_prepass(ctx, job, st, sc));
    BilParams bp{};
    RankParams &p = bp.r;
    MRE_TRY(fill_rank_params(ctx, ix, job, BM, BN, st, p));
    p.ent = sc.ent_full;
    p.D = sc.Kp;
    p.qvec = ctx->qvec.as<float>();
    if (job->Q == 0) return MRE_OK;
    MRE_TRY(ctx->thr.reserve((size_t)job->Q * (sizeof(float2) + sizeof(float)) + 16));
    float2 *thr = ctx->thr.as<float2>();
    float *delta = reinterpret_cast<float *>(thr + job->Q);
    unsigned int *max_norm = reinterpret_cast<unsigned int *>(delta + job->Q);
    p.thr = thr;
    bp.delta = delta;
    MRE_CUDA(cudaMemsetAsync(max_norm, 0, sizeof(unsigned int), st));
    bil_max_rownorm_kernel<<<grid_for(job->E * 32, 256), 256, 0, st>>>(sc.ent_full, job->E, sc.Kp, max_norm);
    bil_threshold_kernel<<<(unsigned)((job->Q + 127) / 128), 128, 0, st>>>(p, max_norm, thr, delta);
    ctx->launches += 2;
    // candidate tables the B tiles stream from
    const float *b_hi = sc.ent_hi, *b_lo = sc.ent_lo;
    int64_t cand_rows = job->E;
    if (!p.all_entities) {
        cand_rows = job->group_cptr[job->n_groups];
        MRE_CHECK_ARG(cand_rows < (1LL << 31), "too many candidate rows");
        const size_t cb = (size_t)std::max<int64_t>(cand_rows, 1) * sc.Kp * sizeof(float);
        MRE_TRY(ctx->ent_aux.reserve(2 * cb));
        float *g_hi = ctx->ent_aux.as<float>(), *g_lo = g_hi + (size_t)std::max<int64_t>(cand_rows, 1) * sc.Kp;
        if (cand_rows > 0) {
            gather_rows_kernel<<<grid_for(cand_rows * (sc.Kp >> 2), 256), 256, 0, st>>>(sc.ent_hi, sc.Kp, job->cand_idx, cand_rows, g_hi);
            gather_rows_kernel<<<grid_for(cand_rows * (sc.Kp >> 2), 256), 256, 0, st>>>(sc.ent_lo, sc.Kp, job->cand_idx, cand_rows, g_lo);
            ctx->launches += 2;
        }
        b_hi = g_hi;
        b_lo = g_lo;
    }
    init_counts_kernel<<<grid_for(4 * job->Q, 256), 256, 0, st>>>(job->counts, 4 * job->Q);
    ctx->launches += 1;
    MRE_TRY(build_tile_filter(ctx, job, p, BM, BN, st));
    const float *q_hi = ctx->qvec2.as<float>(), *q_lo = q_hi + (size_t)job->Q * sc.Kp;
    CUtensorMap tm_ahi, tm_alo, tm_bhi, tm_blo;
    MRE_TRY(make_tmap_f32_2d(&tm_ahi, q_hi, job->Q, sc.Kp, sc.Kp, BM, BK));
    MRE_TRY(make_tmap_f32_2d(&tm_alo, q_lo, job->Q, sc.Kp, sc.Kp, BM, BK));
    MRE_TRY(make_tmap_f32_2d(&tm_bhi, b_hi, std::max<int64_t>(cand_rows, 1), sc.Kp, sc.Kp, BN, BK));
    MRE_TRY(make_tmap_f32_2d(&tm_blo, b_lo, std::max<int64_t>(cand_rows, 1), sc.Kp, sc.Kp, BN, BK));
    static bool configured = false;
    if (!configured) {
        MRE_CUDA(cudaFuncSetAttribute(bilinear_rank_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BIL_SMEM));
        MRE_CUDA(cudaFuncSetAttribute(bilinear_rank_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BIL_SMEM));
        configured = true;
    }
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(p.total_items, ctx->sm_count));
    bp.store = store;
    bp.store_ld = cand_rows;
    // near-tie queue: expected Q*E*P(near) entries with P(near) ~ 5e-5; 10x headroom, at least 4096 per CTA
    const int64_t cap_total = std::min<int64_t>(std::max<int64_t>(job->Q * std::max<int64_t>(cand_rows, 1) / 2048, (int64_t)grid * 4096), 1LL << 26);
    bp.tie_cap = (uint32_t)(cap_total / grid);
    MRE_TRY(ctx->tie_queue.reserve((size_t)bp.tie_cap * grid * sizeof(uint2)));
    bp.tie_queue = ctx->tie_queue.as<uint2>();
    MRE_TRY(ctx->time_begin(st));
    if (store) bilinear_rank_kernel<true><<<grid, BIL_THREADS, BIL_SMEM, st>>>(bp, tm_ahi, tm_alo, tm_bhi, tm_blo);
    else bilinear_rank_kernel<false><<<grid, BIL_THREADS, BIL_SMEM, st>>>(bp, tm_ahi, tm_alo, tm_bhi, tm_blo);
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
    MRE_TRY(bil_prepass(ctx, &one, st, sc));
    bil_predict_kernel<<<(unsigned)((job->E + 127) / 128), 128, 0, st>>>(sc.ent_full, job->E, sc.Kp, ctx->qvec.as<float>(), scores_out);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

}  // namespace mre
