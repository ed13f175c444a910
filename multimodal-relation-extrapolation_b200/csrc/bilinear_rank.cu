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
// back with tcgen05.ld (one query row per thread), compare against the query's true score and count; columns
// closer to the true score than a rigorous error guard are re-scored in scalar FP32, which makes the COUNTS exactly
// those of the sequential FP32 scorer (and consistent with the known-true correction).
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
constexpr size_t BIL_SMEM = 1024 + (size_t)B_STAGES * BSTAGE_BYTES + 16 * sizeof(uint64_t) + 64;
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
                    tma_load_2d(base, &tm_ahi, kb * BK, qrow, full);
                    tma_load_2d(base + A_BYTES, &tm_alo, kb * BK, qrow, full);
                    tma_load_2d(base + 2 * A_BYTES, &tm_bhi, kb * BK, erow, full);
                    tma_load_2d(base + 2 * A_BYTES + B_BYTES, &tm_blo, kb * BK, erow, full);
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
        uint32_t tile = 0;
        for (int64_t item = blockIdx.x; item < p.total_items; item += gridDim.x, tile++) {
            int g, qt, et;
            decode_item(p, item, g, qt, et);
            const GroupDesc gd = p.groups[g];
            const int64_t qbase = gd.q0 + (int64_t)qt * BM;
            const int nq = (int)min((int64_t)BM, gd.q0 + gd.nq - qbase);
            const int ne = (int)min((int64_t)BN, gd.nc - (int64_t)et * BN);
            const bool q_ok = row < nq;
            float sim_true = INFINITY, guard = -1.f;
            if (q_ok) {
                sim_true = -__ldg(&p.thr[qbase + row].x);
                guard = __ldg(bp.delta + qbase + row);
            }
            const uint32_t buf = tile & 1;
            mbar_wait(tfull0 + 8 * buf, (tile >> 1) & 1);
            tc_fence_after();
            int n_lt = 0, n_eq = 0;
            const uint32_t taddr = tmem_base + buf * BN + ((uint32_t)(quarter * 32) << 16);
            // better <=> larger similarity (predict = -sim).  s > thr_hi: counted; s < thr_lo: not; in between: near-tie
            const float thr_hi = sim_true + guard, thr_lo = sim_true - guard;
            auto process = [&](const uint32_t (&v)[32], int c0) {
                const int left = ne - c0;                       // warp-uniform
                if (STORE) {                                    // diagnostic / materialised-score mode
                    if (q_ok) {
                        float *o = bp.store + (qbase + row) * bp.store_ld + gd.c0 + (int64_t)et * BN + c0;
#pragma unroll
                        for (int c = 0; c < 32; c++)
                            if (c < left) o[c] = __uint_as_float(v[c]);
                    }
                    return;
                }
                int gt = 0, ge = 0;
                if (left >= 32) {
#pragma unroll
                    for (int c = 0; c < 32; c++) {
                        const float sc = __uint_as_float(v[c]);
                        gt += sc > thr_hi ? 1 : 0;
                        ge += sc >= thr_lo ? 1 : 0;
                    }
                } else {                                        // last chunk of the last candidate tile: mask the padding
#pragma unroll
                    for (int c = 0; c < 32; c++) {
                        const float sc = __uint_as_float(v[c]);
                        gt += (c < left && sc > thr_hi) ? 1 : 0;
                        ge += (c < left && sc >= thr_lo) ? 1 : 0;
                    }
                }
                n_lt += gt;
                if (ge != gt) {                                 // rare: queue the near-ties of this chunk for the exact re-score
#pragma unroll
                    for (int c = 0; c < 32; c++) {
                        const float sc = __uint_as_float(v[c]);
                        if (c < left && sc >= thr_lo && !(sc > thr_hi)) {
                            const int64_t crow = gd.c0 + (int64_t)et * BN + c0 + c;
                            const int64_t ent_id = p.all_entities ? crow : __ldg(p.cand_idx + crow);
                            const uint32_t slot = atomicAdd(tie_count, 1u);
                            if (slot < bp.tie_cap) {
                                my_queue[slot] = make_uint2((uint32_t)(qbase + row), (uint32_t)ent_id);
                            } else {                            // queue full (pathological ties): re-score in place
                                const float s2 = bil_dot(p.qvec + (qbase + row) * p.D, p.ent + ent_id * p.D, p.D);
                                n_lt += s2 > sim_true ? 1 : 0;
                                n_eq += s2 == sim_true ? 1 : 0;
                            }
                        }
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
                if (n_lt) { atomicAdd(p.counts + q, n_lt); atomicAdd(p.counts + 2 * p.Q + q, n_lt); }
                if (n_eq) { atomicAdd(p.counts + p.Q + q, n_eq); atomicAdd(p.counts + 3 * p.Q + q, n_eq); }
            }
        }
        // exact FP32 re-score of this CTA's queued near-ties: off the tile loop's critical path, one per thread, so the
        // row fetches of ~128 items are in flight together
        asm volatile("bar.sync 1, 128;" ::: "memory");          // the four epilogue warps have finished pushing
        if (!STORE) {
            const uint32_t n_tie = min(*tie_count, bp.tie_cap);
            for (uint32_t i = (warp - EPI_WARP0) * 32 + lane; i < n_tie; i += 128) {
                const uint2 it2 = my_queue[i];
                const int64_t q2 = it2.x, ent_id = it2.y;
                const float s2 = bil_dot(p.qvec + q2 * p.D, p.ent + ent_id * p.D, p.D);
                const float st = -__ldg(&p.thr[q2].x);
                if (s2 > st) { atomicAdd(p.counts + q2, 1); atomicAdd(p.counts + 2 * p.Q + q2, 1); }
                if (s2 == st) { atomicAdd(p.counts + p.Q + q2, 1); atomicAdd(p.counts + 3 * p.Q + q2, 1); }
            }
        }
        // known-true correction with the same FP32 scalar scorer the near-tie path uses (consistent decisions)
        if (!STORE) {
            const int64_t n_warps = (int64_t)gridDim.x * 4;
            for (int64_t q = (int64_t)blockIdx.x * 4 + (warp - EPI_WARP0); q < p.Q; q += n_warps)
                correct_query<true, true>(p, q, lane, [&](int64_t qq, int64_t x) {
                    return -bil_dot(p.qvec + qq * p.D, p.ent + x * p.D, p.D);
                });
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
    const int64_t want = std::max<int64_t>(p.total_items, (p.Q + 3) / 4);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, ctx->sm_count));
    bp.store = store;
    bp.store_ld = cand_rows;
    // near-tie queue: expected Q*E*P(near) entries with P(near) ~ 5e-5; 10x headroom, at least 4096 per CTA
    const int64_t cap_total = std::min<int64_t>(std::max<int64_t>(job->Q * std::max<int64_t>(cand_rows, 1) / 2048, (int64_t)grid * 4096), 1LL << 26);
    bp.tie_cap = (uint32_t)(cap_total / grid);
    MRE_TRY(ctx->counters.reserve((size_t)bp.tie_cap * grid * sizeof(uint2)));
    bp.tie_queue = ctx->counters.as<uint2>();
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
