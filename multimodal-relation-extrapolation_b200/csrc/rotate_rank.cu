// RotatE fused score + rank (a sibling of the TransE kernel: SURVEY 8 f4).
//
// Reference path replaced (paths relative to /root/reference):
//   OpenKE/openke/module/model/RotatE.py:44-78   _calc: phase = r / (rel_embedding_range / pi); the fixed entity rotated by the
//                                                relation (tail query: h o r; head query: conj(r) o t), minus the candidate,
//                                                complex modulus per dimension, summed
//   OpenKE/openke/module/model/RotatE.py:80-91   forward = margin - score, predict = -forward = score - margin (lower = better)
//   OpenKE/openke/base/Test.h:65-192             testHead / testTail on the E-long score vector
// The margin is a constant shift of every score of a query: the counts do not depend on it, so the kernel ranks the distance
// itself.  Per (query, entity, complex dimension): two subtractions, m = fma(di, di, dr * dr), one square root (MUFU-bound: one
// MUFU.SQRT -- 16 per clock per SM -- per complex dimension beside five FP32 instructions) and one add, sequential over
// d per pair -- the same accumulation in the tile kernel, in the threshold of the true entity and in the known-true correction
// pass (rotate_acc), so `s_j < s_true` is decided on identical bits.  A plain shared-memory tile kernel (64 queries x 64
// candidates per CTA, 4 x 4 per thread), not the TMA / packed-FADD2 machinery of the TransE kernel.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.h"
#include "rank_common.cuh"
#include "rank_host.h"

namespace mre {

constexpr int RT_T = 64;          // queries and candidates per tile
constexpr int RT_CH = 16;         // complex dimensions per shared-memory chunk
constexpr int RT_THREADS = 256;   // 16 x 16 threads, a 4 x 4 micro-tile each

// the one definition of a RotatE accumulator: v = [v_re | v_im], e = [e_re | e_im] (Dc complex dimensions each)
#ifndef MRE_ROTATE_SQRT_APPROX
#define MRE_ROTATE_SQRT_APPROX 1
#endif
__device__ __forceinline__ float rotate_step(float acc, float vr, float vi, float er, float ei) {
    const float dr = vr - er, di = vi - ei;
    const float m = fmaf(di, di, dr * dr);
#if MRE_ROTATE_SQRT_APPROX
    // one MUFU.SQRT (2 ulp) instead of MUFU.RSQ + the Newton / special-case sequence of the IEEE square root: the tile loop is then
    // bound by the MUFU unit (16 per clock per SM) rather than by a dozen FP32 instructions per element.  Every decision of a job
    // -- tile kernel, threshold, known-true pass -- goes through this one function, so the counts stay self-consistent; against the
    // reference's torch.norm the difference (~1e-7 relative on a sum of dim terms) is far inside the 1e-5 tie band.
    float s;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(m));
    return acc + s;
#else
    return acc + sqrtf(m);
#endif
}
// two candidates at once, packed (sub / mul / fma / add .f32x2: the same IEEE roundings per lane, half the FP32 instructions --
// at the MUFU bound the scalar form keeps 82 % of the issue slots busy)
__device__ __forceinline__ float rotate_sqrt(float m) {
    float s;
#if MRE_ROTATE_SQRT_APPROX
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(m));
#else
    s = sqrtf(m);
#endif
    return s;
}
__device__ __forceinline__ void rotate_step2(uint64_t &acc, float vr, float vi, uint64_t er, uint64_t ei) {
    uint64_t dr, di, m, sq;
    asm("{\n\t.reg .b64 t;\n\tmov.b64 t, {%1, %1};\n\tsub.rn.f32x2 %0, t, %2;\n\t}" : "=l"(dr) : "f"(vr), "l"(er));
    asm("{\n\t.reg .b64 t;\n\tmov.b64 t, {%1, %1};\n\tsub.rn.f32x2 %0, t, %2;\n\t}" : "=l"(di) : "f"(vi), "l"(ei));
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(m) : "l"(dr));
    asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(m) : "l"(di));
    float m0, m1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(m0), "=f"(m1) : "l"(m));
    const float s0 = rotate_sqrt(m0), s1 = rotate_sqrt(m1);
    asm("mov.b64 %0, {%1, %2};" : "=l"(sq) : "f"(s0), "f"(s1));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(sq));
}
__device__ __forceinline__ float rotate_acc(const float *__restrict__ v, const float *__restrict__ e, int Dc) {
    float acc = 0.f;
    for (int d = 0; d < Dc; d++) acc = rotate_step(acc, v[d], v[Dc + d], __ldg(e + d), __ldg(e + Dc + d));
    return acc;
}

// per query (one thread each): the rotated fixed entity v (RotatE.py:61-71), the true entity's distance with the tile kernel's
// accumulation, the thresholds (lo = s_true, hi = the next float: eq <=> acc == s_true), and the zeroed counters
__global__ void __launch_bounds__(128) rotate_query_kernel(const RankParams p, const float *__restrict__ rel, float phase_div,
                                                           float *__restrict__ qvec, float2 *__restrict__ thr) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= p.Q) return;
    const int side = query_side(p, q);
    const int Dc = (int)(p.D >> 1);
    const int64_t h = p.q_h[q], t = p.q_t[q];
    const float *f = p.ent + (side ? h : t) * p.D, *r = rel + p.q_r[q] * Dc, *e = p.ent + (side ? t : h) * p.D;
    float *v = qvec + q * p.D;
    float acc = 0.f;
    for (int d = 0; d < Dc; d++) {
        const float ph = r[d] / phase_div;
        const float c = cosf(ph), s = sinf(ph);
        const float fr = f[d], fi = f[Dc + d];
        float vr, vi;
        if (side) { vr = fr * c - fi * s; vi = fr * s + fi * c; }      // tail query: h o r          (RotatE.py:67-68)
        else { vr = c * fr + s * fi; vi = c * fi - s * fr; }            // head query: conj(r) o t    (RotatE.py:62-63)
        v[d] = vr;
        v[Dc + d] = vi;
        acc = rotate_step(acc, vr, vi, e[d], e[Dc + d]);
    }
    float hi = acc;
    if (acc == acc && fabsf(acc) < INFINITY) hi = nextafterf(acc, INFINITY);
    thr[q] = make_float2(acc, hi);
#pragma unroll
    for (int c = 0; c < 4; c++) p.counts[(int64_t)c * p.Q + q] = 0;
}

__global__ void __launch_bounds__(RT_THREADS) rotate_rank_kernel(const RankParams p) {
    __shared__ float sQ[2][RT_CH][RT_T + 4], sE[2][RT_CH][RT_T + 4];     // [re | im][d][row]
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int Dc = (int)(p.D >> 1);
    for (int64_t item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        int g, qt, et;
        decode_item(p, item, g, qt, et);
        const GroupDesc gd = p.groups[g];
        const int64_t qbase = gd.q0 + (int64_t)qt * RT_T, cbase = gd.c0 + (int64_t)et * RT_T;
        const int nq = (int)min((int64_t)RT_T, gd.q0 + gd.nq - qbase), ne = (int)min((int64_t)RT_T, gd.nc - (int64_t)et * RT_T);
        uint64_t acc2[4][2];                                       // [query][candidate pair]: two FP32 accumulators per register pair
#pragma unroll
        for (int i = 0; i < 4; i++) acc2[i][0] = acc2[i][1] = 0ull;
        // this thread stages row (threadIdx.x / 4) of both operands, dimensions 4 (threadIdx.x % 4) .. + 3 of every chunk
        const int lrow = threadIdx.x >> 2, ld4 = (threadIdx.x & 3) * 4;
        const float *qrow = lrow < nq ? p.qvec + (qbase + lrow) * p.D : nullptr;
        const float *erow = nullptr;
        if (lrow < ne) erow = p.ent + (p.all_entities ? cbase + lrow : __ldg(p.cand_idx + cbase + lrow)) * p.D;
        for (int c0 = 0; c0 < Dc; c0 += RT_CH) {
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int d = c0 + ld4 + k;
                const bool ok = d < Dc;
                sQ[0][ld4 + k][lrow] = (qrow && ok) ? qrow[d] : 0.f;
                sQ[1][ld4 + k][lrow] = (qrow && ok) ? qrow[Dc + d] : 0.f;
                sE[0][ld4 + k][lrow] = (erow && ok) ? __ldg(erow + d) : 0.f;
                sE[1][ld4 + k][lrow] = (erow && ok) ? __ldg(erow + Dc + d) : 0.f;
            }
            __syncthreads();
            const int nd = min(RT_CH, Dc - c0);
            for (int d = 0; d < nd; d++) {
                const float4 qr = *reinterpret_cast<const float4 *>(&sQ[0][d][ty * 4]), qi = *reinterpret_cast<const float4 *>(&sQ[1][d][ty * 4]);
                const ulonglong2 er = *reinterpret_cast<const ulonglong2 *>(&sE[0][d][tx * 4]), ei = *reinterpret_cast<const ulonglong2 *>(&sE[1][d][tx * 4]);
                const float vr[4] = {qr.x, qr.y, qr.z, qr.w}, vi[4] = {qi.x, qi.y, qi.z, qi.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    rotate_step2(acc2[i][0], vr[i], vi[i], er.x, ei.x);
                    rotate_step2(acc2[i][1], vr[i], vi[i], er.y, ei.y);
                }
            }
        }
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int jp = 0; jp < 2; jp++)
                asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i][2 * jp]), "=f"(acc[i][2 * jp + 1]) : "l"(acc2[i][jp]));
        // ---- compare + count: raw and filtered counters alike (the known-true correction pass takes the known entities back out)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int ql = ty * 4 + i;
            const float2 th = ql < nq ? p.thr[qbase + ql] : make_float2(-INFINITY, -INFINITY);
            int lt = 0, eq = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bool ok = tx * 4 + j < ne;
                const bool l = ok && acc[i][j] < th.x;
                lt += l ? 1 : 0;
                eq += (ok && !l && acc[i][j] < th.y) ? 1 : 0;
            }
            int packed = lt | (eq << 16);
#pragma unroll
            for (int m = 1; m < 16; m <<= 1) packed += __shfl_xor_sync(0xffffffffu, packed, m);
            if (tx == 0 && ql < nq && packed) {
                const int64_t q = qbase + ql;
                const int n_lt = packed & 0xffff, n_eq = packed >> 16;
                if (n_lt) { atomicAdd(p.counts + q, n_lt); atomicAdd(p.counts + 2 * p.Q + q, n_lt); }
                if (n_eq) { atomicAdd(p.counts + p.Q + q, n_eq); atomicAdd(p.counts + 3 * p.Q + q, n_eq); }
            }
        }
    }
}

// the known-true correction (rank_common.cuh) with the RotatE accumulator; every segment takes the entry-per-lane path
struct RotateKnownOp {
    const float *ent, *qvec, *v;
    const float2 *thr;
    int64_t D;
    float2 th;
    static constexpr bool DIRECT_ONLY = true;
    __device__ __forceinline__ void query(int64_t q, int, int64_t, int64_t, int64_t) { v = qvec + q * D; }
    __device__ __forceinline__ void thresholds(int64_t q) { th = thr[q]; }
    __device__ __forceinline__ bool truth_ties() const { return th.x < th.y; }
    __device__ __forceinline__ float direct(int64_t x) const { return rotate_acc(v, ent + x * D, (int)(D >> 1)); }
    __device__ __forceinline__ float vec(int) const { return 0.f; }
    __device__ __forceinline__ float term(float, float) const { return 0.f; }
    __device__ __forceinline__ float fold(float acc, float) const { return acc; }
    __device__ __forceinline__ void classify(float acc, int &lt, int &eq) const {
        if (acc < th.x) lt++;
        else if (acc < th.y) eq++;
    }
};
__global__ void __launch_bounds__(KNOWN_WARPS * 32) rotate_known_score_kernel(const RankParams p, const KnownRuns kr) {
    __shared__ float sT[1][32][33];          // unused by the entry-per-lane path
    __shared__ int64_t sX[1][32];
    RotateKnownOp op{p.ent, p.qvec, nullptr, p.thr, p.D, make_float2(0.f, 0.f)};
    known_score_runs(p, kr, op, sT[0], sX[0]);
}
__global__ void __launch_bounds__(KNOWN_WARPS * 32) rotate_known_compare_kernel(const RankParams p, const KnownRuns kr) {
    RotateKnownOp op{p.ent, p.qvec, nullptr, p.thr, p.D, make_float2(0.f, 0.f)};
    known_compare_runs(p, kr, op);
}
// lists that do not come from the index (MRE_FILTER_CSR, or none: only the true entities): one warp per query, one entry per lane;
// the list walk of known_correction (rank_common.cuh) -- distinct entries of the sorted slice that lie in the query's candidate
// set, the true entity last -- with every survivor scored by the scalar accumulator
__global__ void __launch_bounds__(KNOWN_WARPS * 32) rotate_known_direct_kernel(const RankParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * KNOWN_WARPS + (threadIdx.x >> 5);
    if (q >= p.Q) return;
    const int side = query_side(p, q);
    const int64_t truth = side ? p.q_t[q] : p.q_h[q];
    const int64_t *list = p.filt_idx;
    int64_t lo = 0, hi = 0;
    if (p.filter == MRE_FILTER_CSR) {
        lo = p.filt_ptr[q];
        hi = p.filt_ptr[q + 1];
    }
    const GroupDesc &gd = p.groups[p.all_entities ? 0 : group_of_query(p, q)];
    if (q < gd.q0 || q - gd.q0 >= gd.nq) return;            // the query's own group was empty (dropped)
    RotateKnownOp op{p.ent, p.qvec, nullptr, p.thr, p.D, make_float2(0.f, 0.f)};
    op.query(q, 0, 0, 0, 0);
    op.thresholds(q);
    int k_lt = 0, k_eq = 0;
    for (int64_t i = lo + lane; i <= hi; i += 32) {
        int64_t x = i < hi ? __ldg(list + i) : truth;
        if (i < hi && (x == truth || (i > lo && __ldg(list + i - 1) == x))) continue;     // the truth goes last; duplicates once
        if (x < 0 || x >= p.E) continue;
        if (!p.all_entities) {
            const int64_t k = lower_bound_i64(p.cand_idx, gd.c0, gd.c0 + gd.nc, x);
            if (k >= gd.c0 + gd.nc || __ldg(p.cand_idx + k) != x) continue;
        }
        op.classify(op.direct(x), k_lt, k_eq);
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

int rank_rotate(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, cudaStream_t st) {
    MRE_CHECK_ARG(job->D % 2 == 0, "RotatE entity rows hold [re | im]: D must be even");
    MRE_CHECK_ARG(job->rotate_phase_div > 0.f, "rotate_phase_div (rel_embedding_range / pi) must be positive");
    RankParams p{};
    MRE_TRY(fill_rank_params(ctx, ix, job, RT_T, RT_T, st, p));
    p.ent = job->ent;
    p.D = job->D;
    if (job->Q == 0) return MRE_OK;
    MRE_TRY(ctx->qvec.reserve((size_t)p.Q * p.D * sizeof(float)));
    MRE_TRY(ctx->thr.reserve((size_t)p.Q * sizeof(float2)));
    p.qvec = ctx->qvec.as<float>();
    p.thr = ctx->thr.as<float2>();
    rotate_query_kernel<<<(unsigned)((p.Q + 127) / 128), 128, 0, st>>>(p, job->rel, job->rotate_phase_div, ctx->qvec.as<float>(), ctx->thr.as<float2>());
    const unsigned qgrid = (unsigned)((p.Q + KNOWN_WARPS - 1) / KNOWN_WARPS);
    if (p.filter == MRE_FILTER_INDEX) {
        KnownRuns kr{};
        MRE_TRY(known_runs_scratch(ctx, p, kr));
        rotate_known_score_kernel<<<qgrid, KNOWN_WARPS * 32, 0, st>>>(p, kr);
        rotate_known_compare_kernel<<<qgrid, KNOWN_WARPS * 32, 0, st>>>(p, kr);
        ctx->launches += 2;
    } else {
        rotate_known_direct_kernel<<<qgrid, KNOWN_WARPS * 32, 0, st>>>(p);
        ctx->launches += 1;
    }
    MRE_TRY(ctx->time_begin(st));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(p.total_items, (int64_t)ctx->sm_count * 8));
    rotate_rank_kernel<<<grid, RT_THREADS, 0, st>>>(p);
    MRE_TRY(ctx->time_end(st));
    ctx->launches += 2;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

}  // namespace mre
