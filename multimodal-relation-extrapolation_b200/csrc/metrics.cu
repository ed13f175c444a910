// Rank -> metric sums.
//
// Reference path replaced (paths relative to /root/reference):
//   OpenKE/openke/base/Test.h:102-112,166-177  per-query Hits@10/3/1, rank and reciprocal-rank accumulation
//   OpenKE/openke/base/Test.h:232-277          test_link_prediction: divide by testTotal, average head and tail
//   main.py:263-272, module/zsl_module.py:707-745  the paper's Hits@1/3/10 and Hits@10/5/1 + MRR summaries
// The reference accumulates in float32 globals (Test.h:13-20), which loses integer exactness past 2^24; here every
// sum is an int64 and the reciprocal-rank sum is a float64 reduced in a FIXED order by a single CTA, so the result
// is deterministic.  Slot 6 of every side's sums is the reciprocal-rank sum in 32.32 FIXED POINT (sum floor(2^32 / rank)):
// an integer, so shards of a query set add up to exactly the whole set's value in any order -- the form all-reduced across
// GPUs (MRR error <= 2^-32).  The optional rank histogram is the other all-integer form.
#include "common.h"

namespace mre {

constexpr int MET_THREADS = 1024;

__global__ void __launch_bounds__(MET_THREADS) metrics_kernel(const int32_t *__restrict__ counts, const uint8_t *__restrict__ q_side,
                                                              int side, int64_t Q, int rank_mode, int raw,
                                                              int64_t *__restrict__ sums_out, double *__restrict__ rr_out,
                                                              unsigned long long *__restrict__ hist, int64_t hist_len) {
    __shared__ long long s_int[MET_THREADS / 32][14];
    __shared__ double s_rr[MET_THREADS / 32][2];
    const int32_t *lt = counts + (raw ? 0 : 2) * Q;
    const int32_t *eq = counts + (raw ? 1 : 3) * Q;
    long long acc[2][7] = {{0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0}};
    double rr[2] = {0.0, 0.0};
    // four queries of this thread's stride are fetched together (one exposed load latency per four instead of per one); they
    // are accumulated in the same order as a plain strided loop would, so the float64 sum does not depend on the batching
    constexpr int MB = 4;
    for (int64_t q0 = threadIdx.x; q0 < Q; q0 += (int64_t)MB * MET_THREADS) {
        int32_t va[MB], ve[MB];
        int vs[MB];
#pragma unroll
        for (int k = 0; k < MB; k++) {
            const int64_t q = q0 + (int64_t)k * MET_THREADS;
            const bool ok = q < Q;
            va[k] = ok ? lt[q] : 0;
            ve[k] = ok ? eq[q] : 0;
            vs[k] = ok ? (q_side ? (int)q_side[q] : side) : -1;
        }
#pragma unroll
        for (int k = 0; k < MB; k++) {
            if (vs[k] < 0) continue;
            const int s = vs[k];
            long long a = max(va[k], 0), e = max(ve[k], 0);
            long long rank = a + 1;
            if (rank_mode == MRE_RANK_TIES_HALF) rank += e / 2;
            else if (rank_mode == MRE_RANK_PESSIMISTIC) rank += e;
            const double inv = 1.0 / (double)rank;
            const long long inv_fx = (long long)((1ull << 32) / (unsigned long long)rank);
#pragma unroll
            for (int ss = 0; ss < 2; ss++) {           // static indices: the accumulators stay in registers
                const long long on = s == ss ? 1 : 0;
                acc[ss][0] += on;
                acc[ss][1] += on * rank;
                acc[ss][2] += on & (rank <= 1);
                acc[ss][3] += on & (rank <= 3);
                acc[ss][4] += on & (rank <= 5);
                acc[ss][5] += on & (rank <= 10);
                acc[ss][6] += on * inv_fx;
                rr[ss] += on ? inv : 0.0;
            }
            if (hist) atomicAdd(hist + min((long long)hist_len - 1, rank), 1ull);
        }
    }
    // all 14 quantities together, in a FIXED order: butterfly inside each warp, then the 32 warp partials summed by warp 0
    // in lane order -- deterministic, two block barriers in total
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < 2; s++) {
#pragma unroll
        for (int k = 0; k < 7; k++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[s][k] += __shfl_xor_sync(0xffffffffu, acc[s][k], o);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rr[s] += __shfl_xor_sync(0xffffffffu, rr[s], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < 2; s++) {
#pragma unroll
            for (int k = 0; k < 7; k++) s_int[warp][s * 7 + k] = acc[s][k];
            s_rr[warp][s] = rr[s];
        }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < 14; k++) {
            long long v = s_int[lane][k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) sums_out[(k / 7) * 8 + (k % 7)] = v;
        }
#pragma unroll
        for (int s = 0; s < 2; s++) {
            double v = s_rr[lane][s];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) { rr_out[s] = v; sums_out[s * 8 + 7] = 0; }
        }
    }
}

int metrics(mre_ctx *ctx, const int32_t *counts, const uint8_t *q_side, int32_t side, int64_t Q, int32_t rank_mode,
            int32_t raw, int64_t *sums_out, double *rr_out, int64_t *hist, int64_t hist_len, cudaStream_t st) {
    MRE_CHECK_ARG(counts && sums_out && rr_out, "NULL argument");
    MRE_CHECK_ARG(rank_mode >= MRE_RANK_STRICT && rank_mode <= MRE_RANK_PESSIMISTIC, "unknown rank_mode %d", rank_mode);
    MRE_CHECK_ARG(hist == nullptr || hist_len >= 2, "hist_len must be >= 2");
    MRE_CHECK_ARG(side == 0 || side == 1, "side must be 0 or 1");
    metrics_kernel<<<1, MET_THREADS, 0, st>>>(counts, q_side, side, Q, rank_mode, raw, sums_out, rr_out,
                                              reinterpret_cast<unsigned long long *>(hist), hist_len);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

}  // namespace mre
