// Rank -> metric sums.
//
// Reference path replaced (paths relative to /root/reference):
//   OpenKE/openke/base/Test.h:102-112,166-177  per-query Hits@10/3/1, rank and reciprocal-rank accumulation
//   OpenKE/openke/base/Test.h:232-277          test_link_prediction: divide by testTotal, average head and tail
//   main.py:263-272, module/zsl_module.py:707-745  the paper's Hits@1/3/10 and Hits@10/5/1 + MRR summaries
// The reference accumulates in float32 globals (Test.h:13-20), which loses integer exactness past 2^24; here every
// sum is an int64 and the reciprocal-rank sum is a float64 reduced in a FIXED order (per CTA, then over the CTAs), so the
// result is deterministic.  Slot 6 of every side's sums is the reciprocal-rank sum in 32.32 FIXED POINT (sum floor(2^32 / rank)):
// an integer, so shards of a query set add up to exactly the whole set's value in any order -- the form all-reduced across
// GPUs (MRR error <= 2^-32).  The optional rank histogram is the other all-integer form.
#include <algorithm>

#include "common.h"

namespace mre {

constexpr int MET_THREADS = 256;
constexpr int MET_MAX_BLOCKS = 256;

// Grid of CTAs, each over a contiguous slice of the queries.  Integer sums are combined with int64 atomics (associative: the
// result does not depend on the order); the float64 reciprocal-rank sum is combined in a FIXED order: every CTA reduces its
// slice in a fixed order into a partial, and the last CTA to finish (a ticket counter) adds the partials in CTA order -- so
// rr_out is deterministic for a given (Q, grid), and the grid is a function of Q only.  (One 1024-thread CTA did all of it
// before: ~200 integer/FP64 instructions per query on ONE SM cost 20+ us at 17 596 queries.)
__global__ void __launch_bounds__(MET_THREADS) metrics_kernel(const int32_t *__restrict__ counts, const uint8_t *__restrict__ q_side,
                                                              int side, int64_t Q, int rank_mode, int raw,
                                                              unsigned long long *__restrict__ sums_out, double *__restrict__ rr_out,
                                                              unsigned long long *__restrict__ hist, int64_t hist_len,
                                                              unsigned long long *__restrict__ partial, unsigned int *__restrict__ ticket) {
    __shared__ long long s_int[MET_THREADS / 32][14];
    __shared__ double s_rr[MET_THREADS / 32][2];
    __shared__ bool is_last;
    const int32_t *lt = counts + (raw ? 0 : 2) * Q;
    const int32_t *eq = counts + (raw ? 1 : 3) * Q;
    long long acc[2][7] = {{0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0}};
    double rr[2] = {0.0, 0.0};
    const int64_t per = (Q + gridDim.x - 1) / gridDim.x;
    const int64_t q_lo = (int64_t)blockIdx.x * per, q_hi = min(Q, q_lo + per);
    for (int64_t q = q_lo + threadIdx.x; q < q_hi; q += MET_THREADS) {
        const int s = q_side ? (int)q_side[q] : side;
        const long long a = max(lt[q], 0), e = max(eq[q], 0);
        long long rank = a + 1;
        if (rank_mode == MRE_RANK_TIES_HALF) rank += e / 2;
        else if (rank_mode == MRE_RANK_PESSIMISTIC) rank += e;
        const double inv = 1.0 / (double)rank;
        // ranks fit 32 bits (E < 2^31): one 64 / 32 division instead of the generic 64 / 64 routine
        const long long inv_fx = (long long)((1ull << 32) / (unsigned long long)(unsigned int)rank);
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
    // butterfly inside each warp, then the warp partials summed by warp 0 in lane order: a fixed order
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
    // 14 integer sums + 2 float64 sums of this CTA -> its row of the partials; the LAST CTA to finish (ticket counter) adds the
    // rows in CTA order and writes the outputs: no atomics on the results, no zeroing launch, a fixed order for the doubles
    if (threadIdx.x < 16) {
        unsigned long long bits;
        if (threadIdx.x < 14) {
            long long v = 0;
            for (int w = 0; w < MET_THREADS / 32; w++) v += s_int[w][threadIdx.x];
            bits = (unsigned long long)v;
        } else {
            double v = 0.0;
            for (int w = 0; w < MET_THREADS / 32; w++) v += s_rr[w][threadIdx.x - 14];
            bits = (unsigned long long)__double_as_longlong(v);
        }
        partial[16 * blockIdx.x + threadIdx.x] = bits;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (is_last && threadIdx.x < 16) {
        __threadfence();
        const volatile unsigned long long *pv = partial;
        if (threadIdx.x < 14) {
            long long v = 0;
            for (unsigned b = 0; b < gridDim.x; b++) v += (long long)pv[16 * b + threadIdx.x];
            sums_out[(threadIdx.x / 7) * 8 + (threadIdx.x % 7)] = (unsigned long long)v;
        } else {
            double v = 0.0;
            for (unsigned b = 0; b < gridDim.x; b++) v += __longlong_as_double((long long)pv[16 * b + threadIdx.x]);
            rr_out[threadIdx.x - 14] = v;
            sums_out[(threadIdx.x - 14) * 8 + 7] = 0ull;
        }
        if (threadIdx.x == 0) *ticket = 0u;          // re-armed for the next call
    }
}

int metrics(mre_ctx *ctx, const int32_t *counts, const uint8_t *q_side, int32_t side, int64_t Q, int32_t rank_mode,
            int32_t raw, int64_t *sums_out, double *rr_out, int64_t *hist, int64_t hist_len, cudaStream_t st) {
    MRE_CHECK_ARG(counts && sums_out && rr_out, "NULL argument");
    MRE_CHECK_ARG(rank_mode >= MRE_RANK_STRICT && rank_mode <= MRE_RANK_PESSIMISTIC, "unknown rank_mode %d", rank_mode);
    MRE_CHECK_ARG(hist == nullptr || hist_len >= 2, "hist_len must be >= 2");
    MRE_CHECK_ARG(side == 0 || side == 1, "side must be 0 or 1");
    if (!ctx->met_scratch.p) {       // per-CTA rows of 16 partials + the ticket counter (re-armed by the kernel)
        MRE_TRY(ctx->met_scratch.reserve((size_t)MET_MAX_BLOCKS * 16 * sizeof(unsigned long long) + 64));
        MRE_CUDA(cudaMemset(ctx->met_scratch.p, 0, (size_t)MET_MAX_BLOCKS * 16 * sizeof(unsigned long long) + 64));
    }
    unsigned long long *partial = ctx->met_scratch.as<unsigned long long>();
    unsigned int *ticket = reinterpret_cast<unsigned int *>(partial + MET_MAX_BLOCKS * 16);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((Q + 4 * MET_THREADS - 1) / (4 * MET_THREADS), MET_MAX_BLOCKS));
    metrics_kernel<<<grid, MET_THREADS, 0, st>>>(counts, q_side, side, Q, rank_mode, raw, reinterpret_cast<unsigned long long *>(sums_out),
                                                 rr_out, reinterpret_cast<unsigned long long *>(hist), hist_len, partial, ticket);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

}  // namespace mre
