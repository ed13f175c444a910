// Known-true filter, re-cut per work item: the (query row, candidate column) pairs the rank epilogues must leave out of
// the FILTERED counts, grouped by the 128 x N tile they fall in.
//
// Reference path replaced (paths relative to /root/reference):
//   OpenKE/openke/base/Test.h:80-87,144-151   "if (!_find(j, t, r)) l_filter_s += 1" inside the E-long compare loop
//   OpenKE/openke/base/Corrupt.h:166-177      _find: one binary search in tripleList per counted candidate
//   utils/gen_mode_candidates.py:30-34        the paper's candidate lists with known tails removed
// The reference asks "is candidate j known?" once per counted candidate (E/2 binary searches per query on a random
// model).  Here the question is inverted: every known entity of every query is routed ONCE to the tile that will score
// it (count -> exclusive scan -> fill: a counting sort by work item), and the tile's epilogue marks those accumulator
// slots in a bit mask.  The filtered count is then taken from the very registers the raw count is taken from -- no second
// scoring pass, no inconsistency between two arithmetic paths.  The true entity of each query is routed the same way, so
// it never counts against itself.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.h"
#include "rank_common.cuh"
#include "rank_host.h"

namespace mre {

// visit every (known entity | truth) of query q that lies in q's candidate set: fn(item, row_local, col_local)
template <class Fn>
__device__ __forceinline__ void for_each_known(const RankParams &p, int tile_q, int tile_e, int64_t q, int lane, Fn fn) {
    const int side = query_side(p, q);
    const int64_t h = p.q_h[q], t = p.q_t[q], r = p.q_r[q];
    const int64_t truth = side ? t : h;
    const int64_t *list = nullptr;
    int64_t lo = 0, hi = 0;
    if (p.filter == MRE_FILTER_INDEX) {
        const int64_t *keys = side ? p.hr_key : p.tr_key;
        const int64_t key = (side ? h : t) * p.R + r;
        lo = lower_bound_i64(keys, 0, p.n_all, key);
        hi = lower_bound_i64(keys, lo, p.n_all, key + 1);
        list = side ? p.hr_val : p.tr_val;
    } else if (p.filter == MRE_FILTER_CSR) {
        lo = p.filt_ptr[q];
        hi = p.filt_ptr[q + 1];
        list = p.filt_idx;
    }
    const GroupDesc &gd = p.groups[p.all_entities ? 0 : group_of_query(p, q)];
    if (q < gd.q0 || q - gd.q0 >= gd.nq) return;                       // the query's own group was empty (dropped): nothing is scored for it
    // tile sizes are powers of two (checked on the host): shifts, not 64-bit divisions
    const int sh_q = 31 - __clz(tile_q), sh_e = 31 - __clz(tile_e);
    const int64_t qt = (q - gd.q0) >> sh_q;
    const int row = (int)((q - gd.q0) & (tile_q - 1));
    for (int64_t i = lo + lane; i <= hi; i += 32) {        // index hi stands for the true entity itself
        const int64_t x = i < hi ? __ldg(list + i) : truth;
        if (i < hi && x == truth) continue;                 // listed once, as the last entry
        if (x < 0 || x >= p.E) continue;
        int64_t pos = x;                                    // position inside the group's candidate list
        if (!p.all_entities) {
            const int64_t k = lower_bound_i64(p.cand_idx, gd.c0, gd.c0 + gd.nc, x);
            if (k >= gd.c0 + gd.nc || __ldg(p.cand_idx + k) != x) continue;
            pos = k - gd.c0;
        }
        const int64_t et = pos >> sh_e;
        fn(gd.item0 + et * gd.n_qt + qt, row, (int)(pos & (tile_e - 1)));
    }
}

__global__ void __launch_bounds__(256) tf_count_kernel(const RankParams p, int tile_q, int tile_e, uint32_t *__restrict__ cnt) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < p.Q; q += warps)
        for_each_known(p, tile_q, tile_e, q, lane, [&](int64_t item, int, int) { atomicAdd(cnt + item, 1u); });
}

__global__ void __launch_bounds__(256) tf_fill_kernel(const RankParams p, int tile_q, int tile_e, uint32_t *__restrict__ cnt,
                                                      const uint32_t *__restrict__ ptr, uint32_t *__restrict__ pairs) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < p.Q; q += warps)
        for_each_known(p, tile_q, tile_e, q, lane, [&](int64_t item, int row, int col) {
            const uint32_t slot = atomicSub(cnt + item, 1u) - 1u;    // counts run back down to zero
            pairs[ptr[item] + slot] = ((uint32_t)row << 16) | (uint32_t)col;
        });
}

// exclusive scan of n counters in three steps: per-block scan of SCAN_CHUNK elements, scan of the block totals, add-back
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_PER_THREAD = 16;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_PER_THREAD;

__global__ void __launch_bounds__(SCAN_THREADS) scan_blocks_kernel(const uint32_t *__restrict__ in, int64_t n, uint32_t *__restrict__ out,
                                                                   uint32_t *__restrict__ block_sums) {
    __shared__ uint32_t sh[SCAN_THREADS];
    const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_PER_THREAD;
    uint32_t v[SCAN_PER_THREAD], sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        v[k] = base + k < n ? in[base + k] : 0u;
        sum += v[k];
    }
    sh[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < SCAN_THREADS; off <<= 1) {      // Hillis-Steele inclusive scan of the per-thread sums
        uint32_t add = threadIdx.x >= off ? sh[threadIdx.x - off] : 0u;
        __syncthreads();
        sh[threadIdx.x] += add;
        __syncthreads();
    }
    uint32_t run = sh[threadIdx.x] - sum;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == SCAN_THREADS - 1) block_sums[blockIdx.x] = sh[threadIdx.x];
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(uint32_t *__restrict__ block_sums, int64_t nb, uint32_t *__restrict__ total) {
    __shared__ uint32_t sh[1024];
    uint32_t carry = 0;
    for (int64_t base = 0; base < nb; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = i < nb ? block_sums[i] : 0u;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            uint32_t add = threadIdx.x >= off ? sh[threadIdx.x - off] : 0u;
            __syncthreads();
            sh[threadIdx.x] += add;
            __syncthreads();
        }
        if (i < nb) block_sums[i] = carry + sh[threadIdx.x] - v;
        const uint32_t chunk_total = sh[1023];
        __syncthreads();
        carry += chunk_total;
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(uint32_t *__restrict__ out, int64_t n, const uint32_t *__restrict__ block_sums,
                                                                const uint32_t *__restrict__ total) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK;
    const uint32_t add = block_sums[blockIdx.x];
    for (int k = threadIdx.x; k < SCAN_CHUNK; k += SCAN_THREADS)
        if (base + k < n) out[base + k] += add;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *total;     // ptr[n] = number of pairs
}

// the same exclusive scan in ONE launch when the counters fit one block (n <= 1024 * 16): small jobs are launch-latency bound
constexpr int SMALL_SCAN_MAX = 1024 * 16;
__global__ void __launch_bounds__(1024) scan_small_kernel(const uint32_t *__restrict__ in, int n, uint32_t *__restrict__ out,
                                                          uint32_t *__restrict__ total) {
    __shared__ uint32_t sh[1024];
    const int per = (n + 1023) / 1024, base = threadIdx.x * per;
    uint32_t v[16], sum = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        v[k] = (k < per && base + k < n) ? in[base + k] : 0u;
        sum += v[k];
    }
    sh[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const uint32_t add = threadIdx.x >= off ? sh[threadIdx.x - off] : 0u;
        __syncthreads();
        sh[threadIdx.x] += add;
        __syncthreads();
    }
    uint32_t run = sh[threadIdx.x] - sum;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k < per && base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 1023) {
        out[n] = sh[1023];
        *total = sh[1023];
    }
}

// count -> scan -> fill in ONE cooperative launch (grid-wide barriers between the phases) for jobs whose counters fit one
// block's scan (n <= 16 384 work items: every real dataset of the reference) and whose pair-list capacity is known up front.
// Three dependent launches of a few microseconds of work each cost more in launch gaps and cold-cache round trips than in
// work; here the phases share one launch and the query descriptors stay in L1/L2 between them.
constexpr int TF_FUSED_THREADS = 512;
__global__ void __launch_bounds__(TF_FUSED_THREADS) tf_fused_kernel(const RankParams p, int tile_q, int tile_e, uint32_t *__restrict__ cnt,
                                                                    uint32_t *__restrict__ ptr, uint32_t *__restrict__ total,
                                                                    uint32_t *__restrict__ pairs, int n) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5), w0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (int64_t q = w0; q < p.Q; q += warps)
        for_each_known(p, tile_q, tile_e, q, lane, [&](int64_t item, int, int) { atomicAdd(cnt + item, 1u); });
    grid.sync();
    if (blockIdx.x == 0) {                                  // exclusive scan of the n counters by one block
        __shared__ uint32_t sh[TF_FUSED_THREADS];
        const int per = (n + TF_FUSED_THREADS - 1) / TF_FUSED_THREADS, base = threadIdx.x * per;    // per <= 32
        uint32_t v[32], sum = 0;                            // all of a thread's loads are issued together: one latency, not 32
#pragma unroll
        for (int k = 0; k < 32; k++) {
            v[k] = (k < per && base + k < n) ? cnt[base + k] : 0u;
            sum += v[k];
        }
        sh[threadIdx.x] = sum;
        __syncthreads();
        for (int off = 1; off < TF_FUSED_THREADS; off <<= 1) {
            const uint32_t add = threadIdx.x >= off ? sh[threadIdx.x - off] : 0u;
            __syncthreads();
            sh[threadIdx.x] += add;
            __syncthreads();
        }
        uint32_t run = sh[threadIdx.x] - sum;
#pragma unroll
        for (int k = 0; k < 32; k++) {
            if (k < per && base + k < n) ptr[base + k] = run;
            run += v[k];
        }
        if (threadIdx.x == TF_FUSED_THREADS - 1) { ptr[n] = sh[TF_FUSED_THREADS - 1]; *total = sh[TF_FUSED_THREADS - 1]; }
    }
    grid.sync();
    for (int64_t q = w0; q < p.Q; q += warps)
        for_each_known(p, tile_q, tile_e, q, lane, [&](int64_t item, int row, int col) {
            const uint32_t slot = atomicSub(cnt + item, 1u) - 1u;    // counts run back down to zero
            pairs[ptr[item] + slot] = ((uint32_t)row << 16) | (uint32_t)col;
        });
}

int build_tile_filter(mre_ctx *ctx, const mre_rank_job *job, RankParams &p, int tile_q, int tile_e, cudaStream_t st) {
    p.tf_ptr = nullptr;
    p.tf_pairs = nullptr;
    if (p.Q == 0 || p.total_items == 0) return MRE_OK;
    const int64_t n = p.total_items;
    MRE_CHECK_ARG(n < (1LL << 31), "too many work items (%lld); rank the queries in smaller batches", (long long)n);
    const int64_t nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    // layout of ctx->counters: cnt[n] | ptr[n + 1] | block_sums[nb] | total[1]
    const void *before = ctx->counters.p;
    MRE_TRY(ctx->counters.reserve((size_t)(2 * n + nb + 8) * sizeof(uint32_t)));
    uint32_t *cnt = ctx->counters.as<uint32_t>(), *ptr = cnt + n, *bsum = ptr + n + 1, *total = bsum + nb;
    // tf_fill_kernel runs every counter it filled back down to zero, so the first `counters_armed` words are already zero when
    // the previous job on this context had at least as many work items: no memset node in the steady state
    if (ctx->counters.p != before || ctx->counters_armed < n) MRE_CUDA(cudaMemsetAsync(cnt, 0, (size_t)n * sizeof(uint32_t), st));
    ctx->counters_armed = 0;
    int64_t known_cap = -1;
    if (job->filter == MRE_FILTER_NONE) known_cap = p.Q;
    else if (job->filt_nnz > 0) known_cap = job->filt_nnz + p.Q;
    // The cooperative single-launch form is kept behind mre_ctx_option "tf_fused": a cooperative launch does not run beside the
    // main stream's table / query kernels (the two pre-pass chains serialise: measured 64 + 56 us instead of max(55, 56) us on
    // FB15K-237-ZS DistMult), so the three plain launches on the second stream are the default.
    if (ctx->opt_tf_fused && n <= SMALL_SCAN_MAX && known_cap >= 0 && known_cap < (1LL << 31)) {
        MRE_TRY(ctx->misc2.reserve((size_t)std::max<int64_t>(known_cap, 1) * sizeof(uint32_t)));
        if (ctx->tf_fused_blocks_per_sm == 0) {
            int per_sm = 0;
            MRE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tf_fused_kernel, TF_FUSED_THREADS, 0));
            ctx->tf_fused_blocks_per_sm = std::max(per_sm, 1);
        }
        // cooperative launch: every block must be resident; one block per SM is plenty for the real datasets
        const int cgrid = (int)std::max<int64_t>(1, std::min<int64_t>((p.Q + 15) / 16, (int64_t)ctx->sm_count * std::min(ctx->tf_fused_blocks_per_sm, 2)));
        uint32_t *pairs = ctx->misc2.as<uint32_t>();
        int n32 = (int)n;
        void *args[] = {(void *)&p, (void *)&tile_q, (void *)&tile_e, (void *)&cnt, (void *)&ptr, (void *)&total, (void *)&pairs, (void *)&n32};
        MRE_CUDA(cudaLaunchCooperativeKernel((const void *)tf_fused_kernel, dim3((unsigned)cgrid), dim3(TF_FUSED_THREADS), args, 0, st));
        ctx->launches += 1;
        ctx->counters_armed = n;
        p.tf_ptr = ptr;
        p.tf_pairs = pairs;
        return MRE_OK;
    }
    MRE_CHECK_ARG((tile_q & (tile_q - 1)) == 0 && (tile_e & (tile_e - 1)) == 0, "tile sizes must be powers of two");
    // one warp per query while that stays within a few waves: the per-query chain of dependent loads runs once per warp
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((p.Q + 7) / 8, (int64_t)ctx->sm_count * 32));
    tf_count_kernel<<<grid, 256, 0, st>>>(p, tile_q, tile_e, cnt);
    if (n <= SMALL_SCAN_MAX) {
        scan_small_kernel<<<1, 1024, 0, st>>>(cnt, (int)n, ptr, total);
        ctx->launches += 2;
    } else {
        scan_blocks_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(cnt, n, ptr, bsum);
        scan_sums_kernel<<<1, 1024, 0, st>>>(bsum, nb, total);
        scan_add_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(ptr, n, bsum, total);
        ctx->launches += 4;
    }
    // capacity of the pair list: the caller's bound when it gave one, else read the exact total back (one small sync)
    int64_t cap = -1;
    if (job->filter == MRE_FILTER_NONE) cap = p.Q;
    else if (job->filt_nnz > 0) cap = job->filt_nnz + p.Q;
    if (cap < 0) {
        uint32_t host_total = 0;
        MRE_CUDA(cudaMemcpyAsync(&host_total, total, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        MRE_CUDA(cudaStreamSynchronize(st));
        cap = host_total;
    }
    MRE_CHECK_ARG(cap < (1LL << 31), "too many known-true pairs (%lld); rank the queries in smaller batches", (long long)cap);
    MRE_TRY(ctx->misc2.reserve((size_t)std::max<int64_t>(cap, 1) * sizeof(uint32_t)));
    tf_fill_kernel<<<grid, 256, 0, st>>>(p, tile_q, tile_e, cnt, ptr, ctx->misc2.as<uint32_t>());
    ctx->launches += 1;
    ctx->counters_armed = n;
    MRE_CUDA(cudaGetLastError());
    p.tf_ptr = ptr;
    p.tf_pairs = ctx->misc2.as<uint32_t>();
    return MRE_OK;
}

}  // namespace mre
