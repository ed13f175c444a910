// Index build on the GPU: every sort, the de-duplication and the per-relation counters of importTrainFiles / importTestFiles.
//
// Reference path replaced (paths relative to /root/reference):
//   OpenKE/openke/base/Reader.h:53-160    importTrainFiles: std::sort of trainList by (h,r,t), de-duplication, the (h,r,t) and
//                                         (t,r,h) orders, freqRel / left_mean / right_mean (tph / hpt)
//   OpenKE/openke/base/Reader.h:167-257   importTestFiles: tripleList = test + RAW train + valid sorted by (h,r,t) (the list _find
//                                         bisects), test / valid sorted by (r,h,t)
// mre_index_create (index.cpp) restates them on host threads; here the same lists come out of an LSD radix sort on the device
// (HBM-bound integer work: 12 bytes per triple read and written per 8-bit digit pass), and the filter / sampler tables that
// mre_index_to_device would upload are written in place.  The results are the same bits as the host build (the orders are total
// up to identical triples), which tests/test_index_build_gpu.py asserts column by column.
//
// One pass = three launches: per-tile digit histograms, one exclusive scan over [digit][tile], a stable scatter.  Ids travel as
// int32 structure-of-arrays (h, r, t); a pass sorts by one 8-bit digit of one of the three columns, so the (h,r,t) order is
// digits(t), digits(r), digits(h) least significant first -- ceil(log2 E / 8) * 2 + ceil(log2 R / 8) passes.
// Tile = 8 192 triples per 512-thread CTA (a digit value then owns ~32 consecutive output slots = one 128-byte line per column;
// 2 048-triple tiles wrote isolated 32-byte sectors and ran 2.4x slower once the lists outgrew L2), warp w owning the contiguous
// keys [512 w, 512 w + 512) in sixteen coalesced rounds:
// rank inside a warp round by match.any + popc, across rounds and warps by per-warp running bases in shared memory.
#include <stdio.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "common.h"

namespace mre {

#ifndef MRE_RS_THREADS
#define MRE_RS_THREADS 512
#endif
#ifndef MRE_RS_ROUNDS
#define MRE_RS_ROUNDS 16
#endif
#ifndef MRE_RS_MINB
#define MRE_RS_MINB 3      // resident CTAs per SM the scatter kernel is compiled for (40 registers at 512 threads, no spills): 55 M triples build in 87 / 70 / 64 / 66 ms at 1 / 2 / 3 / 4
#endif
constexpr int RS_THREADS = MRE_RS_THREADS, RS_WARPS = RS_THREADS / 32, RS_ROUNDS = MRE_RS_ROUNDS, RS_TILE = RS_THREADS * RS_ROUNDS;
static_assert(RS_THREADS >= 256, "one thread per digit value");
constexpr int SC_THREADS = 1024, SC_ITEMS = 4, SC_TILE = SC_THREADS * SC_ITEMS;

struct Cols {
    int32_t *h, *r, *t;
};

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const int32_t *__restrict__ key, int64_t n, int shift, uint32_t *__restrict__ tile_hist,
                                                             int n_tiles) {
    __shared__ uint32_t sh[256];
    if (threadIdx.x < 256) sh[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + (threadIdx.x >> 5) * (RS_TILE / RS_WARPS) + (threadIdx.x & 31);
#pragma unroll
    for (int j = 0; j < RS_ROUNDS; j++) {
        const int64_t i = base + j * 32;
        if (i < n) atomicAdd(&sh[((uint32_t)key[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 256) tile_hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = sh[threadIdx.x];   // digit-major: one scan orders digits, then tiles
}

// stable scatter of one tile: `offs` = the exclusive scan of tile_hist
__global__ void __launch_bounds__(RS_THREADS, MRE_RS_MINB) rs_scatter_kernel(const int32_t *__restrict__ key, const Cols in, const Cols out, int64_t n, int shift,
                                                                const uint32_t *__restrict__ offs, int n_tiles) {
    __shared__ uint32_t wbase[RS_WARPS][256];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wbase[0][0])[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + w * (RS_TILE / RS_WARPS) + lane;
    uint32_t dig[RS_ROUNDS];
#pragma unroll
    for (int j = 0; j < RS_ROUNDS; j++) {
        const int64_t i = base + j * 32;
        dig[j] = i < n ? (((uint32_t)key[i] >> shift) & 255u) : 0xffffffffu;
        if (i < n) atomicAdd(&wbase[w][dig[j]], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 256) {   // per digit: the tile's global base, then the warps in order
        const int d = threadIdx.x;
        uint32_t run = offs[(size_t)d * n_tiles + blockIdx.x];
#pragma unroll
        for (int k = 0; k < RS_WARPS; k++) {
            const uint32_t c = wbase[k][d];
            wbase[k][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_ROUNDS; j++) {
        const int64_t i = base + j * 32;
        const bool ok = i < n;
        const unsigned peers = __match_any_sync(0xffffffffu, ok ? dig[j] : 0x100u + (uint32_t)lane);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t pos = 0;
        if (ok) pos = wbase[w][dig[j]] + (uint32_t)rank;
        __syncwarp();
        if (ok && rank == 0) wbase[w][dig[j]] += (uint32_t)__popc(peers);
        __syncwarp();
        if (ok) {
            out.h[pos] = in.h[i];
            out.r[pos] = in.r[i];
            out.t[pos] = in.t[i];
        }
    }
}

// ---- exclusive scan of a uint32 array (in place): tile sums, one CTA over the sums, tile rescan
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *s_warp, uint32_t &total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        s_warp[lane] = s;                                   // inclusive over warps
    }
    __syncthreads();
    total = s_warp[31];
    const uint32_t before = w ? s_warp[w - 1] : 0;
    __syncthreads();
    return before + x - v;
}
__global__ void __launch_bounds__(SC_THREADS) scan_sums_kernel(const uint32_t *__restrict__ a, int64_t m, uint32_t *__restrict__ sums) {
    __shared__ uint32_t s_warp[32];
    const int64_t i0 = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++)
        if (i0 + k < m) v += a[i0 + k];
    uint32_t total;
    block_exclusive_scan(v, s_warp, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SC_THREADS) scan_top_kernel(uint32_t *__restrict__ sums, int64_t nb) {     // one CTA
    __shared__ uint32_t s_warp[32];
    uint32_t carry = 0;
    for (int64_t c0 = 0; c0 < nb; c0 += SC_THREADS) {
        const int64_t i = c0 + threadIdx.x;
        const uint32_t v = i < nb ? sums[i] : 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, s_warp, total);
        if (i < nb) sums[i] = carry + ex;
        carry += total;
    }
}
__global__ void __launch_bounds__(SC_THREADS) scan_apply_kernel(uint32_t *__restrict__ a, int64_t m, const uint32_t *__restrict__ sums) {
    __shared__ uint32_t s_warp[32];
    const int64_t i0 = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    uint32_t x[SC_ITEMS], v = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        x[k] = i0 + k < m ? a[i0 + k] : 0;
        v += x[k];
    }
    uint32_t total;
    uint32_t run = sums[blockIdx.x] + block_exclusive_scan(v, s_warp, total);
#pragma unroll
    for (int k = 0; k < SC_ITEMS; k++) {
        if (i0 + k < m) a[i0 + k] = run;
        run += x[k];
    }
}

// ---- the rest: column conversion, de-duplication, key columns, per-relation counters
__global__ void narrow_kernel(const int64_t *__restrict__ h, const int64_t *__restrict__ t, const int64_t *__restrict__ r, int64_t n, int64_t E,
                              int64_t R, Cols out, int64_t at, int *__restrict__ bad) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = h[i], b = t[i], c = r[i];
        if (a < 0 || a >= E || b < 0 || b >= E || c < 0 || c >= R) atomicMin(bad, (int)min(i, (int64_t)0x7ffffffe));
        out.h[at + i] = (int32_t)a;
        out.t[at + i] = (int32_t)b;
        out.r[at + i] = (int32_t)c;
    }
}
__global__ void flag_kernel(const Cols a, int64_t n, uint32_t *__restrict__ flag) {       // 1 = first of a run of identical triples
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        flag[i] = (i == 0 || a.h[i] != a.h[i - 1] || a.r[i] != a.r[i - 1] || a.t[i] != a.t[i - 1]) ? 1u : 0u;
}
// flag has been scanned (exclusive): element i is kept iff pos[i + 1] != pos[i] (pos[n] = the total, passed in)
__global__ void compact_kernel(const Cols a, int64_t n, const uint32_t *__restrict__ pos, uint32_t total, Cols out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t p = pos[i], q = i + 1 < n ? pos[i + 1] : total;
        if (q != p) {
            out.h[p] = a.h[i];
            out.r[p] = a.r[i];
            out.t[p] = a.t[i];
        }
    }
}
// which = 0: key = h R + r, val = t (and the three columns widened);  1: key = t R + r, val = h
__global__ void keys_kernel(const Cols a, int64_t n, int64_t R, int which, int64_t *__restrict__ key, int64_t *__restrict__ val, int64_t *__restrict__ wh,
                            int64_t *__restrict__ wr, int64_t *__restrict__ wt) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t h = a.h[i], r = a.r[i], t = a.t[i];
        key[i] = (which ? t : h) * R + r;
        if (val) val[i] = which ? h : t;
        if (wh) { wh[i] = h; wr[i] = r; wt[i] = t; }
    }
}
// freqRel, and the number of distinct (fixed entity, r) pairs per relation (Reader.h:142-159), over a list sorted with r inside
// the fixed entity: which = 0 -> the (h,r,t) order counts distinct (h,r); 1 -> the (t,r,h) order counts distinct (t,r)
__global__ void rel_count_kernel(const Cols a, int64_t n, int which, unsigned long long *__restrict__ freq, unsigned long long *__restrict__ distinct) {
    const int lane = threadIdx.x & 31;
    // warp-aggregated: the lanes that hold the same relation elect one to add their total (a few hundred relations take millions of
    // increments: one atomic per (warp, relation) instead of one per triple)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, n_round = (n + stride - 1) / stride * stride;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool ok = i < n;
        const int32_t r = ok ? a.r[i] : -1 - lane;
        const int32_t *f = which ? a.t : a.h;
        const bool first = ok && (i == 0 || f[i] != f[i - 1] || r != a.r[i - 1]);
        const unsigned peers = __match_any_sync(0xffffffffu, r);
        const unsigned firsts = __ballot_sync(0xffffffffu, first) & peers;
        if (ok && (peers & ((1u << lane) - 1u)) == 0) {          // lowest lane of its relation
            if (freq) atomicAdd(freq + r, (unsigned long long)__popc(peers));
            if (firsts) atomicAdd(distinct + r, (unsigned long long)__popc(firsts));
        }
    }
}

static inline int grid1d(int64_t n, int block) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + block - 1) / block, 148 * 16)); }
static inline int bits_for(int64_t count) {              // bits that hold every id in [0, count)
    int b = 1;
    while (b < 31 && (1LL << b) < count) b++;
    return b;
}

struct Sorter {
    Cols a{}, b{};                  // ping-pong columns
    uint32_t *hist = nullptr, *sums = nullptr;
    int64_t cap = 0;
    int bE = 0, bR = 0;
    cudaStream_t st = nullptr;
    int64_t passes = 0;

    int exclusive_scan(uint32_t *x, int64_t m) {
        const int64_t nb = (m + SC_TILE - 1) / SC_TILE;
        scan_sums_kernel<<<(unsigned)nb, SC_THREADS, 0, st>>>(x, m, sums);
        scan_top_kernel<<<1, SC_THREADS, 0, st>>>(sums, nb);
        scan_apply_kernel<<<(unsigned)nb, SC_THREADS, 0, st>>>(x, m, sums);
        return MRE_OK;
    }
    // sorts cur (n triples) by the column order given most significant first ('h', 'r', 't'); the result ends up in `cur`
    int sort(Cols &cur, Cols &other, int64_t n, const char *order) {
        if (n <= 1) return MRE_OK;
        const int n_tiles = (int)((n + RS_TILE - 1) / RS_TILE);
        for (int f = 2; f >= 0; f--) {
            const int bits = order[f] == 'r' ? bR : bE;
            for (int shift = 0; shift < bits; shift += 8) {
                const int32_t *key = order[f] == 'h' ? cur.h : order[f] == 'r' ? cur.r : cur.t;
                rs_hist_kernel<<<n_tiles, RS_THREADS, 0, st>>>(key, n, shift, hist, n_tiles);
                MRE_TRY(exclusive_scan(hist, (int64_t)256 * n_tiles));
                rs_scatter_kernel<<<n_tiles, RS_THREADS, 0, st>>>(key, cur, other, n, shift, hist, n_tiles);
                std::swap(cur, other);
                passes++;
            }
        }
        MRE_CUDA(cudaGetLastError());
        return MRE_OK;
    }
    // drops repeated triples of the sorted `cur`; the result ends up in `cur`, *n_out on the host
    int unique(Cols &cur, Cols &other, int64_t n, int64_t *n_out) {
        *n_out = n;
        if (n <= 1) return MRE_OK;
        flag_kernel<<<grid1d(n, 256), 256, 0, st>>>(cur, n, hist);
        uint32_t last_flag = 0, last_pos = 0;
        MRE_CUDA(cudaMemcpyAsync(&last_flag, hist + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        MRE_TRY(exclusive_scan(hist, n));
        MRE_CUDA(cudaMemcpyAsync(&last_pos, hist + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        MRE_CUDA(cudaStreamSynchronize(st));
        const uint32_t total = last_pos + last_flag;
        compact_kernel<<<grid1d(n, 256), 256, 0, st>>>(cur, n, hist, total, other);
        std::swap(cur, other);
        *n_out = (int64_t)total;
        MRE_CUDA(cudaGetLastError());
        return MRE_OK;
    }
};

static int alloc_cols(Cols &c, int64_t n) {
    const size_t bytes = (size_t)std::max<int64_t>(n, 1) * sizeof(int32_t);
    MRE_CUDA(cudaMalloc((void **)&c.h, bytes));
    MRE_CUDA(cudaMalloc((void **)&c.r, bytes));
    MRE_CUDA(cudaMalloc((void **)&c.t, bytes));
    return MRE_OK;
}
static void free_cols(Cols &c) {
    cudaFree(c.h); cudaFree(c.r); cudaFree(c.t);
    c = Cols{};
}
static int to_host(const Cols &c, int64_t n, std::vector<Triple> &out, std::vector<int32_t> &tmp, cudaStream_t st) {
    out.resize((size_t)n);
    if (n == 0) return MRE_OK;
    tmp.resize((size_t)3 * n);
    MRE_CUDA(cudaMemcpyAsync(tmp.data(), c.h, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    MRE_CUDA(cudaMemcpyAsync(tmp.data() + n, c.r, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    MRE_CUDA(cudaMemcpyAsync(tmp.data() + 2 * n, c.t, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    MRE_CUDA(cudaStreamSynchronize(st));
    const unsigned hc = std::thread::hardware_concurrency();
    const int threads = (int)std::max<int64_t>(1, std::min<int64_t>(hc ? hc : 1, n / (1 << 16)));
    auto widen = [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; i++) out[(size_t)i] = Triple{tmp[(size_t)i], tmp[(size_t)(n + i)], tmp[(size_t)(2 * n + i)]};
    };
    if (threads == 1) {
        widen(0, n);
    } else {
        std::vector<std::thread> pool;
        for (int k = 0; k < threads; k++) pool.emplace_back(widen, n * k / threads, n * (k + 1) / threads);
        for (auto &th : pool) th.join();
    }
    return MRE_OK;
}
template <class T>
static int dmalloc(T **p, int64_t n) {
    MRE_CUDA(cudaMalloc((void **)p, (size_t)std::max<int64_t>(n, 1) * sizeof(T)));
    return MRE_OK;
}

struct Split {
    const int64_t *h, *t, *r;
    int64_t n;
    const char *what;
};

// everything that can fail after the index object exists; the caller destroys `ix` on error
static int build_on_device(mre_index *ix, int device, const Split (&sp)[3] /* train, valid, test */, double *sort_ms) {
    const int64_t E = ix->E, R = ix->R;
    const int64_t n_train = sp[0].n, n_valid = sp[1].n, n_test = sp[2].n, n_all = n_train + n_valid + n_test;
    MRE_CHECK_ARG(E < (1LL << 31) && R < (1LL << 31), "the device build carries ids as int32: E and R must be below 2^31");
    MRE_CHECK_ARG(n_all < (1LL << 31), "the device build addresses triples with 32-bit offsets: at most 2^31 - 1 triples over all splits");
    MRE_TRY(mre_device_ok(device));
    MRE_CUDA(cudaSetDevice(device));
    ix->device = device;                 // from here on mre_index_destroy frees the device columns
    cudaStream_t st = nullptr;           // the legacy default stream: the build is synchronous, like the host build
    Sorter so;
    so.st = st;
    so.bE = bits_for(E);
    so.bR = bits_for(R);
    Cols raw{}, work{}, tmp{};           // raw: [test | train | valid] as given; work / tmp: ping-pong of the sort in flight
    Cols keep_all{}, keep_th{}, keep_tt{}, keep_te{}, keep_va{};   // the sorted lists the host side mirrors, copied back after the device phase
    int64_t *stage = nullptr;
    int *bad = nullptr;
    unsigned long long *cnt = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const int64_t cap = std::max<int64_t>(n_all, 1);
    auto cleanup = [&]() {
        free_cols(raw); free_cols(work); free_cols(tmp);
        free_cols(keep_all); free_cols(keep_th); free_cols(keep_tt); free_cols(keep_te); free_cols(keep_va);
        cudaFree(so.hist); cudaFree(so.sums); cudaFree(stage); cudaFree(bad); cudaFree(cnt);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    };
#define BUILD_TRY(expr)                 \
    do {                                \
        int _rc = (expr);               \
        if (_rc != MRE_OK) {            \
            cleanup();                  \
            return _rc;                 \
        }                               \
    } while (0)
#define BUILD_CUDA(expr) BUILD_TRY([&]() -> int { MRE_CUDA(expr); return MRE_OK; }())
    BUILD_TRY(alloc_cols(raw, cap));
    BUILD_TRY(alloc_cols(work, cap));
    BUILD_TRY(alloc_cols(tmp, cap));
    BUILD_TRY(alloc_cols(keep_all, n_all));
    BUILD_TRY(alloc_cols(keep_th, n_train));
    BUILD_TRY(alloc_cols(keep_tt, n_train));
    BUILD_TRY(alloc_cols(keep_te, n_test));
    BUILD_TRY(alloc_cols(keep_va, n_valid));
    const int64_t n_tiles = (cap + RS_TILE - 1) / RS_TILE;
    const int64_t hist_len = std::max<int64_t>(256 * n_tiles, cap);                  // also the flag / position array of unique()
    BUILD_TRY(dmalloc(&so.hist, hist_len));
    BUILD_TRY(dmalloc(&so.sums, (hist_len + SC_TILE - 1) / SC_TILE + 1));
    const int64_t n_max = std::max(n_train, std::max(n_valid, n_test));
    BUILD_TRY(dmalloc(&stage, 3 * std::max<int64_t>(n_max, 1)));
    BUILD_TRY(dmalloc(&bad, 1));
    BUILD_TRY(dmalloc(&cnt, 3 * R));
    // the tables the index keeps, sized before any duplicate is dropped (allocation stays out of the timed build)
    BUILD_TRY(dmalloc(&ix->d_all_hr_key, n_all));
    BUILD_TRY(dmalloc(&ix->d_all_hr_val, n_all));
    BUILD_TRY(dmalloc(&ix->d_all_tr_key, n_all));
    BUILD_TRY(dmalloc(&ix->d_all_tr_val, n_all));
    BUILD_TRY(dmalloc(&ix->d_tr_h, n_train));
    BUILD_TRY(dmalloc(&ix->d_tr_r, n_train));
    BUILD_TRY(dmalloc(&ix->d_tr_t, n_train));
    BUILD_TRY(dmalloc(&ix->d_tr_hr_key, n_train));
    BUILD_TRY(dmalloc(&ix->d_tr_tr_key, n_train));
    BUILD_TRY(dmalloc(&ix->d_tr_tr_val, n_train));
    BUILD_TRY(dmalloc(&ix->d_bern_prob, R));
    BUILD_CUDA(cudaEventCreate(&e0));
    BUILD_CUDA(cudaEventCreate(&e1));

    // ---- upload + range check + narrow: tripleList's order of insertion is test, train, valid (Reader.h:201-226)
    const int order[3] = {2, 0, 1};
    int64_t at = 0, at_of[3] = {0, 0, 0};
    for (int k = 0; k < 3; k++) {
        const Split &s = sp[order[k]];
        at_of[order[k]] = at;
        if (s.n > 0) {
            const int init = 0x7fffffff;
            BUILD_CUDA(cudaMemcpyAsync(bad, &init, sizeof(int), cudaMemcpyHostToDevice, st));
            BUILD_CUDA(cudaMemcpyAsync(stage, s.h, (size_t)s.n * 8, cudaMemcpyHostToDevice, st));
            BUILD_CUDA(cudaMemcpyAsync(stage + s.n, s.t, (size_t)s.n * 8, cudaMemcpyHostToDevice, st));
            BUILD_CUDA(cudaMemcpyAsync(stage + 2 * s.n, s.r, (size_t)s.n * 8, cudaMemcpyHostToDevice, st));
            narrow_kernel<<<grid1d(s.n, 256), 256, 0, st>>>(stage, stage + s.n, stage + 2 * s.n, s.n, E, R, raw, at, bad);
            int first_bad = 0;
            BUILD_CUDA(cudaMemcpyAsync(&first_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
            BUILD_CUDA(cudaStreamSynchronize(st));
            if (first_bad != 0x7fffffff) {
                const int64_t i = first_bad;
                set_error("%s triple %lld = (h=%lld, t=%lld, r=%lld) out of range for E=%lld, R=%lld", s.what, (long long)i, (long long)s.h[i],
                          (long long)s.t[i], (long long)s.r[i], (long long)E, (long long)R);
                cleanup();
                return MRE_ERR_INVALID;
            }
        }
        at += s.n;
    }
    auto copy_cols = [&](const Cols &src, int64_t off, int64_t n, Cols &dst) -> int {
        if (n == 0) return MRE_OK;
        MRE_CUDA(cudaMemcpyAsync(dst.h, src.h + off, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        MRE_CUDA(cudaMemcpyAsync(dst.r, src.r + off, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        MRE_CUDA(cudaMemcpyAsync(dst.t, src.t + off, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        return MRE_OK;
    };
    std::vector<int32_t> host_tmp;
    BUILD_CUDA(cudaEventRecord(e0, st));

    // ---- tripleList: all splits by (h,r,t), duplicates kept; de-duplicated -> filter tables in both orientations
    BUILD_TRY(copy_cols(raw, 0, n_all, work));
    BUILD_TRY(so.sort(work, tmp, n_all, "hrt"));
    BUILD_TRY(copy_cols(work, 0, n_all, keep_all));
    int64_t n_uniq = 0;
    BUILD_TRY(so.unique(work, tmp, n_all, &n_uniq));
    ix->n_all = n_uniq;
    if (n_uniq) keys_kernel<<<grid1d(n_uniq, 256), 256, 0, st>>>(work, n_uniq, R, 0, ix->d_all_hr_key, ix->d_all_hr_val, nullptr, nullptr, nullptr);
    BUILD_TRY(so.sort(work, tmp, n_uniq, "trh"));
    if (n_uniq) keys_kernel<<<grid1d(n_uniq, 256), 256, 0, st>>>(work, n_uniq, R, 1, ix->d_all_tr_key, ix->d_all_tr_val, nullptr, nullptr, nullptr);

    // ---- trainList: (h,r,t) order, de-duplicated (Reader.h:91-105); trainTail: (t,r,h) (Reader.h:107-109); relation counters
    ix->n_train_raw = n_train;
    BUILD_TRY(copy_cols(raw, at_of[0], n_train, work));
    BUILD_TRY(so.sort(work, tmp, n_train, "hrt"));
    int64_t n_tr = 0;
    BUILD_TRY(so.unique(work, tmp, n_train, &n_tr));
    ix->n_train = n_tr;
    BUILD_TRY(copy_cols(work, 0, n_tr, keep_th));
    BUILD_CUDA(cudaMemsetAsync(cnt, 0, (size_t)3 * R * sizeof(unsigned long long), st));
    if (n_tr) {
        keys_kernel<<<grid1d(n_tr, 256), 256, 0, st>>>(work, n_tr, R, 0, ix->d_tr_hr_key, nullptr, ix->d_tr_h, ix->d_tr_r, ix->d_tr_t);
        rel_count_kernel<<<grid1d(n_tr, 256), 256, 0, st>>>(work, n_tr, 0, cnt, cnt + R);
    }
    BUILD_TRY(so.sort(work, tmp, n_tr, "trh"));
    if (n_tr) {
        keys_kernel<<<grid1d(n_tr, 256), 256, 0, st>>>(work, n_tr, R, 1, ix->d_tr_tr_key, ix->d_tr_tr_val, nullptr, nullptr, nullptr);
        rel_count_kernel<<<grid1d(n_tr, 256), 256, 0, st>>>(work, n_tr, 1, nullptr, cnt + 2 * R);
    }
    BUILD_TRY(copy_cols(work, 0, n_tr, keep_tt));

    // ---- testList / validList by (r,h,t) (Reader.h:227)
    BUILD_TRY(copy_cols(raw, at_of[2], n_test, work));
    BUILD_TRY(so.sort(work, tmp, n_test, "rht"));
    BUILD_TRY(copy_cols(work, 0, n_test, keep_te));
    BUILD_TRY(copy_cols(raw, at_of[1], n_valid, work));
    BUILD_TRY(so.sort(work, tmp, n_valid, "rht"));
    BUILD_TRY(copy_cols(work, 0, n_valid, keep_va));
    BUILD_CUDA(cudaEventRecord(e1, st));

    // ---- the host mirror of the lists (every getter, mre_index_find and the *_host entry points read these)
    BUILD_TRY(to_host(keep_all, n_all, ix->all_head, host_tmp, st));
    BUILD_TRY(to_host(keep_th, n_tr, ix->train_head, host_tmp, st));
    BUILD_TRY(to_host(keep_tt, n_tr, ix->train_tail, host_tmp, st));
    BUILD_TRY(to_host(keep_te, n_test, ix->test, host_tmp, st));
    BUILD_TRY(to_host(keep_va, n_valid, ix->valid, host_tmp, st));

    // ---- tph / hpt / the Bernoulli threshold in the reference's float32 arithmetic (Reader.h:142-159, Base.cpp:113): the
    // reference counts distinct pairs by adding 1.0f to a float, which stops growing at 2^24
    std::vector<unsigned long long> c((size_t)3 * R);
    BUILD_CUDA(cudaMemcpyAsync(c.data(), cnt, c.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    BUILD_CUDA(cudaStreamSynchronize(st));
    ix->left_mean.assign((size_t)R, 0.f);
    ix->right_mean.assign((size_t)R, 0.f);
    ix->bern_prob.assign((size_t)R, 500.f);
    for (int64_t r = 0; r < R; r++) {
        const float lc = (float)std::min<unsigned long long>(c[(size_t)(R + r)], 1ull << 24), rc = (float)std::min<unsigned long long>(c[(size_t)(2 * R + r)], 1ull << 24);
        ix->left_mean[(size_t)r] = (float)(int64_t)c[(size_t)r] / lc;
        ix->right_mean[(size_t)r] = (float)(int64_t)c[(size_t)r] / rc;
        volatile float num = 1000 * ix->right_mean[(size_t)r];
        volatile float den = ix->right_mean[(size_t)r] + ix->left_mean[(size_t)r];
        ix->bern_prob[(size_t)r] = num / den;
    }
    BUILD_CUDA(cudaMemcpy(ix->d_bern_prob, ix->bern_prob.data(), (size_t)R * sizeof(float), cudaMemcpyHostToDevice));
    float ms = 0.f;
    BUILD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (sort_ms) *sort_ms = ms;
    BUILD_CUDA(cudaGetLastError());
    cleanup();
#undef BUILD_TRY
#undef BUILD_CUDA
    return MRE_OK;
}

}  // namespace mre

using namespace mre;

extern "C" int mre_index_create_device(int device, int64_t E, int64_t R, const int64_t *train_h, const int64_t *train_t, const int64_t *train_r,
                                       int64_t n_train, const int64_t *valid_h, const int64_t *valid_t, const int64_t *valid_r, int64_t n_valid,
                                       const int64_t *test_h, const int64_t *test_t, const int64_t *test_r, int64_t n_test, mre_index **out,
                                       double *build_ms) {
    MRE_CHECK_ARG(out != nullptr, "out is NULL");
    MRE_CHECK_ARG(E > 0 && R > 0, "E and R must be positive");
    MRE_CHECK_ARG(n_train >= 0 && n_valid >= 0 && n_test >= 0, "negative split size");
    MRE_CHECK_ARG(E <= (INT64_MAX / 2) / R, "E * R overflows the packed key");
    mre_index *ix = new mre_index();
    ix->E = E;
    ix->R = R;
    const Split sp[3] = {{train_h, train_t, train_r, n_train, "train"}, {valid_h, valid_t, valid_r, n_valid, "valid"}, {test_h, test_t, test_r, n_test, "test"}};
    const int rc = build_on_device(ix, device, sp, build_ms);
    if (rc != MRE_OK) {
        mre_index_destroy(ix);
        return rc;
    }
    *out = ix;
    return MRE_OK;
}

// mre_index_create_from_dir with the build on the GPU: the files are parsed on the host, everything after that as above
extern "C" int mre_index_create_from_dir_device(const char *in_path, int device, mre_index **out, double *build_ms) {
    MRE_CHECK_ARG(in_path && out, "NULL argument");
    std::string dir(in_path);
    if (!dir.empty() && dir.back() != '/') dir += '/';
    int64_t E = 0, R = 0;
    std::vector<Triple> sp[3];
    MRE_TRY(read_benchmark_dir(dir, &E, &R, sp[0], sp[1], sp[2]));
    std::vector<int64_t> col[3][3];
    for (int k = 0; k < 3; k++) {
        for (auto &c : col[k]) c.resize(sp[k].size());
        for (size_t i = 0; i < sp[k].size(); i++) { col[k][0][i] = sp[k][i].h; col[k][1][i] = sp[k][i].t; col[k][2][i] = sp[k][i].r; }
        std::vector<Triple>().swap(sp[k]);
    }
    mre_index *ix = nullptr;
    MRE_TRY(mre_index_create_device(device, E, R, col[0][0].data(), col[0][1].data(), col[0][2].data(), (int64_t)col[0][0].size(), col[1][0].data(),
                                    col[1][1].data(), col[1][2].data(), (int64_t)col[1][0].size(), col[2][0].data(), col[2][1].data(), col[2][2].data(),
                                    (int64_t)col[2][0].size(), &ix, build_ms));
    FILE *f = fopen((dir + "type_constrain.txt").c_str(), "r");      // importTypeFiles (Reader.h:267-317); optional here
    if (f) {
        fclose(f);
        const int rc = mre_index_load_type_constrain(ix, (dir + "type_constrain.txt").c_str());
        if (rc != MRE_OK) {
            mre_index_destroy(ix);
            return rc;
        }
    }
    *out = ix;
    return MRE_OK;
}
