// Data-parallel exchanges over NVLink peer memory: the gradient reduction FUSED into the SGD update, and the integer metric
// all-reduce of the sharded evaluation, for one process per GPU on one NVSwitch box.
//
// Reference path replaced (paths relative to /root/reference):
//   OpenKE/openke/config/Trainer.py:43-54,73-78   loss.backward(); optimizer.step()  (torch.optim.SGD, one process, one GPU)
//   OpenKE/openke/base/Test.h:232-277             the metric sums of test_link_prediction (one process)
// The reference is single-GPU; BASELINE configs[3] asks for the step data-parallel on 1/2/4/8 B200.  A data-parallel step is
// "sum the 11.8 MB gradient buffer over the ranks, then w -= lr * g": with NCCL that is one all-reduce whose ~200 us of ring /
// tree latency is two thirds of the 0.3 ms step.  Here the two are ONE kernel per rank over peer memory: rank r owns the slice
// [r n / N, (r + 1) n / N) of the flat parameter buffer; it loads that slice of every rank's gradient buffer straight over NVLink
// (P2P loads, summed in rank order -- every slice has exactly one summation order, so all ranks hold bit-identical weights),
// applies the update, and stores the new weights into every rank's weight buffer (P2P stores): reduce-scatter + SGD + all-gather
// with the data crossing the switch once in each direction (2 x 11.8 MB x (N - 1) / N per GPU: ~30 us of NVLink time).
// Ordering is two flag rounds in peer memory -- "my gradients are complete" before the loads, "I am done with your buffers"
// before anyone's next kernel may touch them -- written with st.release.sys after __threadfence_system and polled with
// ld.acquire.sys on the local copy; step numbers only grow, so flags are never reset.  Every spin is bounded (a rank that never
// arrives sets the error word instead of hanging the GPU).
//
// Peer memory is plain cudaMalloc + CUDA IPC: mre_peer_group_create returns a 64-byte handle the caller ships to the other
// processes by any means (torch.distributed.all_gather_object in dist.py; a C caller uses its own channel),
// mre_peer_group_connect maps the other ranks' regions.
#include <string.h>

#include <algorithm>

#include "common.h"

namespace mre {

constexpr int PEER_MAX = 16;            // ranks of one exchange (one NVSwitch box holds 8)
constexpr int PEER_THREADS = 512;
constexpr long long PEER_SPIN_CLOCKS = 40000000000LL;    // ~20 s at 1.9 GHz: a peer that has not arrived by then never will

struct PeerPtrs {
    float *w[PEER_MAX];
    float *g[PEER_MAX];
    unsigned int *flags[PEER_MAX];      // per rank: [0, N) arrivals, [N, 2N) departures, [2N] finished-block counter, [2N + 1] error
};

__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float *p) {      // never from a stale L1 line: the buffer is rewritten every step
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer_f1(const float *p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// all ranks' flag `slot0 + p` on MY flag array reach `step`: lanes poll one peer each; false (and the error word set) on timeout
__device__ __forceinline__ bool wait_all(const PeerPtrs &pp, int rank, int world, int slot0, unsigned int step) {
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const unsigned int *f = pp.flags[rank] + slot0 + threadIdx.x;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(f) - step) < 0) {
            if (clock64() - t0 > PEER_SPIN_CLOCKS) {
                ok = 0;
                atomicExch(pp.flags[rank] + 2 * world + 1, 1u);
                break;
            }
            __nanosleep(100);
        }
    }
    __syncthreads();
    return ok != 0;
}

// ONE kernel per rank and step: arrive -> wait -> (reduce my slice over all ranks, update, broadcast) -> depart -> wait -> zero my gradients
__global__ void __launch_bounds__(PEER_THREADS) dp_sgd_kernel(const PeerPtrs pp, int rank, int world, int64_t n, float lr, unsigned int step) {
    unsigned int *mine = pp.flags[rank];
    // ---- my gradients are complete (stream order: the backward kernel ran before this launch): tell every rank
    if (blockIdx.x == 0 && (int)threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(pp.flags[threadIdx.x] + rank, step);
    }
    if (!wait_all(pp, rank, world, 0, step)) return;
    // ---- my slice, in float4 units (n % 4 == 0 is checked on the host)
    const int64_t n4 = n >> 2, base = n4 / world, rem = n4 % world;
    const int64_t lo = rank * base + min((int64_t)rank, rem), hi = lo + base + (rank < rem ? 1 : 0);
    const float4 *wsrc = reinterpret_cast<const float4 *>(pp.w[rank]);
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
        float4 gs[PEER_MAX];
#pragma unroll
        for (int p = 0; p < PEER_MAX; p++)                 // every rank's loads are in flight together
            if (p < world) gs[p] = ld_peer_f4(pp.g[p] + 4 * i);
        float4 s = gs[0];
#pragma unroll
        for (int p = 1; p < PEER_MAX; p++)                 // summed in rank order: one order per element, the same on every GPU
            if (p < world) { s.x += gs[p].x; s.y += gs[p].y; s.z += gs[p].z; s.w += gs[p].w; }
        float4 w = wsrc[i];
        w.x = w.x - lr * s.x; w.y = w.y - lr * s.y; w.z = w.z - lr * s.z; w.w = w.w - lr * s.w;
#pragma unroll
        for (int p = 0; p < PEER_MAX; p++)
            if (p < world) reinterpret_cast<float4 *>(pp.w[p])[i] = w;
    }
    // ---- the last block of this rank to finish announces "I am done with your gradient slices and your weights are written"
    __threadfence_system();
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0) last = atomicAdd(mine + 2 * world, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) mine[2 * world] = 0u;        // re-armed for the next step (nobody reads it before the next launch)
        if ((int)threadIdx.x < world) {
            __threadfence_system();
            st_release_sys(pp.flags[threadIdx.x] + world + rank, step);
        }
    }
    // ---- every rank is done with MY buffers: my gradient buffer may be zeroed (the next backward accumulates into it) and the
    // kernel may end (the next forward reads my weights)
    if (!wait_all(pp, rank, world, world, step)) return;
    float4 *g4 = reinterpret_cast<float4 *>(pp.g[rank]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
        g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// The integer metric sums of the sharded evaluation (mre_metrics' int64 [2, 8]; any small int64 vector), all-reduced over the
// same flags: every rank publishes its vector into its own slot of every rank's exchange area, then sums the N slots it holds.
// count <= 64.  One block.
__global__ void __launch_bounds__(64) peer_allreduce_i64_kernel(const PeerPtrs pp, long long *const *xchg, int rank, int world,
                                                                long long *vec, int count, unsigned int step) {
    const int i = threadIdx.x;
    if (i < count) {
        const long long v = vec[i];
        for (int p = 0; p < world; p++) xchg[p][rank * 64 + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (i < world) st_release_sys(pp.flags[i] + rank, step);
    if (!wait_all(pp, rank, world, 0, step)) return;
    if (i < count) {
        long long s = 0;
        for (int p = 0; p < world; p++) s += *reinterpret_cast<volatile long long *>(xchg[rank] + p * 64 + i);
        vec[i] = s;
    }
    // departures: my slots on the other ranks may only be rewritten (next call) once they have all summed
    __threadfence_system();
    __syncthreads();
    if (i < world) st_release_sys(pp.flags[i] + world + rank, step);
    wait_all(pp, rank, world, world, step);
}

}  // namespace mre

using namespace mre;

struct mre_peer_group {
    int rank = 0, world = 1, device = 0;
    int64_t n = 0;                      // floats of the weight / gradient buffers
    unsigned int step = 0;              // exchanges so far (flags hold the number of the last one)
    void *base[PEER_MAX] = {nullptr};   // every rank's region: [w n | g n | xchg int64 world * 64 | flags]
    size_t off_g = 0, off_x = 0, off_f = 0, bytes = 0;
    long long **d_xchg = nullptr;       // device copy of the exchange-area pointers
    bool local = false;                 // connect_local: the peers' regions are plain pointers of this process
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static void peer_layout(mre_peer_group *g) {
    g->off_g = align256((size_t)g->n * sizeof(float));
    g->off_x = g->off_g + align256((size_t)g->n * sizeof(float));
    g->off_f = g->off_x + align256((size_t)g->world * 64 * sizeof(long long));
    g->bytes = g->off_f + align256((size_t)(2 * g->world + 2) * sizeof(unsigned int));
}

extern "C" {

int mre_peer_group_create(mre_ctx *ctx, int32_t rank, int32_t world, int64_t n_floats, mre_peer_group **out, unsigned char *handle_out) {
    MRE_CHECK_ARG(ctx && out && handle_out, "NULL argument");
    MRE_CHECK_ARG(world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world, "rank / world out of range (at most %d ranks)", PEER_MAX);
    MRE_CHECK_ARG(n_floats > 0 && n_floats % 4 == 0, "the buffers must hold a positive multiple of 4 floats");
    MRE_CUDA(cudaSetDevice(ctx->device));
    mre_peer_group *g = new mre_peer_group();
    g->rank = rank; g->world = world; g->device = ctx->device; g->n = n_floats;
    peer_layout(g);
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, g->bytes);
    if (e != cudaSuccess) { delete g; mre::set_error("cudaMalloc of the peer region failed: %s", cudaGetErrorString(e)); return MRE_ERR_CUDA; }
    cudaMemset(p, 0, g->bytes);
    cudaDeviceSynchronize();
    g->base[rank] = p;
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); delete g; mre::set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); return MRE_ERR_CUDA; }
    static_assert(sizeof(cudaIpcMemHandle_t) == MRE_PEER_HANDLE_BYTES, "handle size");
    memcpy(handle_out, &h, sizeof(h));
    *out = g;
    return MRE_OK;
}

int mre_peer_group_connect(mre_peer_group *g, const unsigned char *handles) {
    MRE_CHECK_ARG(g && handles, "NULL argument");
    MRE_CUDA(cudaSetDevice(g->device));
    for (int p = 0; p < g->world; p++) {
        if (p == g->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)p * MRE_PEER_HANDLE_BYTES, sizeof(h));
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { mre::set_error("cudaIpcOpenMemHandle of rank %d failed: %s", p, cudaGetErrorString(e)); return MRE_ERR_CUDA; }
        g->base[p] = ptr;
    }
    long long *x[PEER_MAX] = {nullptr};
    for (int p = 0; p < g->world; p++) x[p] = reinterpret_cast<long long *>(static_cast<char *>(g->base[p]) + g->off_x);
    MRE_CUDA(cudaMalloc((void **)&g->d_xchg, sizeof(x)));
    MRE_CUDA(cudaMemcpy(g->d_xchg, x, sizeof(x), cudaMemcpyHostToDevice));
    return MRE_OK;
}

/* test hook: a group that lives in ONE process (every "rank" on the same device); regions are handed over directly */
int mre_peer_group_connect_local(mre_peer_group *g, mre_peer_group *const *all) {
    MRE_CHECK_ARG(g && all, "NULL argument");
    for (int p = 0; p < g->world; p++) {
        MRE_CHECK_ARG(all[p] && all[p]->n == g->n && all[p]->world == g->world, "group %d does not match", p);
        g->base[p] = all[p]->base[p];
    }
    g->local = true;
    long long *x[PEER_MAX] = {nullptr};
    for (int p = 0; p < g->world; p++) x[p] = reinterpret_cast<long long *>(static_cast<char *>(g->base[p]) + g->off_x);
    MRE_CUDA(cudaSetDevice(g->device));
    MRE_CUDA(cudaMalloc((void **)&g->d_xchg, sizeof(x)));
    MRE_CUDA(cudaMemcpy(g->d_xchg, x, sizeof(x), cudaMemcpyHostToDevice));
    return MRE_OK;
}

void mre_peer_group_destroy(mre_peer_group *g) {
    if (!g) return;
    cudaSetDevice(g->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < g->world; p++)
        if (p != g->rank && g->base[p] && !g->local) cudaIpcCloseMemHandle(g->base[p]);
    cudaGetLastError();
    if (g->d_xchg) cudaFree(g->d_xchg);
    if (g->base[g->rank]) cudaFree(g->base[g->rank]);
    delete g;
}

float *mre_peer_weights(mre_peer_group *g) { return g ? static_cast<float *>(g->base[g->rank]) : nullptr; }
float *mre_peer_grads(mre_peer_group *g) { return g ? reinterpret_cast<float *>(static_cast<char *>(g->base[g->rank]) + g->off_g) : nullptr; }

static void fill_ptrs(const mre_peer_group *g, PeerPtrs &pp) {
    for (int p = 0; p < PEER_MAX; p++) {
        char *b = p < g->world ? static_cast<char *>(g->base[p]) : nullptr;
        pp.w[p] = reinterpret_cast<float *>(b);
        pp.g[p] = b ? reinterpret_cast<float *>(b + g->off_g) : nullptr;
        pp.flags[p] = b ? reinterpret_cast<unsigned int *>(b + g->off_f) : nullptr;
    }
}

int mre_dp_sgd_step(mre_ctx *ctx, mre_peer_group *g, float lr, int32_t max_blocks, void *stream) {
    MRE_CHECK_ARG(ctx && g, "NULL argument");
    MRE_CHECK_ARG(ctx->device == g->device, "the peer group lives on device %d", g->device);
    for (int p = 0; p < g->world; p++) MRE_CHECK_ARG(g->base[p] != nullptr, "rank %d's region is not connected", p);
    MRE_CUDA(cudaSetDevice(ctx->device));
    PeerPtrs pp;
    fill_ptrs(g, pp);
    g->step += 1;
    // every block spins on flags of other GPUs: the whole grid must be resident -- one block per SM
    int grid = ctx->sm_count;
    if (max_blocks > 0) grid = std::min(grid, max_blocks);
    dp_sgd_kernel<<<grid, PEER_THREADS, 0, (cudaStream_t)stream>>>(pp, g->rank, g->world, g->n, lr, g->step);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

int mre_peer_allreduce_i64(mre_ctx *ctx, mre_peer_group *g, int64_t *vec, int32_t count, void *stream) {
    MRE_CHECK_ARG(ctx && g && vec, "NULL argument");
    MRE_CHECK_ARG(count > 0 && count <= 64, "count must be in [1, 64]");
    MRE_CHECK_ARG(ctx->device == g->device && g->d_xchg, "the peer group is not connected on this device");
    MRE_CUDA(cudaSetDevice(ctx->device));
    PeerPtrs pp;
    fill_ptrs(g, pp);
    g->step += 1;
    peer_allreduce_i64_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(pp, g->d_xchg, g->rank, g->world, reinterpret_cast<long long *>(vec), count, g->step);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

/* 0 when no exchange of this group ever timed out (reads the error word; synchronises the stream's device) */
int mre_peer_group_error(mre_peer_group *g) {
    if (!g) return MRE_ERR_INVALID;
    cudaSetDevice(g->device);
    unsigned int e = 0;
    const unsigned int *f = reinterpret_cast<unsigned int *>(static_cast<char *>(g->base[g->rank]) + g->off_f) + 2 * g->world + 1;
    if (cudaMemcpy(&e, f, sizeof(e), cudaMemcpyDeviceToHost) != cudaSuccess) { mre::set_error("reading the peer error word failed"); return MRE_ERR_CUDA; }
    if (e) { mre::set_error("a peer exchange timed out: some rank never arrived"); return MRE_ERR_CUDA; }
    return MRE_OK;
}

}  // extern "C"
