// ZSL candidate scorer + ranker: Extractor (neighbour encoder + pair encoder + support encoder) -> cosine-mean against the
// generated relation vectors -> rank of the true candidate, for every test triple of a relation sweep in a few launches.
//
// Reference path replaced (paths relative to /root/reference):
//   module/zsl_module.py:46-59,61-67,69-106  Extractor.neighbor_encoder / entity_encoder / forward (query half)
//   module/submodule.py:240-258              SupportEncoder: LayerNorm(proj2(relu(proj1(x))) + x)
//   module/zsl_module.py:662-745             ZSLmodule.eval: per test triple build [C, 2] query pairs, gather the [C, 50, 2]
//                                            neighbour tensors of head and candidates, run the Extractor, sklearn
//                                            cosine_similarity(cand_vecs, relation_vecs).mean(1), argsort, rank of index 0
// The reference runs one Extractor forward per test triple, re-encoding the head's and every candidate's 50 neighbours each
// time.  Here the encoder is SPLIT where it is linear: reshape_layer([N_h | tanh fc1(h) | tanh fc2(c) | N_c]) = A_h + B_c with
// per-ENTITY halves A (its contribution as a pair's head) and B (as the candidate), computed once per entity; what remains
// per (head, candidate) pair is the 200 -> 400 -> 200 support encoder, LayerNorm and the cosine mean -- two FP32 tile GEMMs
// over all pairs of the sweep, the second with LayerNorm + cosine-mean fused in its epilogue (whole rows stay in one warp),
// and a per-triple compare/count.  FP32 throughout (the reference's arithmetic); scores agree to rounding.
#include <math.h>

#include <algorithm>

#include "common.h"
#include "device_utils.cuh"

namespace mre {

struct ZslDims {
    int D, H, D2;   // embedding / model dim (200), D / 2, 2 D
};

// ------------------------------------------------------------------------------------------ per-entity halves
// one CTA per entity; dot products are per-thread sequential fmaf over the weight row (tiny work: ~140 K MAC per entity)
__global__ void __launch_bounds__(128) zsl_entity_kernel(const mre_zsl_model m, const int64_t *__restrict__ ent_symbol,
                                                         const int64_t *__restrict__ conn, const float *__restrict__ deg,
                                                         int64_t n_ent, int max_nb, float *__restrict__ A, float *__restrict__ B) {
    extern __shared__ float sh[];
    const int D = (int)m.D, H = D / 2;
    float *s_sum = sh, *s_self = sh + D, *s_N = sh + 2 * D, *s_T1 = s_N + H, *s_T2 = s_T1 + H;
    const int64_t e = blockIdx.x;
    if (e >= n_ent) return;
    const int64_t sym = ent_symbol[e];
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < max_nb; j++) acc += m.symbol_emb[conn[e * max_nb + j] * D + d];   // pad id -> the zero row
        s_sum[d] = acc;
        s_self[d] = m.symbol_emb[sym * D + d];
    }
    __syncthreads();
    const float dg = deg[e];
    for (int o = threadIdx.x; o < H; o += blockDim.x) {
        float g = 0.f, a = 0.f, b = 0.f;
        for (int d = 0; d < D; d++) {
            g = fmaf(m.gcn_w[o * D + d], s_sum[d], g);
            a = fmaf(m.fc1_w[o * D + d], s_self[d], a);
            b = fmaf(m.fc2_w[o * D + d], s_self[d], b);
        }
        s_N[o] = tanhf((g + (float)max_nb * m.gcn_b[o]) / dg);      // every neighbour slot (pads too) carries the Linear's bias
        s_T1[o] = tanhf(a + m.fc1_b[o]);
        s_T2[o] = tanhf(b + m.fc2_b[o]);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
        const float *w = m.reshape_w + (int64_t)k * 2 * D;          // [D, 2 D]: columns [N_left | T1 | T2 | N_right]
        float a = 0.f, b = 0.f;
        for (int o = 0; o < H; o++) {
            a = fmaf(w[o], s_N[o], a);
            a = fmaf(w[H + o], s_T1[o], a);
            b = fmaf(w[2 * H + o], s_T2[o], b);
            b = fmaf(w[3 * H + o], s_N[o], b);
        }
        A[e * D + k] = a;
        B[e * D + k] = b + m.reshape_b[k];
    }
}

// pair -> triple (the candidate list it belongs to)
__global__ void zsl_pair_triple_kernel(const int64_t *__restrict__ cand_ptr, int64_t T, int64_t P, int32_t *__restrict__ pair_triple) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = T;                                     // last t with cand_ptr[t] <= p
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (cand_ptr[mid] <= p) lo = mid; else hi = mid;
        }
        pair_triple[p] = (int32_t)lo;
    }
}

// ||r_k|| of every generated relation vector
__global__ void zsl_relnorm_kernel(const float *__restrict__ rel_vecs, int64_t n, int D, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int d = 0; d < D; d++) s = fmaf(rel_vecs[i * D + d], rel_vecs[i * D + d], s);
    out[i] = sqrtf(s);
}

// ------------------------------------------------------------------------------------------ the two tile GEMMs
constexpr int ZBK = 16;

// 8 x 8 register micro-tile update from one k-slice held in shared memory (a: [ZBK][BM], b: [ZBK][BN], both k-major).
// Packed: the accumulators are 8 x 4 pairs of adjacent columns and one step is fma.rn.f32x2 (a, a) * (b_j, b_j+1) + acc
// (SASS FFMA2 with a scalar-broadcast operand): half the issue slots of 64 scalar FFMAs, same roundings.
__device__ __forceinline__ void ffma2(unsigned long long &acc, float a, unsigned long long b) {
    asm("{\n\t.reg .b64 t;\n\tmov.b64 t, {%1, %1};\n\tfma.rn.f32x2 %0, t, %2, %0;\n\t}" : "+l"(acc) : "f"(a), "l"(b));
}
template <int BM, int BN>
__device__ __forceinline__ void zsl_mma_tile(const float *__restrict__ sa, const float *__restrict__ sb, int ra, int cb,
                                             unsigned long long (&acc)[8][4]) {
#pragma unroll
    for (int k = 0; k < ZBK; k++) {
        const float4 a0 = *reinterpret_cast<const float4 *>(sa + k * BM + ra), a1 = *reinterpret_cast<const float4 *>(sa + k * BM + ra + 4);
        // the thread's 8 columns are cb .. cb + 3 and BN / 2 + cb .. + 3: each 128-bit load is contiguous across the lanes
        const ulonglong2 b0 = *reinterpret_cast<const ulonglong2 *>(sb + k * BN + cb), b1 = *reinterpret_cast<const ulonglong2 *>(sb + k * BN + BN / 2 + cb);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const unsigned long long b[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) ffma2(acc[i][j], a[i], b[j]);
    }
}
// column j of accumulator row i
__device__ __forceinline__ float zsl_acc(const unsigned long long (&acc)[8][4], int i, int j) {
    return __uint_as_float((j & 1) ? (uint32_t)(acc[i][j >> 1] >> 32) : (uint32_t)acc[i][j >> 1]);
}

// layer 1: Hid[p, n] = relu(sum_k X[p, k] W1[n, k] + b1[n]),  X[p, :] = A[head(p), :] + B[cand(p), :].  128 x 128 tiles.
__global__ void __launch_bounds__(256, 2) zsl_layer1_kernel(const mre_zsl_model m, const float *__restrict__ A, const float *__restrict__ B,
                                                         const int64_t *__restrict__ q_head, const int64_t *__restrict__ cand,
                                                         const int32_t *__restrict__ pair_triple, int64_t p0, int64_t P,
                                                         float *__restrict__ hid) {
    constexpr int BM = 128, BN = 128;
    __shared__ __align__(16) float sa[ZBK * BM], sb[ZBK * BN];
    const int D = (int)m.D, N = 2 * D;
    const int64_t row0 = p0 + (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    // loader role: row lr of the tile, k-quad lq (two float4 per thread and operand)
    const int lr = tid & 127, lq = (tid >> 7) * 2;      // a warp stores 32 consecutive rows of one k: conflict-free transposition
    const int64_t prow = row0 + lr;
    const bool row_ok = prow < p0 + P;
    const float *xa = nullptr, *xb = nullptr;
    if (row_ok) {
        xa = A + q_head[pair_triple[prow]] * D;
        xb = B + cand[prow] * D;
    }
    const int wn = n0 + lr;
    const float *wrow = wn < N ? m.proj1_w + (int64_t)wn * D : nullptr;
    unsigned long long acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0ull;
    // the next k-slice travels global -> registers while the current one is multiplied out of shared memory
    float4 nx[2], nw[2];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int k = k0 + (lq + u) * 4;
            nx[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            nw[u] = nx[u];
            if (row_ok && k < D) {
                const float4 p = *reinterpret_cast<const float4 *>(xa + k), q = *reinterpret_cast<const float4 *>(xb + k);
                nx[u] = make_float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w);
            }
            if (wrow && k < D) nw[u] = *reinterpret_cast<const float4 *>(wrow + k);
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int kk = (lq + u) * 4;
            sa[(kk + 0) * BM + lr] = nx[u].x; sa[(kk + 1) * BM + lr] = nx[u].y; sa[(kk + 2) * BM + lr] = nx[u].z; sa[(kk + 3) * BM + lr] = nx[u].w;
            sb[(kk + 0) * BN + lr] = nw[u].x; sb[(kk + 1) * BN + lr] = nw[u].y; sb[(kk + 2) * BN + lr] = nw[u].z; sb[(kk + 3) * BN + lr] = nw[u].w;
        }
    };
    fetch(0);
    stage();
    __syncthreads();
    for (int k0 = 0; k0 < D; k0 += ZBK) {
        const bool more = k0 + ZBK < D;
        if (more) fetch(k0 + ZBK);
        zsl_mma_tile<BM, BN>(sa, sb, ty * 8, tx * 4, acc);
        __syncthreads();
        if (more) {
            stage();
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int64_t p = row0 + ty * 8 + i;
        if (p >= p0 + P) continue;
        float *o = hid + (p - p0) * N + n0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int c = (j < 4 ? 0 : BN / 2) + tx * 4 + (j & 3);
            if (n0 + c < N) o[c] = fmaxf(zsl_acc(acc, i, j) + m.proj1_b[n0 + c], 0.f);
        }
    }
}

// layer 2 + LayerNorm + cosine mean: score[p] = mean_k cos(LN(Hid[p, :] W2^T + b2 + X[p, :]), r_k).  64 x 256 tiles: a warp owns
// 8 whole rows (lane = 8-column slice), so every row reduction is a warp shuffle.
__global__ void __launch_bounds__(256, 2) zsl_layer2_kernel(const mre_zsl_model m, const float *__restrict__ A, const float *__restrict__ B,
                                                         const int64_t *__restrict__ q_head, const int64_t *__restrict__ q_rel,
                                                         const int64_t *__restrict__ cand, const int32_t *__restrict__ pair_triple,
                                                         const float *__restrict__ hid, const float *__restrict__ rel_vecs,
                                                         const float *__restrict__ rel_norm, int n_vec, int64_t p0, int64_t P,
                                                         float *__restrict__ score) {
    constexpr int BM = 64, BN = 256;
    __shared__ __align__(16) float sa[ZBK * BM], sb[ZBK * BN];
    const int D = (int)m.D, K = 2 * D;
    const int64_t row0 = p0 + (int64_t)blockIdx.x * BM;
    const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;
    unsigned long long acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0ull;
    // loaders: hidden rows 64 x 16 = 256 float4 (one per thread); W2 rows 256 x 16 = 1024 float4 (four per thread)
    const int ar = tid & 63, aq = (tid >> 6) * 4;        // a warp stores 32 consecutive rows of one k
    const int64_t arow = row0 + ar;
    const float *hrow = arow < p0 + P ? hid + (arow - p0) * K : nullptr;
    float4 nh, nw[4];
    auto fetch = [&](int k0) {
        nh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hrow && k0 + aq < K) nh = *reinterpret_cast<const float4 *>(hrow + k0 + aq);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int n = tid, kq = u * 4;
            nw[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < D && k0 + kq < K) nw[u] = *reinterpret_cast<const float4 *>(m.proj2_w + (int64_t)n * K + k0 + kq);
        }
    };
    auto stage = [&]() {
        sa[(aq + 0) * BM + ar] = nh.x; sa[(aq + 1) * BM + ar] = nh.y; sa[(aq + 2) * BM + ar] = nh.z; sa[(aq + 3) * BM + ar] = nh.w;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int n = tid, kq = u * 4;
            sb[(kq + 0) * BN + n] = nw[u].x; sb[(kq + 1) * BN + n] = nw[u].y; sb[(kq + 2) * BN + n] = nw[u].z; sb[(kq + 3) * BN + n] = nw[u].w;
        }
    };
    fetch(0);
    stage();
    __syncthreads();
    for (int k0 = 0; k0 < K; k0 += ZBK) {
        const bool more = k0 + ZBK < K;
        if (more) fetch(k0 + ZBK);
        zsl_mma_tile<BM, BN>(sa, sb, ty * 8, tx * 4, acc);
        __syncthreads();
        if (more) {
            stage();
            __syncthreads();
        }
    }
    // ---- epilogue: this lane holds columns 4 tx .. 4 tx + 3 and 128 + 4 tx .. + 3 of eight rows
    const int ca = tx * 4, cb2 = BN / 2 + tx * 4;                  // first column of the lane's two 4-column groups
    const bool ok_a = ca < D, ok_b = cb2 < D;                      // D is a multiple of 8 (so of 4): a group is all in or all out
    float b2[8], g[8], be[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int c = (j < 4 ? ca : cb2) + (j & 3);
        const bool ok = j < 4 ? ok_a : ok_b;
        b2[j] = ok ? m.proj2_b[c] : 0.f;
        g[j] = ok ? m.ln_g[c] : 0.f;
        be[j] = ok ? m.ln_b[c] : 0.f;
    }
    const float inv_d = 1.f / (float)D;
#pragma unroll                                                      // (static indices: the accumulators stay in registers)
    for (int i = 0; i < 8; i++) {
        const int64_t p = row0 + ty * 8 + i;
        if (p >= p0 + P) continue;                                  // warp-uniform
        const int t = pair_triple[p];
        const float *xa = A + q_head[t] * D, *xb = B + cand[p] * D;
        float y[8], s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int c = (j < 4 ? ca : cb2) + (j & 3);
            const bool ok = j < 4 ? ok_a : ok_b;
            y[j] = ok ? zsl_acc(acc, i, j) + b2[j] + (xa[c] + xb[c]) : 0.f;
            s += y[j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mu = s * inv_d;
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float d = (j < 4 ? ok_a : ok_b) ? y[j] - mu : 0.f;
            v = fmaf(d, d, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const float rstd = rsqrtf(v * inv_d + m.ln_eps);
        float ln[8], nn = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            ln[j] = (j < 4 ? ok_a : ok_b) ? (y[j] - mu) * rstd * g[j] + be[j] : 0.f;
            nn = fmaf(ln[j], ln[j], nn);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
        const float *rv = rel_vecs + q_rel[t] * (int64_t)n_vec * D;
        const float *rn = rel_norm + q_rel[t] * (int64_t)n_vec;
        float total = 0.f;
        for (int k = 0; k < n_vec; k++) {
            float d = 0.f;
            if (ok_a) {
                const float4 r0 = *reinterpret_cast<const float4 *>(rv + (int64_t)k * D + ca);
                d = ln[0] * r0.x + ln[1] * r0.y + ln[2] * r0.z + ln[3] * r0.w;
            }
            if (ok_b) {
                const float4 r1 = *reinterpret_cast<const float4 *>(rv + (int64_t)k * D + cb2);
                d += ln[4] * r1.x + ln[5] * r1.y + ln[6] * r1.z + ln[7] * r1.w;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            const float den = sqrtf(nn) * rn[k];
            total += den > 0.f ? d / den : 0.f;
        }
        if (tx == 0) score[p] = total / (float)n_vec;
    }
}

// per test triple: how many candidates score higher than / equal to the true one (candidate 0 of its list); one warp per triple
__global__ void __launch_bounds__(256) zsl_count_kernel(const float *__restrict__ score, const int64_t *__restrict__ cand_ptr, int64_t T,
                                                        int32_t *__restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= T) return;
    const int64_t lo = cand_ptr[t], hi = cand_ptr[t + 1];
    int gt = 0, eq = 0;
    if (hi > lo) {
        const float s0 = score[lo];
        for (int64_t i = lo + 1 + lane; i < hi; i += 32) {
            gt += score[i] > s0 ? 1 : 0;
            eq += score[i] == s0 ? 1 : 0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        gt += __shfl_xor_sync(0xffffffffu, gt, o);
        eq += __shfl_xor_sync(0xffffffffu, eq, o);
    }
    if (lane == 0) {                                               // the [4][T] layout mre_metrics reads: raw_lt, raw_eq, filt_lt, filt_eq
        counts[t] = gt; counts[T + t] = eq; counts[2 * T + t] = gt; counts[3 * T + t] = eq;
    }
}

static int check_model(const mre_zsl_model *m) {
    MRE_CHECK_ARG(m != nullptr, "model is NULL");
    MRE_CHECK_ARG(m->D > 0 && m->D % 8 == 0 && m->D <= 256, "the model dimension must be a multiple of 8, at most 256 (reference: 200)");
    MRE_CHECK_ARG(m->symbol_emb && m->gcn_w && m->gcn_b && m->fc1_w && m->fc1_b && m->fc2_w && m->fc2_b && m->reshape_w && m->reshape_b &&
                      m->proj1_w && m->proj1_b && m->proj2_w && m->proj2_b && m->ln_g && m->ln_b,
                  "a model tensor is NULL");
    return MRE_OK;
}

int zsl_entity_features(mre_ctx *ctx, const mre_zsl_model *m, const int64_t *ent_symbol, const int64_t *conn, const float *deg,
                        int64_t n_ent, int32_t max_nb, float *A, float *B, cudaStream_t st) {
    MRE_TRY(check_model(m));
    MRE_CHECK_ARG(n_ent >= 0 && max_nb >= 0, "negative size");
    if (n_ent == 0) return MRE_OK;
    MRE_CHECK_ARG(ent_symbol && conn && deg && A && B, "NULL argument");
    const size_t smem = (size_t)(2 * m->D + 3 * (m->D / 2)) * sizeof(float);
    zsl_entity_kernel<<<(unsigned)n_ent, 128, smem, st>>>(*m, ent_symbol, conn, deg, n_ent, max_nb, A, B);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

int zsl_rank(mre_ctx *ctx, const mre_zsl_model *m, const float *A, const float *B, const int64_t *q_head, const int64_t *q_rel,
             const int64_t *cand_ptr, const int64_t *cand_idx, int64_t T, int64_t P, const float *rel_vecs, int64_t n_rel,
             int32_t n_vec, float *scores, int32_t *counts, cudaStream_t st) {
    MRE_TRY(check_model(m));
    MRE_CHECK_ARG(T >= 0 && P >= 0 && n_rel >= 0 && n_vec > 0, "bad size");
    if (T == 0) return MRE_OK;
    MRE_CHECK_ARG(A && B && q_head && q_rel && cand_ptr && counts && rel_vecs, "NULL argument");
    MRE_CHECK_ARG(P == 0 || cand_idx, "cand_idx is NULL");
    MRE_CHECK_ARG(P < (1LL << 31), "too many (head, candidate) pairs for one call");
    const int D = (int)m->D;
    const int64_t chunk = std::min<int64_t>(std::max<int64_t>(P, 1), 1 << 18);
    // scratch: pair -> triple map, relation-vector norms, scores (when the caller does not want them), hidden activations
    MRE_TRY(ctx->misc2.reserve((size_t)std::max<int64_t>(P, 1) * (sizeof(int32_t) + sizeof(float)) + (size_t)n_rel * n_vec * sizeof(float) + 64));
    int32_t *pair_triple = ctx->misc2.as<int32_t>();
    float *sc = scores ? scores : reinterpret_cast<float *>(pair_triple + std::max<int64_t>(P, 1));
    float *rel_norm = reinterpret_cast<float *>(pair_triple + std::max<int64_t>(P, 1)) + std::max<int64_t>(P, 1);
    MRE_TRY(ctx->ent_aux.reserve((size_t)chunk * 2 * D * sizeof(float)));
    float *hid = ctx->ent_aux.as<float>();
    if (P > 0) zsl_pair_triple_kernel<<<(unsigned)std::min<int64_t>((P + 255) / 256, 148 * 16), 256, 0, st>>>(cand_ptr, T, P, pair_triple);
    if (n_rel > 0) zsl_relnorm_kernel<<<(unsigned)((n_rel * n_vec + 127) / 128), 128, 0, st>>>(rel_vecs, n_rel * n_vec, D, rel_norm);
    ctx->launches += 2;
    MRE_TRY(ctx->time_begin(st));
    for (int64_t p0 = 0; p0 < P; p0 += chunk) {
        const int64_t n = std::min(chunk, P - p0);
        dim3 g1((unsigned)((n + 127) / 128), (unsigned)((2 * D + 127) / 128));
        zsl_layer1_kernel<<<g1, 256, 0, st>>>(*m, A, B, q_head, cand_idx, pair_triple, p0, n, hid);
        zsl_layer2_kernel<<<(unsigned)((n + 63) / 64), 256, 0, st>>>(*m, A, B, q_head, q_rel, cand_idx, pair_triple, hid, rel_vecs, rel_norm,
                                                                      n_vec, p0, n, sc);
        ctx->launches += 2;
    }
    MRE_TRY(ctx->time_end(st));
    zsl_count_kernel<<<(unsigned)((T + 7) / 8), 256, 0, st>>>(sc, cand_ptr, T, counts);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

}  // namespace mre
