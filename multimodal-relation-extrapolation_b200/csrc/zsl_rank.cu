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
// per-ENTITY halves A (its contribution as a pair's head) and B (as the candidate), computed once per entity; proj1 splits the
// same way (A1 = W1 A + b1, B1 = W1 B), and the cosine mean is one dot product with sum_k r_k / ||r_k||.  What remains per
// (head, candidate) pair -- relu(A1_h + B1_c) W2^T, the residual, LayerNorm, that dot product -- is zsl_tc_kernel below: a
// 3 x TF32 tcgen05 contraction over CTA pairs with the epilogue read out of TMEM (FP32-level accuracy: scores within 5e-8 of
// the reference's), then a per-triple compare/count.  The FP32 CUDA-core kernels the scorer started as (zsl_layer1_kernel in
// pair mode + zsl_layer2_kernel: two tile GEMMs over all pairs) serve model widths the tensor-core tiles do not fit and, behind
// MRE_DEV_ZSL_FP32=1, as the full-size cross-check of tests/test_zsl.py.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.h"
#include "device_utils.cuh"
#include "tma_host.h"

namespace mre {

struct ZslDims {
    int D, H, D2;   // embedding / model dim (200), D / 2, 2 D
};

// ------------------------------------------------------------------------------------------ per-entity halves
// ZE_TILE entities per CTA: every weight element a thread reads feeds ZE_TILE accumulators (the weights, 560 KB, do not fit
// L1: one entity per CTA re-read them from L2 14 208 times and took 4.5 ms).  Dot products stay per-thread sequential fmaf
// over the weight row, in the same order for every entity, so the halves do not depend on the tiling.
constexpr int ZE_TILE = 8;
__global__ void __launch_bounds__(128) zsl_entity_kernel(const mre_zsl_model m, const int64_t *__restrict__ ent_symbol,
                                                         const int64_t *__restrict__ conn, const float *__restrict__ deg,
                                                         int64_t n_ent, int max_nb, float *__restrict__ A, float *__restrict__ B) {
    extern __shared__ __align__(16) float sh[];
    const int D = (int)m.D, H = D / 2;
    float *s_sum = sh, *s_self = sh + ZE_TILE * D, *s_N = sh + 2 * ZE_TILE * D, *s_T1 = s_N + ZE_TILE * H, *s_T2 = s_T1 + ZE_TILE * H;
    const int64_t e0 = (int64_t)blockIdx.x * ZE_TILE;
    const int ne = (int)min((int64_t)ZE_TILE, n_ent - e0);
    for (int i = threadIdx.x; i < ZE_TILE * D; i += blockDim.x) {
        const int te = i / D, d = i - te * D;
        float acc = 0.f, self = 0.f;
        if (te < ne) {
            const int64_t e = e0 + te;
            for (int j = 0; j < max_nb; j++) acc += m.symbol_emb[conn[e * max_nb + j] * D + d];   // pad id -> the zero row
            self = m.symbol_emb[ent_symbol[e] * D + d];
        }
        s_sum[i] = acc;
        s_self[i] = self;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < H; o += blockDim.x) {
        float g[ZE_TILE], a[ZE_TILE], b[ZE_TILE];
#pragma unroll
        for (int te = 0; te < ZE_TILE; te++) g[te] = a[te] = b[te] = 0.f;
        for (int d = 0; d < D; d += 4) {
            const float4 wg = *reinterpret_cast<const float4 *>(m.gcn_w + o * D + d), wa = *reinterpret_cast<const float4 *>(m.fc1_w + o * D + d);
            const float4 wb = *reinterpret_cast<const float4 *>(m.fc2_w + o * D + d);
#pragma unroll
            for (int te = 0; te < ZE_TILE; te++) {
                const float4 ss = *reinterpret_cast<const float4 *>(s_sum + te * D + d), sf = *reinterpret_cast<const float4 *>(s_self + te * D + d);
                g[te] = fmaf(wg.w, ss.w, fmaf(wg.z, ss.z, fmaf(wg.y, ss.y, fmaf(wg.x, ss.x, g[te]))));
                a[te] = fmaf(wa.w, sf.w, fmaf(wa.z, sf.z, fmaf(wa.y, sf.y, fmaf(wa.x, sf.x, a[te]))));
                b[te] = fmaf(wb.w, sf.w, fmaf(wb.z, sf.z, fmaf(wb.y, sf.y, fmaf(wb.x, sf.x, b[te]))));
            }
        }
#pragma unroll
        for (int te = 0; te < ZE_TILE; te++) {
            const float dg = te < ne ? deg[e0 + te] : 1.f;
            s_N[te * H + o] = tanhf((g[te] + (float)max_nb * m.gcn_b[o]) / dg);   // every neighbour slot (pads too) carries the Linear's bias
            s_T1[te * H + o] = tanhf(a[te] + m.fc1_b[o]);
            s_T2[te * H + o] = tanhf(b[te] + m.fc2_b[o]);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
        const float *w = m.reshape_w + (int64_t)k * 2 * D;          // [D, 2 D]: columns [N_left | T1 | T2 | N_right]
        float a[ZE_TILE], b[ZE_TILE];
#pragma unroll
        for (int te = 0; te < ZE_TILE; te++) a[te] = b[te] = 0.f;
        for (int o = 0; o < H; o++) {
            const float w0 = w[o], w1 = w[H + o], w2 = w[2 * H + o], w3 = w[3 * H + o];
#pragma unroll
            for (int te = 0; te < ZE_TILE; te++) {
                const float n = s_N[te * H + o];
                a[te] = fmaf(w1, s_T1[te * H + o], fmaf(w0, n, a[te]));
                b[te] = fmaf(w3, n, fmaf(w2, s_T2[te * H + o], b[te]));
            }
        }
#pragma unroll
        for (int te = 0; te < ZE_TILE; te++)
            if (te < ne) {
                A[(e0 + te) * D + k] = a[te];
                B[(e0 + te) * D + k] = b[te] + m.reshape_b[k];
            }
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
                                                         float *__restrict__ hid, int relu, int bias) {
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
        if (pair_triple) {
            xa = A + q_head[pair_triple[prow]] * D;
            xb = B + cand[prow] * D;
        } else {
            xa = A + prow * D;                                      // table mode: row p of A alone (the per-entity hidden halves)
        }
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
                const float4 p = *reinterpret_cast<const float4 *>(xa + k);
                const float4 q = xb ? *reinterpret_cast<const float4 *>(xb + k) : make_float4(0.f, 0.f, 0.f, 0.f);
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
            if (n0 + c < N) {
                const float v = zsl_acc(acc, i, j) + (bias ? m.proj1_b[n0 + c] : 0.f);
                o[c] = relu ? fmaxf(v, 0.f) : v;
            }
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

// ------------------------------------------------------------------------------------------ tensor-core path
// proj1 is linear too, so the hidden layer splits the same way the pair vector does: relu(W1 (A_h + B_c) + b1) =
// relu(A1_h + B1_c) with per-ENTITY rows A1 = W1 A + b1, B1 = W1 B (two small FP32 GEMMs over the entity table).  What is left
// per pair is ONE contraction, Hid[p, :] W2^T (K = 2 D = 400, N = D = 200), run on the 5th-generation tensor cores as 3 x TF32:
// every FP32 operand is split hi + lo (both round-to-nearest TF32; hi + lo carries 22 significand bits) and the product is
// hi*hi + hi*lo + lo*hi with FP32 accumulation in TMEM -- the dropped lo*lo term is 2^-22 relative, FP32-level accuracy.
// W2 carries one extra output row, its column sums, so accumulator column D is sum_d y[d]: the LayerNorm mean is known before
// the row is read, and the whole epilogue is ONE pass over TMEM on centred values d = u - mu:
//     var = sum d^2 / D,  z = d rstd g + be,  z . rsum = rstd sum (d g) rsum + be . rsum,  |z|^2 = rstd^2 sum (d g)^2 + 2 rstd sum (d g) be + |be|^2
// (rsum = sum_k r_k / ||r_k||: the cosine mean is linear in the normalised relation vectors).
//   Two CTAs (a cluster, cta_group::2) walk 256-pair tiles together: each stages its own 128 pair rows and HALF of every W2
//   k-block (NP / 2 rows x 16 floats, hi and lo), one tcgen05.mma of the leader drives both SMs' tensor cores (M = 256), and each
//   CTA keeps two 128 x NP accumulators in TMEM, so the epilogue of tile i runs under the MMAs of tile i + 1.  17 warps per CTA:
//     warps 0-7     operand producers: four lanes share a pair row (one 16-byte quarter of its 64-byte k-block each); gather
//                   A1[head], B1[cand] (L2-resident), add, relu, split, store the hi / lo rows in the canonical K-major
//                   SWIZZLE_64B layout (the contraction's left operand never sees HBM), then arrive on the LEADER's full
//                   barrier; lane 0 of warp 0 also issues the TMA loads of this CTA's W2 half
//     warps 8-15    epilogue, thread = pair row (TMEM lane), two warps per 32-lane quarter splitting the columns; partial sums
//                   meet in shared memory
//     warp 16       MMA issuer (leader CTA): 6 tcgen05.mma per k-block (2 k-steps x 3 split products), commits multicast to both
#ifndef MRE_ZT_STAGES
#define MRE_ZT_STAGES 4
#endif
#ifndef MRE_ZT_EPI_WARPS
#define MRE_ZT_EPI_WARPS 4      // 4: one epilogue warp per TMEM lane quarter, all columns (13 warps: 128 registers per thread); 8: two, splitting the columns
#endif
constexpr int ZT_ROWS = 256, ZT_CTA_ROWS = 128, ZT_BK = 16, ZT_STAGES = MRE_ZT_STAGES, ZT_PROD_WARPS = 8, ZT_EPI_WARPS = MRE_ZT_EPI_WARPS;
constexpr int ZT_HALVES = ZT_EPI_WARPS / 4;
static_assert(ZT_EPI_WARPS == 4 || ZT_EPI_WARPS == 8, "one or two epilogue warps per TMEM lane quarter");
constexpr int ZT_PASSES = ZT_CTA_ROWS / (ZT_PROD_WARPS * 8);      // a producer warp covers its rows 8 at a time
constexpr int ZT_THREADS = 32 * (ZT_PROD_WARPS + ZT_EPI_WARPS + 1);
#ifndef MRE_ZT_DIAG
#define MRE_ZT_DIAG 0      // developer timing variants: 1 no residual gather, 4 no B1 gather, 16 no epilogue work
#endif
constexpr int ZT_XPITCH = 36;                                      // floats per staged residual row (32 + 4: conflict-free both ways)
constexpr int ZT_A_BYTES = ZT_CTA_ROWS * ZT_BK * 4;                // one of the hi / lo operand tiles of a stage (this CTA's rows)

// The epilogue's three per-column vectors (proj2 bias, LayerNorm gain and bias) are the same for every row of every tile.  Read from
// shared memory they cost two L1TEX data-pipe wavefronts per warp-uniform LDS.128 (165 M per sweep on the pipe that bounds the
// kernel, profiles/r2_zsl_gather_experiments.md); from the constant bank they are LDC reads through the constant cache and cost
// the data pipe nothing.  One slot per context (contexts are not shared between concurrently running streams -- the same contract
// as the context's scratch buffers), filled by a stream-ordered device-to-device copy before the launch.  Sixteen slots, handed out
// round-robin at mre_ctx_create: seventeen or more contexts scoring DIFFERENT models at the same time on one device would share one.
#ifndef MRE_ZT_CONST_VEC
#define MRE_ZT_CONST_VEC 1
#endif
constexpr int ZT_CONST_SLOTS = 16, ZT_CONST_D = 224;
__constant__ float c_zsl_vec[ZT_CONST_SLOTS][3][ZT_CONST_D];

struct ZslTcParams {
    const float *A1, *B1, *A, *B, *sA, *sB;                         // sA / sB: row sums of A / B
    const int64_t *q_head, *q_rel, *cand;
    const int32_t *pair_triple;
    const float *rsum, *bR;                                         // [n_rel, D]: sum_k r_k / ||r_k||;  [n_rel]: ln_b . rsum
    const float *b2, *ln_g, *ln_b;
    float ln_eps, inv_nvec;
    int D, K, NP, nkb, cslot;
    uint32_t idesc;
    int64_t P, tiles;
    float *score;
};

// W2 -> TF32 hi / lo; row D holds the column sums (accumulator column D = sum of the row's outputs), the rest of the padding zeros
__global__ void zsl_split_w2_kernel(const float *__restrict__ w, int D, int K, int NP, float *__restrict__ hi, float *__restrict__ lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NP * K) return;
    const int row = i / K, k = i - row * K;
    float v = 0.f;
    if (row < D) v = w[i];
    else if (row == D)
        for (int d = 0; d < D; d++) v += w[d * K + k];
    const float h = tf32_rna(v);
    hi[i] = h;
    lo[i] = tf32_rna(v - h);
}

// out[i] = sum of row i (one warp per row, fixed order)
__global__ void zsl_rowsum_kernel(const float *__restrict__ tab, int64_t n, int D, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    float s = 0.f;
    for (int d = threadIdx.x & 31; d < D; d += 32) s += tab[i * D + d];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) out[i] = s;
}

// rsum[r, :] = sum_k rel_vecs[r, k, :] / ||rel_vecs[r, k, :]||   (zero vectors contribute nothing, as sklearn's normalize leaves
// them);  bR[r] = ln_b . rsum[r, :]
__global__ void zsl_relsum_kernel(const float *__restrict__ rel_vecs, int n_vec, int D, const float *__restrict__ ln_b,
                                  float *__restrict__ rsum, float *__restrict__ bR) {
    __shared__ float s_inv[64];
    const int64_t r = blockIdx.x;
    const float *rv = rel_vecs + r * (int64_t)n_vec * D;
    for (int k0 = 0; k0 < n_vec; k0 += 64) {
        const int nk = min(64, n_vec - k0);
        __syncthreads();
        if ((int)threadIdx.x < nk) {
            float s = 0.f;
            for (int d = 0; d < D; d++) s = fmaf(rv[(int64_t)(k0 + threadIdx.x) * D + d], rv[(int64_t)(k0 + threadIdx.x) * D + d], s);
            const float n = sqrtf(s);
            s_inv[threadIdx.x] = n > 0.f ? 1.f / n : 0.f;
        }
        __syncthreads();
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            float acc = k0 ? rsum[r * D + d] : 0.f;
            for (int k = 0; k < nk; k++) acc = fmaf(rv[(int64_t)(k0 + k) * D + d], s_inv[k], acc);
            rsum[r * D + d] = acc;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int d = 0; d < D; d++) s = fmaf(ln_b[d], rsum[r * D + d], s);
        bR[r] = s;
    }
}

// round-to-nearest (ties away) TF32 of a finite non-negative float in two integer ops (cvt.rna.tf32 adds an inf / nan guard)
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }

// D[tmem of both CTAs] (+)= A * B^T, kind::tf32, over a CTA pair (M = 256, each CTA holds 128 rows of A / D and half of B's rows)
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// (launch bound 512 with 13 warps: 4 warps at most share a scheduler's 512 registers per lane -> 128 per thread; 17 warps: 96)
__global__ void __launch_bounds__(ZT_THREADS > 512 ? ZT_THREADS : 512, 1) zsl_tc_kernel(const __grid_constant__ CUtensorMap tm_hi,
                                                               const __grid_constant__ CUtensorMap tm_lo, const ZslTcParams p) {
    extern __shared__ uint8_t zt_smem[];
    __shared__ __align__(8) uint64_t bars[2 * ZT_STAGES + 4];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_vec[3][256];
    __shared__ float s_scal[2];
    __shared__ __align__(16) float s_stage[ZT_EPI_WARPS][32][ZT_XPITCH];   // per epilogue warp: 32 rows x 32 residual columns, transposed on the way
    __shared__ int s_cand[ZT_EPI_WARPS][32];
    __shared__ __align__(16) float4 s_part[2][4][32];                      // [tile parity][quarter][row]: the upper column half's partial sums
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int64_t worker = blockIdx.x >> 1, n_workers = gridDim.x >> 1;
    const uint32_t base = (smem_u32(zt_smem) + 1023u) & ~1023u;
    const uint32_t w_bytes = (uint32_t)(p.NP / 2) * ZT_BK * 4, stage_bytes = 2 * ZT_A_BYTES + 2 * w_bytes;
    auto full = [&](int s) { return smem_u32(&bars[s]); };
    auto empty = [&](int s) { return smem_u32(&bars[ZT_STAGES + s]); };
    auto acc_full = [&](int b) { return smem_u32(&bars[2 * ZT_STAGES + b]); };
    auto acc_empty = [&](int b) { return smem_u32(&bars[2 * ZT_STAGES + 2 + b]); };
    if (tid == 0) {
        for (int s = 0; s < ZT_STAGES; s++) {
            mbar_init(full(s), 2 * ZT_PROD_WARPS + 1);              // (leader's copy is the one used) both CTAs' producer warps + the expect_tx arrive
            mbar_init(empty(s), 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(acc_full(b), 1);
            mbar_init(acc_empty(b), 2 * ZT_EPI_WARPS);              // (leader's copy) the epilogue warps of both CTAs
        }
        fence_barrier_init();
        tma_prefetch_desc(&tm_hi);
        tma_prefetch_desc(&tm_lo);
    }
    for (int i = tid; i < 256; i += ZT_THREADS) {
        s_vec[0][i] = i < p.D ? p.b2[i] : 0.f;
        s_vec[1][i] = i < p.D ? p.ln_g[i] : 0.f;
        s_vec[2][i] = i < p.D ? p.ln_b[i] : 0.f;
    }
    if (warp == 1) {                                                // sum b2 and |ln_b|^2
        float a = 0.f, b = 0.f;
        for (int d = lane; d < p.D; d += 32) {
            a += p.b2[d];
            b = fmaf(p.ln_b[d], p.ln_b[d], b);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (lane == 0) { s_scal[0] = a; s_scal[1] = b; }
    }
    constexpr int MMA_WARP = ZT_PROD_WARPS + ZT_EPI_WARPS;
    if (warp == MMA_WARP) tmem_alloc_pair(smem_u32(&tmem_slot), 512);
    tc_fence_before();
    cluster_sync_all();                                             // both CTAs' barriers are initialised before either signals the other's
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp < ZT_PROD_WARPS) {
        // ---- operand producers (+ this CTA's W2 half)
        const int rr = lane >> 2, j = lane & 3;
        const uint32_t sw = (uint32_t)(rr >> 1) & 3u;
        const uint32_t row_off = (uint32_t)(warp * 8 * ZT_PASSES + rr) * (ZT_BK * 4) + (((uint32_t)j ^ sw) << 4);
        const uint32_t full_leader0 = mapa_shared(full(0), 0);
        int s = 0;
        uint32_t ph = 0;
        const float4 *a1[ZT_PASSES], *b1[ZT_PASSES], *na1[ZT_PASSES], *nb1[ZT_PASSES];
        auto row_ptrs = [&](int64_t tile, const float4 *(&pa)[ZT_PASSES], const float4 *(&pb)[ZT_PASSES]) {
#pragma unroll
            for (int i = 0; i < ZT_PASSES; i++) {
                const int64_t pr = min(tile * ZT_ROWS + rank * ZT_CTA_ROWS + warp * 8 * ZT_PASSES + i * 8 + rr, p.P - 1);   // rows past the end recompute the last pair
                pa[i] = reinterpret_cast<const float4 *>(p.A1 + p.q_head[p.pair_triple[pr]] * p.K) + j;
                pb[i] = reinterpret_cast<const float4 *>(p.B1 + p.cand[pr] * p.K) + j;
            }
        };
        if (worker < p.tiles) row_ptrs(worker, a1, b1);
        for (int64_t tile = worker; tile < p.tiles; tile += n_workers) {
            if (tile + n_workers < p.tiles) row_ptrs(tile + n_workers, na1, nb1);   // the next tile's index chain resolves under this tile
            float4 bq[4][2 * ZT_PASSES];                           // [.][2 i] = the candidate's quarter k-block, [.][2 i + 1] = the head's
            auto ld = [&](float4 (&d)[2 * ZT_PASSES], int kb) {
#pragma unroll
                for (int i = 0; i < ZT_PASSES; i++) {
                    d[2 * i] = (MRE_ZT_DIAG & 4) ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(b1[i] + kb * 4);
                    d[2 * i + 1] = __ldg(a1[i] + kb * 4);
                }
            };
            auto put = [&](const float4 (&q)[2 * ZT_PASSES], int kb) {
                mbar_wait(empty(s), ph ^ 1);
                if (warp == 0 && lane == 0) {                       // the bytes of both CTAs' halves are counted by the leader's barrier
                    if (rank == 0) mbar_arrive_expect_tx(full(s), 4 * w_bytes);
                    const uint32_t w = base + s * stage_bytes + 2 * ZT_A_BYTES, fl = full_leader0 + 8 * s;
                    tma_load_2d_pair(w, &tm_hi, kb * ZT_BK, (int)rank * (p.NP / 2), fl, L2_EVICT_LAST);
                    tma_load_2d_pair(w + w_bytes, &tm_lo, kb * ZT_BK, (int)rank * (p.NP / 2), fl, L2_EVICT_LAST);
                }
                const uint32_t a_hi = base + s * stage_bytes + row_off, a_lo = a_hi + ZT_A_BYTES;
#pragma unroll
                for (int i = 0; i < ZT_PASSES; i++) {
                    const float v0 = fmaxf(q[2 * i + 1].x + q[2 * i].x, 0.f), v1 = fmaxf(q[2 * i + 1].y + q[2 * i].y, 0.f);
                    const float v2 = fmaxf(q[2 * i + 1].z + q[2 * i].z, 0.f), v3 = fmaxf(q[2 * i + 1].w + q[2 * i].w, 0.f);
                    const float h0 = tf32_hi(v0), h1 = tf32_hi(v1), h2 = tf32_hi(v2), h3 = tf32_hi(v3);
                    const uint32_t off = (uint32_t)i * 8 * (ZT_BK * 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + off), "f"(h0), "f"(h1), "f"(h2), "f"(h3) : "memory");
                    // lo = v - hi is exact in FP32; the tensor core reads its leading 11 significand bits (2^-21 |v| at worst)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + off), "f"(v0 - h0), "f"(v1 - h1), "f"(v2 - h2), "f"(v3 - h3)
                                 : "memory");
                }
                fence_proxy_async();                                // generic-proxy stores -> visible to the tensor core's async proxy
                __syncwarp();
                // (relaxed: a release.cluster arrive costs a MEMBAR.ALL.GPU per k-block; the proxy fence above has already made
                // the rows visible to the tensor core before the arrive is issued)
                if (lane == 0) mbar_arrive_cluster_relaxed(full_leader0 + 8 * s);
                if (++s == ZT_STAGES) { s = 0; ph ^= 1; }
            };
            ld(bq[0], 0);
            if (p.nkb > 1) ld(bq[1], 1);
            if (p.nkb > 2) ld(bq[2], 2);
            for (int kb = 0; kb < p.nkb; kb += 4) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (kb + u < p.nkb) {
                        if (kb + u + 3 < p.nkb) ld(bq[(u + 3) % 4], kb + u + 3);
                        put(bq[u], kb + u);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < ZT_PASSES; i++) { a1[i] = na1[i]; b1[i] = nb1[i]; }
        }
    } else if (warp == MMA_WARP) {
        // ---- MMA issuer: only the leader CTA issues, one instruction drives both SMs' tensor cores
        if (rank == 0 && lane == 0) {
            int s = 0;
            uint32_t ph = 0, n = 0;
            for (int64_t tile = worker; tile < p.tiles; tile += n_workers, n++) {
                const uint32_t buf = n & 1;
                mbar_wait(acc_empty(buf), ((n >> 1) & 1) ^ 1);      // both CTAs' epilogues have drained this accumulator buffer
                tc_fence_after();
                const uint32_t d = tmem + buf * (uint32_t)p.NP;
                for (int kb = 0; kb < p.nkb; kb++) {
                    mbar_wait(full(s), ph);
                    tc_fence_after();
                    const uint32_t a_hi = base + s * stage_bytes, a_lo = a_hi + ZT_A_BYTES, w_hi = a_hi + 2 * ZT_A_BYTES, w_lo = w_hi + w_bytes;
#pragma unroll
                    for (int kk = 0; kk < ZT_BK / 8; kk++) {
                        const uint64_t da_hi = umma_desc_k64(a_hi) + kk * 2, da_lo = umma_desc_k64(a_lo) + kk * 2;
                        const uint64_t db_hi = umma_desc_k64(w_hi) + kk * 2, db_lo = umma_desc_k64(w_lo) + kk * 2;
                        umma_tf32_pair(d, da_lo, db_hi, p.idesc, (kb | kk) != 0);     // the small cross terms first
                        umma_tf32_pair(d, da_hi, db_lo, p.idesc, 1);
                        umma_tf32_pair(d, da_hi, db_hi, p.idesc, 1);
                    }
                    umma_commit_pair(empty(s), 3);                  // the stage is reusable in both CTAs once these MMAs retire
                    if (++s == ZT_STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit_pair(acc_full(buf), 3);
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue: thread = TMEM lane = pair row; the two warps of a lane quarter split the columns
        // (17 warps cap the kernel at 96 registers per thread -- 5 warps on one scheduler; setmaxnreg.inc cannot help: it draws
        // from the CTA's own pool, and the gather warps have nothing to give back)
        const int ew = warp - ZT_PROD_WARPS, q = warp & 3, half = ew >> 2, r = q * 32 + lane;
        const int nch = p.D / 8, c_split = ((nch + 1) / 2) * 8;
        const int c_lo = half ? c_split : 0, c_hi = (half || ZT_HALVES == 1) ? p.D : c_split;
        const int ngrp = (c_hi - c_lo + 31) / 32;
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        const float inv_d = 1.f / (float)p.D;
        float (*stg)[ZT_XPITCH] = s_stage[ew];
        const int lr = lane >> 3, lc = (lane & 7) * 4;             // coalesced residual loads: 8 lanes cover 128 bytes of one row
        const uint32_t acc_empty_leader0 = mapa_shared(acc_empty(0), 0);
        uint32_t n = 0;
        for (int64_t tile = worker; tile < p.tiles; tile += n_workers, n++) {
            const uint32_t buf = n & 1;
            const int64_t pr0 = tile * ZT_ROWS + rank * ZT_CTA_ROWS + r, pr = min(pr0, p.P - 1);
            const int t = p.pair_triple[pr];
            const int64_t head = p.q_head[t], cnd = p.cand[pr], rel = p.q_rel[t];
            const float4 *xa = reinterpret_cast<const float4 *>(p.A + head * p.D);
            const float4 *rs = reinterpret_cast<const float4 *>(p.rsum + rel * p.D);
            const float sum_x = p.sA[head] + p.sB[cnd], b_r = p.bR[rel];
            __syncwarp();
            s_cand[ew][lane] = (int)cnd;
            __syncwarp();
            float4 xr[8];                                           // the candidates' residual columns of the next 32-column group
            auto ldx = [&](int g) {
#pragma unroll
                for (int i = 0; i < 8; i++)
                    xr[i] = (!(MRE_ZT_DIAG & 1) && c_lo + g * 32 + lc < c_hi)
                                ? __ldg(reinterpret_cast<const float4 *>(p.B + (int64_t)s_cand[ew][4 * i + lr] * p.D + (c_lo + g * 32 + lc)))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            ldx(0);
            mbar_wait(acc_full(buf), (n >> 1) & 1);
            tc_fence_after();
            const uint32_t t0 = t_lane + buf * (uint32_t)p.NP;
            float var = 0.f, q2 = 0.f, sgb = 0.f, dotp = 0.f;
            if (!(MRE_ZT_DIAG & 16)) {
                float mu;
                {
                    uint32_t v[8];
                    tmem_ld_32x8(t0 + p.D, v);                      // column D: sum_d (Hid W2^T)[d], from the column-sum row of W2
                    tmem_ld_wait();
                    mu = ((__uint_as_float(v[0]) + s_scal[0]) + sum_x) * inv_d;
                }
                for (int g = 0; g < ngrp; g++) {
                    const int c0 = c_lo + g * 32;
                    __syncwarp();                                   // the previous group's rows have been read
#pragma unroll
                    for (int i = 0; i < 8; i++) *reinterpret_cast<float4 *>(&stg[4 * i + lr][lc]) = xr[i];
                    __syncwarp();
                    if (g + 1 < ngrp) ldx(g + 1);
                    uint32_t v[4][8];
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (c0 + 8 * k < c_hi) tmem_ld_32x8(t0 + c0 + 8 * k, v[k]);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (c0 + 8 * k < c_hi) {
#pragma unroll
                            for (int h = 0; h < 2; h++) {
                                const int c = c0 + 8 * k + 4 * h;
                                const float4 a = (MRE_ZT_DIAG & 1) ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(xa + c / 4);
                                const float4 b = *reinterpret_cast<const float4 *>(&stg[lane][8 * k + 4 * h]), rr4 = __ldg(rs + c / 4);
#if MRE_ZT_CONST_VEC
                                const float4 b2 = *reinterpret_cast<const float4 *>(&c_zsl_vec[p.cslot][0][c]);
                                const float4 gg = *reinterpret_cast<const float4 *>(&c_zsl_vec[p.cslot][1][c]);
                                const float4 be = *reinterpret_cast<const float4 *>(&c_zsl_vec[p.cslot][2][c]);
#else
                                const float4 b2 = *reinterpret_cast<const float4 *>(&s_vec[0][c]);
                                const float4 gg = *reinterpret_cast<const float4 *>(&s_vec[1][c]);
                                const float4 be = *reinterpret_cast<const float4 *>(&s_vec[2][c]);
#endif
                                const float xs[4] = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
                                const float b2s[4] = {b2.x, b2.y, b2.z, b2.w}, gs[4] = {gg.x, gg.y, gg.z, gg.w};
                                const float bes[4] = {be.x, be.y, be.z, be.w}, rrs[4] = {rr4.x, rr4.y, rr4.z, rr4.w};
#pragma unroll
                                for (int i = 0; i < 4; i++) {
                                    const float u = (__uint_as_float(v[k][4 * h + i]) + b2s[i]) + xs[i];
                                    const float d = u - mu, e = d * gs[i];
                                    var = fmaf(d, d, var);
                                    q2 = fmaf(e, e, q2);
                                    sgb = fmaf(e, bes[i], sgb);
                                    dotp = fmaf(e, rrs[i], dotp);
                                }
                            }
                        }
                }
            } else {
                dotp = xr[0].x;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(acc_empty_leader0 + 8 * buf);   // this warp is done with the accumulator buffer
            // the upper column half hands its partial sums to the lower one (buffers alternate with the tile parity: the
            // writer can be at most one tile ahead of the reader)
            if (ZT_HALVES == 2 && half) s_part[n & 1][q][lane] = make_float4(var, q2, sgb, dotp);
            if (ZT_HALVES == 2) asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
            if (!half && pr0 < p.P) {
                if (ZT_HALVES == 2) {
                    const float4 o = s_part[n & 1][q][lane];
                    var += o.x; q2 += o.y; sgb += o.z; dotp += o.w;
                }
                const float rstd = rsqrtf(var * inv_d + p.ln_eps);
                const float dot = fmaf(rstd, dotp, b_r);
                const float nn = fmaf(rstd * rstd, q2, fmaf(2.f * rstd, sgb, s_scal[1]));
                p.score[pr0] = nn > 0.f ? dot / sqrtf(nn) * p.inv_nvec : 0.f;
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();                                             // neither CTA may retire while its partner can still touch its barriers / shared memory
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc_pair(tmem, 512);
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
    const size_t smem = (size_t)ZE_TILE * (2 * m->D + 3 * (m->D / 2)) * sizeof(float);
    zsl_entity_kernel<<<(unsigned)((n_ent + ZE_TILE - 1) / ZE_TILE), 128, smem, st>>>(*m, ent_symbol, conn, deg, n_ent, max_nb, A, B);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

int zsl_rank(mre_ctx *ctx, const mre_zsl_model *m, const float *A, const float *B, int64_t n_ent, const int64_t *q_head,
             const int64_t *q_rel, const int64_t *cand_ptr, const int64_t *cand_idx, int64_t T, int64_t P, const float *rel_vecs,
             int64_t n_rel, int32_t n_vec, float *scores, int32_t *counts, cudaStream_t st) {
    MRE_TRY(check_model(m));
    MRE_CHECK_ARG(T >= 0 && P >= 0 && n_rel >= 0 && n_vec > 0 && n_ent >= 0, "bad size");
    if (T == 0) return MRE_OK;
    MRE_CHECK_ARG(A && B && q_head && q_rel && cand_ptr && counts && rel_vecs, "NULL argument");
    MRE_CHECK_ARG(P == 0 || cand_idx, "cand_idx is NULL");
    MRE_CHECK_ARG(P == 0 || n_ent > 0, "pairs given but the entity table is empty");
    MRE_CHECK_ARG(P < (1LL << 31), "too many (head, candidate) pairs for one call");
    MRE_CHECK_ARG(n_ent < (1LL << 31), "entity ids must fit 31 bits");
    const int D = (int)m->D, K = 2 * D, NP = (D + 16) / 16 * 16;   // >= D + 1: one spare output row for the column sums
    // mre_ctx_option "zsl_fp32": the CUDA-core FP32 tile GEMMs (also taken by wide models: three operand stages of wider tiles
    // do not fit the SM's shared memory)
    const bool fp32_path = ctx->opt_zsl_fp32 != 0 || NP > 224;
    const int64_t P1 = (std::max<int64_t>(P, 1) + 3) / 4 * 4;                // keeps the float4-read relation sums 16-byte aligned
    // scratch: pair -> triple map, scores (when the caller does not want them), relation-vector norms / normalised sums
    MRE_TRY(ctx->misc2.reserve((size_t)P1 * (sizeof(int32_t) + sizeof(float)) + (size_t)n_rel * std::max(n_vec, D + 1) * sizeof(float) + 64));
    int32_t *pair_triple = ctx->misc2.as<int32_t>();
    float *sc = scores ? scores : reinterpret_cast<float *>(pair_triple + P1);
    float *rel_aux = reinterpret_cast<float *>(pair_triple + P1) + P1;
    if (P > 0) zsl_pair_triple_kernel<<<(unsigned)std::min<int64_t>((P + 255) / 256, 148 * 16), 256, 0, st>>>(cand_ptr, T, P, pair_triple);
    ctx->launches += 1;
    if (fp32_path) {
        const int64_t chunk = std::min<int64_t>(P1, 1 << 18);
        MRE_TRY(ctx->ent_aux.reserve((size_t)chunk * K * sizeof(float)));
        float *hid = ctx->ent_aux.as<float>();
        if (n_rel > 0) zsl_relnorm_kernel<<<(unsigned)((n_rel * n_vec + 127) / 128), 128, 0, st>>>(rel_vecs, n_rel * n_vec, D, rel_aux);
        ctx->launches += 1;
        MRE_TRY(ctx->time_begin(st));
        for (int64_t p0 = 0; p0 < P; p0 += chunk) {
            const int64_t n = std::min(chunk, P - p0);
            dim3 g1((unsigned)((n + 127) / 128), (unsigned)((K + 127) / 128));
            zsl_layer1_kernel<<<g1, 256, 0, st>>>(*m, A, B, q_head, cand_idx, pair_triple, p0, n, hid, 1, 1);
            zsl_layer2_kernel<<<(unsigned)((n + 63) / 64), 256, 0, st>>>(*m, A, B, q_head, q_rel, cand_idx, pair_triple, hid, rel_vecs, rel_aux,
                                                                          n_vec, p0, n, sc);
            ctx->launches += 2;
        }
        MRE_TRY(ctx->time_end(st));
    } else if (P > 0) {
        // per-entity hidden halves A1 = W1 A + b1, B1 = W1 B, and the TF32 hi / lo split of W2 (rows zero-padded to NP)
        const int64_t n4 = (n_ent + 3) / 4 * 4;
        MRE_TRY(ctx->ent_aux.reserve(((size_t)2 * n_ent * K + (size_t)2 * NP * K + 2 * n4) * sizeof(float) + 256));
        float *A1 = ctx->ent_aux.as<float>(), *B1 = A1 + n_ent * K, *w_hi = B1 + n_ent * K, *w_lo = w_hi + (size_t)NP * K;
        float *sA = w_lo + (size_t)NP * K, *sB = sA + n4;
        dim3 g1((unsigned)((n_ent + 127) / 128), (unsigned)((K + 127) / 128));
        zsl_layer1_kernel<<<g1, 256, 0, st>>>(*m, A, nullptr, nullptr, nullptr, nullptr, 0, n_ent, A1, 0, 1);
        zsl_layer1_kernel<<<g1, 256, 0, st>>>(*m, B, nullptr, nullptr, nullptr, nullptr, 0, n_ent, B1, 0, 0);
        zsl_split_w2_kernel<<<(NP * K + 255) / 256, 256, 0, st>>>(m->proj2_w, D, K, NP, w_hi, w_lo);
        zsl_rowsum_kernel<<<(unsigned)((n_ent + 7) / 8), 256, 0, st>>>(A, n_ent, D, sA);
        zsl_rowsum_kernel<<<(unsigned)((n_ent + 7) / 8), 256, 0, st>>>(B, n_ent, D, sB);
        if (n_rel > 0) zsl_relsum_kernel<<<(unsigned)n_rel, 128, 0, st>>>(rel_vecs, n_vec, D, m->ln_b, rel_aux, rel_aux + n_rel * D);
        ctx->launches += 6;
        CUtensorMap tm_hi, tm_lo;
        MRE_TRY(make_tmap_f32_2d(&tm_hi, w_hi, NP, K, K, NP / 2, ZT_BK));   // a CTA of the pair loads half of the rows
        MRE_TRY(make_tmap_f32_2d(&tm_lo, w_lo, NP, K, K, NP / 2, ZT_BK));
        ZslTcParams tp;
        tp.A1 = A1; tp.B1 = B1; tp.A = A; tp.B = B; tp.sA = sA; tp.sB = sB;
        tp.q_head = q_head; tp.q_rel = q_rel; tp.cand = cand_idx; tp.pair_triple = pair_triple;
        tp.rsum = rel_aux; tp.bR = rel_aux + n_rel * D; tp.b2 = m->proj2_b; tp.ln_g = m->ln_g; tp.ln_b = m->ln_b;
        tp.ln_eps = m->ln_eps; tp.inv_nvec = 1.f / (float)n_vec;
        tp.D = D; tp.K = K; tp.NP = NP; tp.nkb = K / ZT_BK;
        tp.cslot = ctx->zsl_const_slot;
#if MRE_ZT_CONST_VEC
        {   // stream-ordered: the previous launch of this context has read its slot before these copies run
            const size_t row = (size_t)ZT_CONST_D * sizeof(float), base_off = (size_t)tp.cslot * 3 * row;
            MRE_CUDA(cudaMemcpyToSymbolAsync(c_zsl_vec, m->proj2_b, (size_t)D * sizeof(float), base_off, cudaMemcpyDeviceToDevice, st));
            MRE_CUDA(cudaMemcpyToSymbolAsync(c_zsl_vec, m->ln_g, (size_t)D * sizeof(float), base_off + row, cudaMemcpyDeviceToDevice, st));
            MRE_CUDA(cudaMemcpyToSymbolAsync(c_zsl_vec, m->ln_b, (size_t)D * sizeof(float), base_off + 2 * row, cudaMemcpyDeviceToDevice, st));
        }
#endif
        tp.idesc = umma_idesc_tf32(256, NP);
        tp.P = P; tp.tiles = (P + ZT_ROWS - 1) / ZT_ROWS;
        tp.score = sc;
        const size_t smem = (size_t)ZT_STAGES * (2 * ZT_A_BYTES + 2 * (size_t)(NP / 2) * ZT_BK * 4) + 1024;
        MRE_CUDA(cudaFuncSetAttribute(zsl_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2u * (unsigned)std::max<int64_t>(1, std::min<int64_t>(tp.tiles, sms / 2)));
        cfg.blockDim = dim3(ZT_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MRE_TRY(ctx->time_begin(st));
        MRE_CUDA(cudaLaunchKernelEx(&cfg, zsl_tc_kernel, tm_hi, tm_lo, tp));
        ctx->launches += 1;
        MRE_TRY(ctx->time_end(st));
    }
    zsl_count_kernel<<<(unsigned)((T + 7) / 8), 256, 0, st>>>(sc, cand_ptr, T, counts);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

}  // namespace mre
