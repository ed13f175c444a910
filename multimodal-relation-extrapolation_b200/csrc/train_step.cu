// Fused TransE margin-loss training step (forward + backward) and the SGD update.
//
// Reference path replaced (paths relative to /root/reference):
//   OpenKE/openke/module/model/TransE.py:46-74             gather rows, F.normalize, ||h + r - t||_p
//   OpenKE/openke/module/strategy/NegativeSampling.py:13-32  p = score[:B], n[b,k] = score[B + k*B + b]
//   OpenKE/openke/module/loss/MarginLoss.py:24-28          mean(max(p - n, -margin)) + margin
//   OpenKE/openke/config/Trainer.py:43-54                  loss.backward(); optimizer.step() (SGD)
//   module/NegativeSampling.py:142-157, module/loss.py:20-24  the paper's copy of the same scorer and loss
// The reference runs ~25 eager kernels per step and materialises three [n, D] gathered copies plus their
// normalised versions, the score vector, and autograd's saved tensors.  Here a step is a forward kernel (one warp per
// triple: row norms by warp shuffle, score), two tiny kernels over the n scores (loss, dLoss/dscore) and a backward
// kernel (one warp per triple: gradient through the norm and the normalisation, scattered with atomics into the dense
// gradient tables).  Nothing of size [n, D] is ever written.  HBM/L2-bound: per triple 3 rows read twice and 3 rows
// of atomics.
#include <algorithm>

#include "common.h"

namespace mre {

constexpr float NORM_EPS = 1e-12f;  // F.normalize's eps (TransE.py:48-50)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

struct RowNorms {
    float h, r, t;  // max(||x||_2, eps), or 1 when normalisation is off
};

__device__ __forceinline__ RowNorms row_norms(const float *__restrict__ vh, const float *__restrict__ vr,
                                              const float *__restrict__ vt, int D, int lane, int normalize) {
    RowNorms n{1.f, 1.f, 1.f};
    if (!normalize) return n;
    float sh = 0.f, sr = 0.f, st = 0.f;
    for (int d = lane; d < D; d += 32) {
        float a = vh[d], b = vr[d], c = vt[d];
        sh = fmaf(a, a, sh); sr = fmaf(b, b, sr); st = fmaf(c, c, st);
    }
    n.h = fmaxf(sqrtf(warp_sum(sh)), NORM_EPS);
    n.r = fmaxf(sqrtf(warp_sum(sr)), NORM_EPS);
    n.t = fmaxf(sqrtf(warp_sum(st)), NORM_EPS);
    return n;
}

// forward: score[i] = || h^ + r^ - t^ ||_p, one warp per triple
template <int P>
__global__ void __launch_bounds__(256) transe_fwd_kernel(const float *__restrict__ ent, const float *__restrict__ rel, int D,
                                                         const int64_t *__restrict__ bh, const int64_t *__restrict__ bt,
                                                         const int64_t *__restrict__ br, int64_t n, int normalize,
                                                         float *__restrict__ score) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
        const float *vh = ent + bh[i] * D, *vt = ent + bt[i] * D, *vr = rel + br[i] * D;
        const RowNorms nr = row_norms(vh, vr, vt, D, lane, normalize);
        float acc = 0.f;
        for (int d = lane; d < D; d += 32) {
            float u = (vh[d] / nr.h + vr[d] / nr.r) - vt[d] / nr.t;
            acc = P == 1 ? acc + fabsf(u) : fmaf(u, u, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) score[i] = P == 1 ? acc : sqrtf(acc);
    }
}

// loss = mean_{b,k} max(p_b - n_bk, -m) + m, float64 accumulation; one thread per (b, k)
__global__ void __launch_bounds__(256) margin_loss_kernel(const float *__restrict__ score, int64_t B, int64_t neg, float margin,
                                                          double *__restrict__ acc) {
    __shared__ double part[8];
    double s = 0.0;
    const int64_t total = B * neg;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i % B;
        const float v = score[b] - score[B + i];
        s += (double)(v > -margin ? v : -margin);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += part[w];
        atomicAdd(acc, t);
    }
}

__global__ void finish_loss_kernel(const double *acc, int64_t B, int64_t neg, float margin, float *loss_out) {
    loss_out[0] = (float)(acc[0] / (double)(B * neg) + (double)margin);
}

// dLoss/dscore for the margin loss on the strategy's layout: positives collect one term per active negative of their
// row, negatives get -1/(B*neg) when active (MarginLoss.py:24-28 through strategy/NegativeSampling.py:13-21)
__global__ void __launch_bounds__(256) margin_grad_kernel(const float *__restrict__ score, int64_t B, int64_t neg, float margin,
                                                          float *__restrict__ dscore) {
    const int64_t n = B * (1 + neg);
    const float inv = 1.0f / (float)(B * neg);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float c;
        if (i < B) {
            const float p = score[i];
            int active = 0;
            for (int64_t k = 0; k < neg; k++) active += (p - score[B + k * B + i] > -margin) ? 1 : 0;
            c = (float)active * inv;
        } else {
            const int64_t b = (i - B) % B;
            c = (score[b] - score[i] > -margin) ? -inv : 0.f;
        }
        dscore[i] = c;
    }
}

// backward of the TransE score through the norm and the normalisation: one warp per triple, c = dLoss/dscore_i
template <int P>
__global__ void __launch_bounds__(256) transe_bwd_kernel(const float *__restrict__ ent, const float *__restrict__ rel, int D,
                                                         const int64_t *__restrict__ bh, const int64_t *__restrict__ bt,
                                                         const int64_t *__restrict__ br, int64_t n, int normalize,
                                                         const float *__restrict__ score, const float *__restrict__ dscore,
                                                         float *__restrict__ grad_ent, float *__restrict__ grad_rel) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
        const float c = dscore[i];
        if (c == 0.f) continue;
        const int64_t ih = bh[i], it = bt[i], ir = br[i];
        const float *vh = ent + ih * D, *vt = ent + it * D, *vr = rel + ir * D;
        const RowNorms nr = row_norms(vh, vr, vt, D, lane, normalize);
        // g = c * d||u||_p/du ; dots of g with the normalised rows (for the projection in normalize's backward)
        float un = 1.f;
        if (P == 2) {
            const float s = score[i];
            un = s > 0.f ? 1.f / s : 0.f;
        }
        float dh = 0.f, dr = 0.f, dt = 0.f;
        if (normalize) {
            for (int d = lane; d < D; d += 32) {
                const float xh = vh[d] / nr.h, xr = vr[d] / nr.r, xt = vt[d] / nr.t;
                const float u = (xh + xr) - xt;
                const float g = P == 1 ? (u > 0.f ? c : (u < 0.f ? -c : 0.f)) : c * u * un;
                dh = fmaf(xh, g, dh); dr = fmaf(xr, g, dr); dt = fmaf(xt, g, dt);
            }
            dh = warp_sum(dh); dr = warp_sum(dr); dt = warp_sum(dt);
        }
        float *gh = grad_ent + ih * D, *gt = grad_ent + it * D, *gr = grad_rel + ir * D;
        for (int d = lane; d < D; d += 32) {
            const float xh = vh[d] / nr.h, xr = vr[d] / nr.r, xt = vt[d] / nr.t;
            const float u = (xh + xr) - xt;
            const float g = P == 1 ? (u > 0.f ? c : (u < 0.f ? -c : 0.f)) : c * u * un;
            float g_h = g, g_r = g, g_t = -g;
            if (normalize) {
                // d(x / max(||x||, eps))/dx applied to the incoming gradient; below eps the map is x / eps
                g_h = nr.h > NORM_EPS ? (g - xh * dh) / nr.h : g / NORM_EPS;
                g_r = nr.r > NORM_EPS ? (g - xr * dr) / nr.r : g / NORM_EPS;
                g_t = nr.t > NORM_EPS ? (-g + xt * dt) / nr.t : -g / NORM_EPS;
            }
            atomicAdd(gh + d, g_h);
            atomicAdd(gr + d, g_r);
            atomicAdd(gt + d, g_t);
        }
    }
}

// forward scores of the similarity models for explicit triples (Model.forward, DistMult.py:46-57, ComplEx.py:29-40):
// one warp per triple; the value is the raw similarity (predict negates it)
__global__ void __launch_bounds__(256) bilinear_fwd_kernel(int scorer, const float *__restrict__ ent, const float *__restrict__ ent_im,
                                                           const float *__restrict__ rel, const float *__restrict__ rel_im, int D,
                                                           const int64_t *__restrict__ bh, const int64_t *__restrict__ bt,
                                                           const int64_t *__restrict__ br, int64_t n, float *__restrict__ score) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
        const int64_t h = bh[i] * D, t = bt[i] * D, r = br[i] * D;
        float acc = 0.f;
        for (int d = lane; d < D; d += 32) {
            if (scorer == MRE_DISTMULT) {
                acc = acc + (ent[h + d] * rel[r + d]) * ent[t + d];
            } else {
                const float hr = ent[h + d], hi = ent_im[h + d], tr = ent[t + d], ti = ent_im[t + d], rr = rel[r + d], ri = rel_im[r + d];
                acc = acc + (hr * tr * rr + hi * ti * rr + hr * ti * ri - hi * tr * ri);
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) score[i] = acc;
    }
}

__global__ void sgd_kernel(float *__restrict__ w, float *__restrict__ g, int64_t n, float lr) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        w[i] = w[i] - lr * g[i];
        g[i] = 0.f;
    }
}

static int launch_grid(mre_ctx *ctx, int64_t n) { return (int)std::min<int64_t>((n + 7) / 8, (int64_t)ctx->sm_count * 16); }

int score_triples(mre_ctx *ctx, int scorer, const float *ent, const float *ent_im, const float *rel, const float *rel_im, int64_t D,
                  const int64_t *h, const int64_t *t, const int64_t *r, int64_t n, int32_t p_norm, int32_t normalize, float *score,
                  cudaStream_t st) {
    MRE_CHECK_ARG(ent && rel && h && t && r && score, "NULL argument");
    MRE_CHECK_ARG(D > 0 && D < (1 << 30) && n >= 0, "bad shape");
    if (n == 0) return MRE_OK;
    const int grid = launch_grid(ctx, n);
    if (scorer == MRE_TRANSE) {
        MRE_CHECK_ARG(p_norm == 1 || p_norm == 2, "p_norm must be 1 or 2");
        if (p_norm == 1) transe_fwd_kernel<1><<<grid, 256, 0, st>>>(ent, rel, (int)D, h, t, r, n, normalize, score);
        else transe_fwd_kernel<2><<<grid, 256, 0, st>>>(ent, rel, (int)D, h, t, r, n, normalize, score);
    } else {
        MRE_CHECK_ARG(scorer == MRE_DISTMULT || (scorer == MRE_COMPLEX && ent_im && rel_im), "bad scorer / missing ComplEx tables");
        bilinear_fwd_kernel<<<grid, 256, 0, st>>>(scorer, ent, ent_im, rel, rel_im, (int)D, h, t, r, n, score);
    }
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

int transe_backward(mre_ctx *ctx, const float *ent, const float *rel, int64_t D, const int64_t *h, const int64_t *t, const int64_t *r,
                    int64_t n, int32_t p_norm, int32_t normalize, const float *score, const float *dscore, float *grad_ent,
                    float *grad_rel, cudaStream_t st) {
    MRE_CHECK_ARG(ent && rel && h && t && r && score && dscore && grad_ent && grad_rel, "NULL argument");
    MRE_CHECK_ARG(p_norm == 1 || p_norm == 2, "p_norm must be 1 or 2");
    if (n == 0) return MRE_OK;
    const int grid = launch_grid(ctx, n);
    if (p_norm == 1) transe_bwd_kernel<1><<<grid, 256, 0, st>>>(ent, rel, (int)D, h, t, r, n, normalize, score, dscore, grad_ent, grad_rel);
    else transe_bwd_kernel<2><<<grid, 256, 0, st>>>(ent, rel, (int)D, h, t, r, n, normalize, score, dscore, grad_ent, grad_rel);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

int transe_margin_step(mre_ctx *ctx, const float *ent, const float *rel, int64_t E, int64_t R, int64_t D, const int64_t *h,
                       const int64_t *t, const int64_t *r, int64_t B, int64_t neg, float margin, int32_t p_norm,
                       int32_t normalize, float *grad_ent, float *grad_rel, float *loss_out, float *scores_out,
                       cudaStream_t st) {
    MRE_CHECK_ARG(ent && rel && h && t && r && grad_ent && grad_rel && loss_out, "NULL argument");
    MRE_CHECK_ARG(E > 0 && R > 0 && D > 0 && D < (1 << 30), "bad table shape");
    MRE_CHECK_ARG(B > 0 && neg > 0, "B and neg must be positive");
    MRE_CHECK_ARG(p_norm == 1 || p_norm == 2, "p_norm must be 1 or 2");
    const int64_t n = B * (1 + neg);
    MRE_TRY(ctx->qvec.reserve((size_t)2 * n * sizeof(float)));
    float *dscore = ctx->qvec.as<float>();
    float *score = scores_out ? scores_out : dscore + n;
    MRE_TRY(ctx->misc.reserve(256));
    double *acc = ctx->misc.as<double>();
    MRE_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
    MRE_TRY(ctx->time_begin(st));
    MRE_TRY(score_triples(ctx, MRE_TRANSE, ent, nullptr, rel, nullptr, D, h, t, r, n, p_norm, normalize, score, st));
    const int lgrid = (int)std::min<int64_t>((B * neg + 255) / 256, (int64_t)ctx->sm_count * 4);
    margin_loss_kernel<<<lgrid, 256, 0, st>>>(score, B, neg, margin, acc);
    finish_loss_kernel<<<1, 1, 0, st>>>(acc, B, neg, margin, loss_out);
    margin_grad_kernel<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8), 256, 0, st>>>(score, B, neg, margin, dscore);
    ctx->launches += 3;
    MRE_TRY(transe_backward(ctx, ent, rel, D, h, t, r, n, p_norm, normalize, score, dscore, grad_ent, grad_rel, st));
    MRE_TRY(ctx->time_end(st));
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

int sgd_update(mre_ctx *ctx, float *w, float *g, int64_t n, float lr, cudaStream_t st) {
    MRE_CHECK_ARG(w && g && n >= 0, "bad argument");
    if (n == 0) return MRE_OK;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
    sgd_kernel<<<grid, 256, 0, st>>>(w, g, n, lr);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

}  // namespace mre
