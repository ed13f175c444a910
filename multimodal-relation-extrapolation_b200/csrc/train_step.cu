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

// The same forward for D % 4 == 0 and D <= 128 * MAXV: a lane owns float4 groups, the three rows are fetched once with 128-bit
// loads and stay in registers through the norm and score phases (the kernel above reads them twice).
template <int P, int MAXV>
__global__ void __launch_bounds__(256) transe_fwd_vec_kernel(const float *__restrict__ ent, const float *__restrict__ rel, int D,
                                                             const int64_t *__restrict__ bh, const int64_t *__restrict__ bt,
                                                             const int64_t *__restrict__ br, int64_t n, int normalize,
                                                             float *__restrict__ score) {
    const int lane = threadIdx.x & 31;
    const int nv = D >> 2;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
        const float4 *vh = reinterpret_cast<const float4 *>(ent + bh[i] * D), *vt = reinterpret_cast<const float4 *>(ent + bt[i] * D),
                     *vr = reinterpret_cast<const float4 *>(rel + br[i] * D);
        float xh[MAXV][4], xr[MAXV][4], xt[MAXV][4];
        float sh = 0.f, sr = 0.f, st = 0.f;
#pragma unroll
        for (int k = 0; k < MAXV; k++) {
            const int v = lane + 32 * k;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 a = v < nv ? vh[v] : z, b = v < nv ? vr[v] : z, e = v < nv ? vt[v] : z;
            xh[k][0] = a.x; xh[k][1] = a.y; xh[k][2] = a.z; xh[k][3] = a.w;
            xr[k][0] = b.x; xr[k][1] = b.y; xr[k][2] = b.z; xr[k][3] = b.w;
            xt[k][0] = e.x; xt[k][1] = e.y; xt[k][2] = e.z; xt[k][3] = e.w;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                sh = fmaf(xh[k][j], xh[k][j], sh); sr = fmaf(xr[k][j], xr[k][j], sr); st = fmaf(xt[k][j], xt[k][j], st);
            }
        }
        float nh = 1.f, nr = 1.f, nt = 1.f;
        if (normalize) {
            nh = fmaxf(sqrtf(warp_sum(sh)), NORM_EPS);
            nr = fmaxf(sqrtf(warp_sum(sr)), NORM_EPS);
            nt = fmaxf(sqrtf(warp_sum(st)), NORM_EPS);
        }
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < MAXV; k++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float u = normalize ? (xh[k][j] / nh + xr[k][j] / nr) - xt[k][j] / nt : (xh[k][j] + xr[k][j]) - xt[k][j];
                acc = P == 1 ? acc + fabsf(u) : fmaf(u, u, acc);
            }
        acc = warp_sum(acc);
        if (lane == 0) score[i] = P == 1 ? acc : sqrtf(acc);
    }
}

// Negative-sampling losses on the strategy's score layout (p_b = score[b], n_bk = score[B + k*B + b],
// strategy/NegativeSampling.py:13-21), forward value and dLoss/dscore in ONE launch -- one thread per positive row b:
//   MARGIN    mean_b sum_k w_bk max(p_b - n_bk, -m) + m                     MarginLoss.py:24-28 (module/loss.py:20-24)
//   SIGMOID   -(mean_b logsig(p_b) + mean_b sum_k w_bk logsig(-n_bk)) / 2   SigmoidLoss.py:22-26
//   SOFTPLUS   (mean_b softplus(-p_b) + mean_b sum_k w_bk softplus(n_bk)) / 2   SoftplusLoss.py:22-26
// w_bk = 1/neg, or the detached self-adversarial weights softmax_k(-T n_bk) (margin, MarginLoss.py:21-22) /
// softmax_k(+T n_bk) (sigmoid, softplus; SigmoidLoss.py:19-20) when adv != 0.  The loss is accumulated in float64; the
// last block to finish writes loss_out and re-arms the accumulator, so a step needs no memset and no finishing launch.
__device__ __forceinline__ float log_sigmoid(float x) {          // nn.LogSigmoid: min(x, 0) - log1p(exp(-|x|))
    return fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float softplus20(float x) {           // nn.Softplus(beta = 1, threshold = 20)
    return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

template <int KIND>
__global__ void __launch_bounds__(256) ns_loss_kernel(const float *__restrict__ score, int64_t B, int64_t neg, float margin,
                                                       int adv, float temperature, float *__restrict__ dscore,
                                                       double *__restrict__ acc, unsigned int *__restrict__ done,
                                                       float *__restrict__ loss_out) {
    __shared__ double part[8];
    double local = 0.0;
    const float invB = 1.0f / (float)B;
    const float sgn = KIND == MRE_LOSS_MARGIN ? -temperature : temperature;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
        const float p = score[b];
        const float *nrow = score + B + b;
        float zmax = -INFINITY, zsum = 0.f;
        if (adv) {                                   // softmax over the row's negatives, two passes (max, then sum)
            for (int64_t k = 0; k < neg; k++) zmax = fmaxf(zmax, sgn * nrow[k * B]);
            for (int64_t k = 0; k < neg; k++) zsum += expf(sgn * nrow[k * B] - zmax);
        }
        const float wu = 1.0f / (float)neg;
        float row = 0.f, dp = 0.f;
        for (int64_t k = 0; k < neg; k++) {
            const float nk = nrow[k * B];
            const float w = adv ? expf(sgn * nk - zmax) / zsum : wu;
            float term, dn;
            if (KIND == MRE_LOSS_MARGIN) {
                const float v = p - nk;
                const bool active = v > -margin;
                term = active ? v : -margin;
                dn = active ? -w * invB : 0.f;
                dp += active ? w * invB : 0.f;
            } else if (KIND == MRE_LOSS_SIGMOID) {
                term = -0.5f * log_sigmoid(-nk);
                dn = 0.5f * w * invB * sigmoidf(nk);
            } else {
                term = 0.5f * softplus20(nk);
                dn = 0.5f * w * invB * (nk > 20.f ? 1.f : sigmoidf(nk));
            }
            row += w * term;
            if (dscore) dscore[B + k * B + b] = dn;
        }
        if (KIND == MRE_LOSS_SIGMOID) {
            row += -0.5f * log_sigmoid(p);
            dp = -0.5f * invB * sigmoidf(-p);
        } else if (KIND == MRE_LOSS_SOFTPLUS) {
            row += 0.5f * softplus20(-p);
            dp = -0.5f * invB * (-p > 20.f ? 1.f : sigmoidf(-p));
        }
        if (dscore) dscore[b] = dp;
        local += (double)row;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) local += __shfl_xor_sync(0xffffffffu, local, m);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += part[w];
        atomicAdd(acc, t);
        __threadfence();
        if (atomicAdd(done, 1u) == gridDim.x - 1) {          // last block: publish the loss, re-arm for the next call
            const double total = atomicAdd(acc, 0.0);
            loss_out[0] = (float)(total / (double)B + (KIND == MRE_LOSS_MARGIN ? (double)margin : 0.0));
            *acc = 0.0;
            *done = 0u;
        }
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

// The same backward for D % 4 == 0 and D <= 128 * MAXV, the shape every config of the reference trains at (D = 200): a lane
// owns float4 groups, the three rows are fetched ONCE with 128-bit loads and stay in registers through the norm, dot and
// gradient phases (the kernel above re-reads and re-normalises them in each: nine divisions per element, three of them here),
// and the gradients leave as 128-bit vector reductions (RED.ADD.F32x4: a quarter of the atomic instructions).  The kernel
// above was issue-bound (72 % of issue slots, 1 290 warp instructions per triple); this one executes about a third of them.
template <int P, int MAXV>
__global__ void __launch_bounds__(256) transe_bwd_vec_kernel(const float *__restrict__ ent, const float *__restrict__ rel, int D,
                                                             const int64_t *__restrict__ bh, const int64_t *__restrict__ bt,
                                                             const int64_t *__restrict__ br, int64_t n, int normalize,
                                                             const float *__restrict__ score, const float *__restrict__ dscore,
                                                             float *__restrict__ grad_ent, float *__restrict__ grad_rel) {
    const int lane = threadIdx.x & 31;
    const int nv = D >> 2;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
        const float c = dscore[i];
        if (c == 0.f) continue;
        const int64_t ih = bh[i], it = bt[i], ir = br[i];
        const float4 *vh = reinterpret_cast<const float4 *>(ent + ih * D), *vt = reinterpret_cast<const float4 *>(ent + it * D),
                     *vr = reinterpret_cast<const float4 *>(rel + ir * D);
        float xh[MAXV][4], xr[MAXV][4], xt[MAXV][4];
        float sh = 0.f, sr = 0.f, st = 0.f;
#pragma unroll
        for (int k = 0; k < MAXV; k++) {
            const int v = lane + 32 * k;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 a = v < nv ? vh[v] : z, b = v < nv ? vr[v] : z, e = v < nv ? vt[v] : z;
            xh[k][0] = a.x; xh[k][1] = a.y; xh[k][2] = a.z; xh[k][3] = a.w;
            xr[k][0] = b.x; xr[k][1] = b.y; xr[k][2] = b.z; xr[k][3] = b.w;
            xt[k][0] = e.x; xt[k][1] = e.y; xt[k][2] = e.z; xt[k][3] = e.w;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                sh = fmaf(xh[k][j], xh[k][j], sh); sr = fmaf(xr[k][j], xr[k][j], sr); st = fmaf(xt[k][j], xt[k][j], st);
            }
        }
        float nh = 1.f, nr = 1.f, nt = 1.f;
        if (normalize) {
            nh = fmaxf(sqrtf(warp_sum(sh)), NORM_EPS);
            nr = fmaxf(sqrtf(warp_sum(sr)), NORM_EPS);
            nt = fmaxf(sqrtf(warp_sum(st)), NORM_EPS);
#pragma unroll
            for (int k = 0; k < MAXV; k++)
#pragma unroll
                for (int j = 0; j < 4; j++) { xh[k][j] = xh[k][j] / nh; xr[k][j] = xr[k][j] / nr; xt[k][j] = xt[k][j] / nt; }
        }
        float un = 1.f;
        if (P == 2) {
            const float s = score[i];
            un = s > 0.f ? 1.f / s : 0.f;
        }
        float g[MAXV][4];
        float dh = 0.f, dr = 0.f, dt = 0.f;
#pragma unroll
        for (int k = 0; k < MAXV; k++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float u = (xh[k][j] + xr[k][j]) - xt[k][j];
                g[k][j] = P == 1 ? (u > 0.f ? c : (u < 0.f ? -c : 0.f)) : c * u * un;
                dh = fmaf(xh[k][j], g[k][j], dh); dr = fmaf(xr[k][j], g[k][j], dr); dt = fmaf(xt[k][j], g[k][j], dt);
            }
        float ah = 1.f, ar = 1.f, at = 1.f;                 // d(x / max(||x||, eps))/dx: (g - x^ (x^ . g)) / ||x||; below eps the map is x / eps
        if (normalize) {
            dh = warp_sum(dh); dr = warp_sum(dr); dt = warp_sum(dt);
            if (!(nh > NORM_EPS)) dh = 0.f;
            if (!(nr > NORM_EPS)) dr = 0.f;
            if (!(nt > NORM_EPS)) dt = 0.f;
            ah = 1.f / nh; ar = 1.f / nr; at = 1.f / nt;
        } else {
            dh = dr = dt = 0.f;
        }
        float4 *gh = reinterpret_cast<float4 *>(grad_ent + ih * D), *gt = reinterpret_cast<float4 *>(grad_ent + it * D),
               *gr = reinterpret_cast<float4 *>(grad_rel + ir * D);
#pragma unroll
        for (int k = 0; k < MAXV; k++) {
            const int v = lane + 32 * k;
            if (v < nv) {
                float oh[4], orr[4], ot[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    oh[j] = (g[k][j] - xh[k][j] * dh) * ah;
                    orr[j] = (g[k][j] - xr[k][j] * dr) * ar;
                    ot[j] = (-g[k][j] + xt[k][j] * dt) * at;
                }
                atomicAdd(gh + v, make_float4(oh[0], oh[1], oh[2], oh[3]));
                atomicAdd(gr + v, make_float4(orr[0], orr[1], orr[2], orr[3]));
                atomicAdd(gt + v, make_float4(ot[0], ot[1], ot[2], ot[3]));
            }
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

// backward of the similarity models (autograd through DistMult.py:34-44 / ComplEx.py:20-27): one warp per triple,
// c = dLoss/dscore_i, scattered with float atomics into the dense gradient tables
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(int scorer, const float *__restrict__ ent, const float *__restrict__ ent_im,
                                                           const float *__restrict__ rel, const float *__restrict__ rel_im, int D,
                                                           const int64_t *__restrict__ bh, const int64_t *__restrict__ bt,
                                                           const int64_t *__restrict__ br, int64_t n, const float *__restrict__ dscore,
                                                           float *__restrict__ g_ent, float *__restrict__ g_ent_im,
                                                           float *__restrict__ g_rel, float *__restrict__ g_rel_im) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
        const float c = dscore[i];
        if (c == 0.f) continue;
        const int64_t h = bh[i] * D, t = bt[i] * D, r = br[i] * D;
        for (int d = lane; d < D; d += 32) {
            if (scorer == MRE_DISTMULT) {
                const float vh = ent[h + d], vt = ent[t + d], vr = rel[r + d];
                atomicAdd(g_ent + h + d, c * (vr * vt));
                atomicAdd(g_ent + t + d, c * (vh * vr));
                atomicAdd(g_rel + r + d, c * (vh * vt));
            } else {
                const float hr = ent[h + d], hi = ent_im[h + d], tr = ent[t + d], ti = ent_im[t + d], rr = rel[r + d], ri = rel_im[r + d];
                atomicAdd(g_ent + h + d, c * (tr * rr + ti * ri));
                atomicAdd(g_ent_im + h + d, c * (ti * rr - tr * ri));
                atomicAdd(g_ent + t + d, c * (hr * rr - hi * ri));
                atomicAdd(g_ent_im + t + d, c * (hi * rr + hr * ri));
                atomicAdd(g_rel + r + d, c * (hr * tr + hi * ti));
                atomicAdd(g_rel_im + r + d, c * (hr * ti - hi * tr));
            }
        }
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
        const bool vec = D % 4 == 0 && D <= 512 && ((uintptr_t)ent % 16 == 0) && ((uintptr_t)rel % 16 == 0);
#define MRE_FWD(KERN) KERN<<<grid, 256, 0, st>>>(ent, rel, (int)D, h, t, r, n, normalize, score)
        if (vec && D <= 256) { if (p_norm == 1) MRE_FWD((transe_fwd_vec_kernel<1, 2>)); else MRE_FWD((transe_fwd_vec_kernel<2, 2>)); }
        else if (vec) { if (p_norm == 1) MRE_FWD((transe_fwd_vec_kernel<1, 4>)); else MRE_FWD((transe_fwd_vec_kernel<2, 4>)); }
        else if (p_norm == 1) MRE_FWD(transe_fwd_kernel<1>);
        else MRE_FWD(transe_fwd_kernel<2>);
#undef MRE_FWD
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
    // rows and gradient rows 16-byte aligned (D % 4 == 0 on cudaMalloc'ed tables) and short enough to live in registers: the
    // vector form; anything else: the scalar form
    const bool vec = D % 4 == 0 && D <= 512 && ((uintptr_t)ent % 16 == 0) && ((uintptr_t)rel % 16 == 0) && ((uintptr_t)grad_ent % 16 == 0) &&
                     ((uintptr_t)grad_rel % 16 == 0);
#define MRE_BWD(P, KERN) KERN<<<grid, 256, 0, st>>>(ent, rel, (int)D, h, t, r, n, normalize, score, dscore, grad_ent, grad_rel)
    if (vec && D <= 256) { if (p_norm == 1) MRE_BWD(1, (transe_bwd_vec_kernel<1, 2>)); else MRE_BWD(2, (transe_bwd_vec_kernel<2, 2>)); }
    else if (vec) { if (p_norm == 1) MRE_BWD(1, (transe_bwd_vec_kernel<1, 4>)); else MRE_BWD(2, (transe_bwd_vec_kernel<2, 4>)); }
    else if (p_norm == 1) MRE_BWD(1, transe_bwd_kernel<1>);
    else MRE_BWD(2, transe_bwd_kernel<2>);
#undef MRE_BWD
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

int ns_loss(mre_ctx *ctx, int32_t kind, const float *score, int64_t B, int64_t neg, float margin, int32_t adv, float temperature,
            float *loss_out, float *dscore, cudaStream_t st) {
    MRE_CHECK_ARG(score && loss_out, "NULL argument");
    MRE_CHECK_ARG(B > 0 && neg > 0, "B and neg must be positive");
    MRE_CHECK_ARG(kind >= MRE_LOSS_MARGIN && kind <= MRE_LOSS_SOFTPLUS, "unknown loss kind %d", kind);
    if (!ctx->loss_acc.p) {                               // accumulator + finished-block counter, re-armed by the kernel itself
        MRE_TRY(ctx->loss_acc.reserve(64));
        MRE_CUDA(cudaMemset(ctx->loss_acc.p, 0, 64));
    }
    double *acc = ctx->loss_acc.as<double>();
    unsigned int *done = (unsigned int *)(acc + 1);
    const int grid = (int)std::min<int64_t>((B + 255) / 256, (int64_t)ctx->sm_count * 4);
    if (kind == MRE_LOSS_MARGIN) ns_loss_kernel<MRE_LOSS_MARGIN><<<grid, 256, 0, st>>>(score, B, neg, margin, adv, temperature, dscore, acc, done, loss_out);
    else if (kind == MRE_LOSS_SIGMOID) ns_loss_kernel<MRE_LOSS_SIGMOID><<<grid, 256, 0, st>>>(score, B, neg, margin, adv, temperature, dscore, acc, done, loss_out);
    else ns_loss_kernel<MRE_LOSS_SOFTPLUS><<<grid, 256, 0, st>>>(score, B, neg, margin, adv, temperature, dscore, acc, done, loss_out);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

int bilinear_backward(mre_ctx *ctx, int scorer, const float *ent, const float *ent_im, const float *rel, const float *rel_im, int64_t D,
                      const int64_t *h, const int64_t *t, const int64_t *r, int64_t n, const float *dscore, float *g_ent,
                      float *g_ent_im, float *g_rel, float *g_rel_im, cudaStream_t st) {
    MRE_CHECK_ARG(ent && rel && h && t && r && dscore && g_ent && g_rel, "NULL argument");
    MRE_CHECK_ARG(scorer == MRE_DISTMULT || (scorer == MRE_COMPLEX && ent_im && rel_im && g_ent_im && g_rel_im),
                  "bad scorer / missing ComplEx tables");
    MRE_CHECK_ARG(D > 0 && D < (1 << 30) && n >= 0, "bad shape");
    if (n == 0) return MRE_OK;
    bilinear_bwd_kernel<<<launch_grid(ctx, n), 256, 0, st>>>(scorer, ent, ent_im, rel, rel_im, (int)D, h, t, r, n, dscore, g_ent, g_ent_im,
                                                             g_rel, g_rel_im);
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
    MRE_TRY(ctx->time_begin(st));
    MRE_TRY(score_triples(ctx, MRE_TRANSE, ent, nullptr, rel, nullptr, D, h, t, r, n, p_norm, normalize, score, st));
    MRE_TRY(ns_loss(ctx, MRE_LOSS_MARGIN, score, B, neg, margin, 0, 0.f, loss_out, dscore, st));
    MRE_TRY(transe_backward(ctx, ent, rel, D, h, t, r, n, p_norm, normalize, score, dscore, grad_ent, grad_rel, st));
    MRE_TRY(ctx->time_end(st));
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}

int ns_train_step(mre_ctx *ctx, int32_t scorer, const float *ent, const float *ent_im, const float *rel, const float *rel_im,
                  int64_t D, const int64_t *h, const int64_t *t, const int64_t *r, int64_t B, int64_t neg, int32_t loss_kind,
                  float margin, int32_t adv, float temperature, int32_t p_norm, int32_t normalize, float *g_ent, float *g_ent_im,
                  float *g_rel, float *g_rel_im, float *loss_out, float *scores_out, cudaStream_t st) {
    MRE_CHECK_ARG(ent && rel && h && t && r && g_ent && g_rel && loss_out, "NULL argument");
    MRE_CHECK_ARG(B > 0 && neg > 0 && D > 0 && D < (1 << 30), "bad shape");
    const int64_t n = B * (1 + neg);
    MRE_TRY(ctx->qvec.reserve((size_t)2 * n * sizeof(float)));
    float *dscore = ctx->qvec.as<float>();
    float *score = scores_out ? scores_out : dscore + n;
    MRE_TRY(ctx->time_begin(st));
    MRE_TRY(score_triples(ctx, scorer, ent, ent_im, rel, rel_im, D, h, t, r, n, p_norm, normalize, score, st));
    MRE_TRY(ns_loss(ctx, loss_kind, score, B, neg, margin, adv, temperature, loss_out, dscore, st));
    if (scorer == MRE_TRANSE)
        MRE_TRY(transe_backward(ctx, ent, rel, D, h, t, r, n, p_norm, normalize, score, dscore, g_ent, g_rel, st));
    else
        MRE_TRY(bilinear_backward(ctx, scorer, ent, ent_im, rel, rel_im, D, h, t, r, n, dscore, g_ent, g_ent_im, g_rel, g_rel_im, st));
    MRE_TRY(ctx->time_end(st));
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
