// Per-relation entity tables of the translation models that project entities before translating them:
//   OpenKE/openke/module/model/TransH.py:66-74   _transfer: e - (e . w^) w^,  w^ = F.normalize(norm_vector[r])
//   OpenKE/openke/module/model/TransD.py:92-109  _transfer: F.normalize(e + (e . e_p) r_p)     (dim_e == dim_r)
//   TransH.py:51-55 / TransD.py:77-81            _calc: F.normalize of h, t again when norm_flag
// The reference applies the projection to the E rows of every 1-vs-all query; the projected table is a function of the RELATION
// only, so it is built once per relation of the job -- out[slot * E + e, :] for the job's slot-th relation -- and the TransE
// rank kernel then streams it as that relation's candidate group (openke/module/model/Model.py: RelationProjected).
// One warp per output row: lanes across d, dot products and norms by warp shuffles.
#include <algorithm>

#include "common.h"

namespace mre {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

constexpr int PROJ_MAX_PER_LANE = 16;    // D <= 512

__global__ void __launch_bounds__(256) relation_project_kernel(int kind, const float *__restrict__ ent, const float *__restrict__ ent_aux,
                                                               const float *__restrict__ rel_aux, const int64_t *__restrict__ rels,
                                                               int64_t n_rel, int64_t E, int D, int renorm, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t rows = n_rel * E;
    for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * 8) {
        const int64_t slot = row / E, e = row - slot * E;
        const float *x = ent + e * D, *w = rel_aux + __ldg(rels + slot) * D;
        float xv[PROJ_MAX_PER_LANE], wv[PROJ_MAX_PER_LANE];
        float dot = 0.f, wn = 0.f;
#pragma unroll
        for (int k = 0; k < PROJ_MAX_PER_LANE; k++) {
            const int d = lane + 32 * k;
            xv[k] = d < D ? __ldg(x + d) : 0.f;
            wv[k] = d < D ? __ldg(w + d) : 0.f;
            wn = fmaf(wv[k], wv[k], wn);
        }
        if (kind == 0) {                                   // TransH: w^ = w / max(||w||, 1e-12);  e - (e . w^) w^
            const float nrm = fmaxf(sqrtf(warp_sum(wn)), 1e-12f);
#pragma unroll
            for (int k = 0; k < PROJ_MAX_PER_LANE; k++) {
                wv[k] = wv[k] / nrm;
                dot = fmaf(xv[k], wv[k], dot);
            }
            dot = warp_sum(dot);
#pragma unroll
            for (int k = 0; k < PROJ_MAX_PER_LANE; k++) xv[k] = xv[k] - dot * wv[k];
        } else {                                           // TransD: normalize(e + (e . e_p) r_p)
            const float *xp = ent_aux + e * D;
#pragma unroll
            for (int k = 0; k < PROJ_MAX_PER_LANE; k++) {
                const int d = lane + 32 * k;
                dot = fmaf(xv[k], d < D ? __ldg(xp + d) : 0.f, dot);
            }
            dot = warp_sum(dot);
            float ss = 0.f;
#pragma unroll
            for (int k = 0; k < PROJ_MAX_PER_LANE; k++) {
                xv[k] = xv[k] + dot * wv[k];
                ss = fmaf(xv[k], xv[k], ss);
            }
            const float nrm = fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
#pragma unroll
            for (int k = 0; k < PROJ_MAX_PER_LANE; k++) xv[k] = xv[k] / nrm;
        }
        if (renorm) {                                      // _calc's F.normalize(., 2, -1) on the projected vector
            float ss = 0.f;
#pragma unroll
            for (int k = 0; k < PROJ_MAX_PER_LANE; k++) ss = fmaf(xv[k], xv[k], ss);
            const float nrm = fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
#pragma unroll
            for (int k = 0; k < PROJ_MAX_PER_LANE; k++) xv[k] = xv[k] / nrm;
        }
        float *o = out + row * D;
#pragma unroll
        for (int k = 0; k < PROJ_MAX_PER_LANE; k++) {
            const int d = lane + 32 * k;
            if (d < D) o[d] = xv[k];
        }
    }
}

}  // namespace mre

extern "C" int mre_relation_project(mre_ctx *ctx, int32_t kind, const float *ent, const float *ent_aux, const float *rel_aux,
                                    const int64_t *rels, int64_t n_rel, int64_t E, int64_t D, int32_t renormalize, float *out,
                                    void *stream) {
    MRE_CHECK_ARG(ctx && ent && rel_aux && rels && out, "NULL argument");
    MRE_CHECK_ARG(kind == MRE_PROJECT_TRANSH || (kind == MRE_PROJECT_TRANSD && ent_aux), "kind must be MRE_PROJECT_TRANSH or MRE_PROJECT_TRANSD (with ent_aux)");
    MRE_CHECK_ARG(n_rel >= 0 && E > 0 && D > 0 && D <= 32 * mre::PROJ_MAX_PER_LANE, "bad shape (D <= %d)", 32 * mre::PROJ_MAX_PER_LANE);
    if (n_rel == 0) return MRE_OK;
    MRE_CUDA(cudaSetDevice(ctx->device));
    const int64_t rows = n_rel * E;
    const int grid = (int)std::min<int64_t>((rows + 7) / 8, (int64_t)ctx->sm_count * 32);
    mre::relation_project_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kind, ent, ent_aux, rel_aux, rels, n_rel, E, (int)D, renormalize, out);
    ctx->launches += 1;
    MRE_CUDA(cudaGetLastError());
    return MRE_OK;
}
