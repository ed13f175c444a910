#include "common.h"
namespace mre {
int rank_bilinear(mre_ctx *, const mre_index *, const mre_rank_job *, cudaStream_t) { set_error("bilinear not built yet"); return MRE_ERR_UNSUPPORTED; }
int predict_bilinear(mre_ctx *, const mre_rank_job *, int64_t, float *, cudaStream_t) { set_error("bilinear not built yet"); return MRE_ERR_UNSUPPORTED; }
int transe_margin_step(mre_ctx *, const float *, const float *, int64_t, int64_t, int64_t, const int64_t *, const int64_t *, const int64_t *, int64_t, int64_t, float, int32_t, int32_t, float *, float *, float *, float *, cudaStream_t) { set_error("train step not built yet"); return MRE_ERR_UNSUPPORTED; }
int sgd_update(mre_ctx *, float *, float *, int64_t, float, cudaStream_t) { set_error("not built yet"); return MRE_ERR_UNSUPPORTED; }
int probe_tf32_peak(mre_ctx *, double *) { set_error("not built yet"); return MRE_ERR_UNSUPPORTED; }
}
