#include "common.h"
namespace mre {
int rank_bilinear(mre_ctx *, const mre_index *, const mre_rank_job *, cudaStream_t) { set_error("bilinear not built yet"); return MRE_ERR_UNSUPPORTED; }
int predict_bilinear(mre_ctx *, const mre_rank_job *, int64_t, float *, cudaStream_t) { set_error("bilinear not built yet"); return MRE_ERR_UNSUPPORTED; }
int probe_tf32_peak(mre_ctx *, double *) { set_error("not built yet"); return MRE_ERR_UNSUPPORTED; }
}
