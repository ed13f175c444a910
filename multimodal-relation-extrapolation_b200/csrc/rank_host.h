// Host-side helpers shared by the rank kernels' launchers.
#pragma once
#include "common.h"
#include "rank_common.cuh"

namespace mre {
// validates the job's group / filter arguments, uploads the group descriptors and fills the common RankParams fields
int fill_rank_params(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, int tile_q, int tile_e, cudaStream_t st,
                     RankParams &p);
// routes every (query, known entity | truth) pair to the work item that scores it; sets p.tf_ptr / p.tf_pairs
int build_tile_filter(mre_ctx *ctx, const mre_rank_job *job, RankParams &p, int tile_q, int tile_e, cudaStream_t st);
}  // namespace mre
