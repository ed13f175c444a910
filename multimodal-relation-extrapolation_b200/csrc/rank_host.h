// Host-side helpers shared by the rank kernels' launchers.
#pragma once
#include <vector>

#include "common.h"
#include "rank_common.cuh"

namespace mre {
// validates the job's group / filter arguments, uploads the group descriptors and fills the common RankParams fields
int fill_rank_params(mre_ctx *ctx, const mre_index *ix, const mre_rank_job *job, int tile_q, int tile_e, cudaStream_t st,
                     RankParams &p, std::vector<GroupDesc> *groups_out = nullptr);
// scratch of the shared-run known-true pass (transe_rank.cu)
int known_runs_scratch(mre_ctx *ctx, const RankParams &p, KnownRuns &kr);
// the flattened known-true correction serves jobs whose list prefix is known on the host: no lists at all (only the true
// entities), or MRE_FILTER_CSR with the caller's filt_nnz
inline bool known_is_flat(const RankParams &p) {
    return p.filter == MRE_FILTER_NONE || (p.filter == MRE_FILTER_CSR && p.filt_nnz > 0 && p.filt_nnz + p.Q < (1LL << 40));
}
}  // namespace mre
