// noc_ts_inst.cu — instantiation of the streamed tensor-core rollout (noc_ts_rollout.cuh) for the 50-agent swarm (d = 150).
#include "noc_ts_rollout.cuh"

namespace noc {
int launch_ts_swarm50(const TsArgs& A, const PhiRaw<float>& raw, int D, int r, int smem_limit, cudaStream_t st, double* out_sums) {
    return launch_ts<TsShape<50>>(A, raw, D, r, smem_limit, st, out_sums);
}
}  // namespace noc
