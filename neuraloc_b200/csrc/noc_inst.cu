// noc_inst.cu — instantiates the rollout kernel for ONE tile configuration (-DNOC_CFG_ID=0..8) so that the
// configurations compile in parallel.
#include "noc_launch.cuh"

#ifndef NOC_CFG_ID
#error "compile with -DNOC_CFG_ID=<0..8>"
#endif

namespace noc {
#define NOC_DEF_LAUNCH(ID, CFG, REAL)                                                                              \
    int launch_cfg_##ID(const RolloutArgs<REAL>& A, const PhiRaw<REAL>* raw, int kmode, size_t smem, cudaStream_t st, \
                        double* out_sums) {                                                                        \
        return launch_cfg<CFG, REAL>(A, raw, kmode, smem, st, out_sums);                                           \
    }
#if NOC_CFG_ID == 0
NOC_DEF_LAUNCH(0, CfgF_S4, float)
#elif NOC_CFG_ID == 1
NOC_DEF_LAUNCH(1, CfgF_S8, float)
#elif NOC_CFG_ID == 2
NOC_DEF_LAUNCH(2, CfgF_M, float)
#elif NOC_CFG_ID == 3
NOC_DEF_LAUNCH(3, CfgF_L, float)
#elif NOC_CFG_ID == 4
NOC_DEF_LAUNCH(4, CfgD_S8, double)
#elif NOC_CFG_ID == 5
NOC_DEF_LAUNCH(5, CfgD_M, double)
#elif NOC_CFG_ID == 6
NOC_DEF_LAUNCH(6, CfgD_L, double)
#elif NOC_CFG_ID == 7
NOC_DEF_LAUNCH(7, CfgF_S8Z, float)
#elif NOC_CFG_ID == 8
NOC_DEF_LAUNCH(8, CfgF_S4Z, float)
#endif
}  // namespace noc
