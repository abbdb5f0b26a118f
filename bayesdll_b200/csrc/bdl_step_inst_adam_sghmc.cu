// bdl_step_inst_adam_sghmc.cu -- instantiates every build of the fused step for (BDL_ADAM_SGHMC, SGD momentum buffer = false); see bdl_step.cuh.
#define BDL_STEP_INSTANTIATE
#include "bdl_step.cuh"

#ifndef BDL_AB_SLIM      // slim A/B builds (tools/ab_builds.py) instantiate the two kernels they time from bdl_step.cu
namespace bdl {
template int launch_nd<BDL_ADAM_SGHMC, false>(const StepParams&, bool, int, cudaStream_t);
}  // namespace bdl
#endif
