// spec_one.cu -- ONE instantiation of the plate kernel per object file, selected on the compile line:
//   -DJB_SPEC='StaticSpec<RatsDesc<5>>' -DJB_BLK=128 -DJB_TMA=false -DJB_NAME=launch_rats_t5_128
// (small translation units compile in parallel instead of one big one.)
#include "jbugs_launch.cuh"
namespace jbugs {
#ifndef JB_GRAD
#define JB_GRAD true
#endif
#ifdef JB_ONE_LAUNCH  // the one-launch kernel for small single-plate models instead of the plate kernel
cudaError_t JB_NAME(JBUGS_FUSED_PARAMS) { return launch_fused_blk<JB_SPEC, JB_BLK>(JBUGS_FUSED_ARGS); }
#else
cudaError_t JB_NAME(JBUGS_LAUNCH_PARAMS) { return launch_plate_blk<JB_SPEC, JB_BLK, JB_TMA, JB_GRAD>(JBUGS_LAUNCH_ARGS); }
#endif
}  // namespace jbugs
