// interpreter instantiation of the plate kernel, 32 chains per CTA
#include "jbugs_launch.cuh"
namespace jbugs {
cudaError_t launch_dyn_32(JBUGS_LAUNCH_PARAMS) { return launch_plate_blk<DynSpec, 32>(JBUGS_LAUNCH_ARGS); }
}  // namespace jbugs
