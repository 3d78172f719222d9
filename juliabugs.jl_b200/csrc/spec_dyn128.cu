// interpreter instantiation of the plate kernel, 128 chains per CTA
#include "jbugs_launch.cuh"
namespace jbugs {
cudaError_t launch_dyn_128(JBUGS_LAUNCH_PARAMS) { return launch_plate_blk<DynSpec, 128>(JBUGS_LAUNCH_ARGS); }
}  // namespace jbugs
