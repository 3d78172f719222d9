// interpreter instantiation of the plate kernel (any plate of the plan class); one translation unit per CTA width
#include "jbugs_specs.cuh"
namespace jbugs {
cudaError_t launch_dyn_32(JBUGS_LAUNCH_PARAMS);
cudaError_t launch_dyn_64(JBUGS_LAUNCH_PARAMS);
cudaError_t launch_dyn_128(JBUGS_LAUNCH_PARAMS);
cudaError_t launch_dyn(JBUGS_LAUNCH_PARAMS) {
    switch (plate_block(C)) {
    case 32: return launch_dyn_32(JBUGS_LAUNCH_ARGS);
    case 64: return launch_dyn_64(JBUGS_LAUNCH_ARGS);
    default: return launch_dyn_128(JBUGS_LAUNCH_ARGS);
    }
}
}  // namespace jbugs
