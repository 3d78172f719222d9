// Rats-shaped plates: normal likelihood, identity link, two group-level normal random effects
#include "jbugs_launch.cuh"
namespace jbugs {
cudaError_t launch_rats_t5(JBUGS_LAUNCH_PARAMS) { return launch_plate_t<StaticSpec<RatsDesc<5>>>(JBUGS_LAUNCH_ARGS); }
cudaError_t launch_rats(JBUGS_LAUNCH_PARAMS) { return launch_plate_t<StaticSpec<RatsDesc<0>>>(JBUGS_LAUNCH_ARGS); }
}  // namespace jbugs
