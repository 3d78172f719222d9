// Epil-shaped plates: Poisson likelihood, log link, subject and subject-by-visit normal random effects
#include "jbugs_launch.cuh"
namespace jbugs {
cudaError_t launch_epil(JBUGS_LAUNCH_PARAMS) { return launch_plate_t<StaticSpec<EpilDesc<0>>>(JBUGS_LAUNCH_ARGS); }
}  // namespace jbugs
