// Seeds / Surgical (binomial-logit with a normal random effect) and Pumps (Poisson-gamma)
#include "jbugs_launch.cuh"
namespace jbugs {
cudaError_t launch_seeds(JBUGS_LAUNCH_PARAMS) { return launch_plate_t<StaticSpec<SeedsDesc>>(JBUGS_LAUNCH_ARGS); }
cudaError_t launch_surgical(JBUGS_LAUNCH_PARAMS) { return launch_plate_t<StaticSpec<SurgicalDesc>>(JBUGS_LAUNCH_ARGS); }
cudaError_t launch_pumps(JBUGS_LAUNCH_PARAMS) { return launch_plate_t<StaticSpec<PumpsDesc>>(JBUGS_LAUNCH_ARGS); }
}  // namespace jbugs
