// spec_dispatch.cu -- per-spec launchers: pick the CTA width (chains per CTA) from the batch size.
#include "jbugs_specs.cuh"
namespace jbugs {
#define JB_DISPATCH(name)                                        \
    cudaError_t name##_32(JBUGS_LAUNCH_PARAMS);                  \
    cudaError_t name##_64(JBUGS_LAUNCH_PARAMS);                  \
    cudaError_t name##_128(JBUGS_LAUNCH_PARAMS);                 \
    cudaError_t name##_128t(JBUGS_LAUNCH_PARAMS);                \
    cudaError_t name(JBUGS_LAUNCH_PARAMS) {                      \
        switch (plate_block(C)) {                                \
        case 32: return name##_32(JBUGS_LAUNCH_ARGS);            \
        case 64: return name##_64(JBUGS_LAUNCH_ARGS);            \
        default: /* TMA-staged instantiation on request */       \
            return pp.v.bulk_ok ? name##_128t(JBUGS_LAUNCH_ARGS) : name##_128(JBUGS_LAUNCH_ARGS); \
        }                                                        \
    }
// the same, with a forward-only kernel for logdensity without the gradient (128-chain CTAs)
#define JB_DISPATCH_NG(name)                                     \
    cudaError_t name##_32(JBUGS_LAUNCH_PARAMS);                  \
    cudaError_t name##_64(JBUGS_LAUNCH_PARAMS);                  \
    cudaError_t name##_128(JBUGS_LAUNCH_PARAMS);                 \
    cudaError_t name##_128t(JBUGS_LAUNCH_PARAMS);                \
    cudaError_t name##_128ng(JBUGS_LAUNCH_PARAMS);               \
    cudaError_t name(JBUGS_LAUNCH_PARAMS) {                      \
        switch (plate_block(C)) {                                \
        case 32: return name##_32(JBUGS_LAUNCH_ARGS);            \
        case 64: return name##_64(JBUGS_LAUNCH_ARGS);            \
        default:                                                 \
            if (!want_grad) return name##_128ng(JBUGS_LAUNCH_ARGS); \
            return pp.v.bulk_ok ? name##_128t(JBUGS_LAUNCH_ARGS) : name##_128(JBUGS_LAUNCH_ARGS); \
        }                                                        \
    }
JB_DISPATCH(launch_dyn)
JB_DISPATCH(launch_x)
JB_DISPATCH(launch_xw)
JB_DISPATCH_NG(launch_rats_t5)
JB_DISPATCH(launch_rats)
JB_DISPATCH_NG(launch_seeds)
JB_DISPATCH(launch_surgical)
JB_DISPATCH(launch_pumps)
JB_DISPATCH(launch_epil)
JB_DISPATCH(launch_epil_t4)
// one-launch kernels for small single-plate models
#define JB_DISPATCH_FUSED(name)                                  \
    cudaError_t name##_32(JBUGS_FUSED_PARAMS);                   \
    cudaError_t name##_64(JBUGS_FUSED_PARAMS);                   \
    cudaError_t name##_128(JBUGS_FUSED_PARAMS);                  \
    cudaError_t name(JBUGS_FUSED_PARAMS) {                       \
        switch (plate_block(C)) {                                \
        case 32: return name##_32(JBUGS_FUSED_ARGS);             \
        case 64: return name##_64(JBUGS_FUSED_ARGS);             \
        default: return name##_128(JBUGS_FUSED_ARGS);            \
        }                                                        \
    }
JB_DISPATCH_FUSED(launch_fused_dyn)
JB_DISPATCH_FUSED(launch_fused_pumps)
JB_DISPATCH_FUSED(launch_fused_epil)
JB_DISPATCH_FUSED(launch_fused_epil_t4)
}  // namespace jbugs
