// jbugs_specs.cuh -- table of plate-kernel instantiations and the run-time selection among them.
//
// Variant 0 is the interpreter (`DynSpec`): any plate of the plan class.  The other variants are
// compile-time specialisations of the SAME kernel body for the plate shapes of the configurations
// named in BASELINE.json; they are selected when the run-time shape is identical to the static one.
#pragma once

#include <cstring>

#include "jbugs_kernels.cuh"

namespace jbugs {

typedef void (*plate_launch_fn)(dim3 grid, cudaStream_t st, const PlateParams& pp, const double* theta, double* grad,
                                long long ld, long long C, const double* gvals, long long ldg, double* partial,
                                int* flags, int want_grad);

template <class S>
static void launch_plate_t(dim3 grid, cudaStream_t st, const PlateParams& pp, const double* theta, double* grad,
                           long long ld, long long C, const double* gvals, long long ldg, double* partial, int* flags,
                           int want_grad) {
    plate_kernel<S><<<grid, BLOCK, 0, st>>>(pp, theta, grad, ld, C, gvals, ldg, partial, flags, want_grad);
}

struct SpecEntry {
    const char* name;
    plate_launch_fn launch;
    bool (*matches)(const PlateShape&);
    int fused;  // FUSED_* code: tells the host whether the plate needs obs_const
};

static bool match_never(const PlateShape&) { return false; }

#include "jbugs_static_specs.inc"

static const SpecEntry g_specs[] = {
    {"dynamic", &launch_plate_t<DynSpec>, &match_never, FUSED_NONE},
    JBUGS_STATIC_SPEC_ENTRIES
};
static const int g_num_specs = (int)(sizeof(g_specs) / sizeof(g_specs[0]));

static int select_spec(const PlateShape& sh) {
    for (int k = 1; k < g_num_specs; ++k)
        if (g_specs[k].matches(sh)) return k;
    return 0;
}
static const char* spec_name(int variant) { return g_specs[variant].name; }

static void launch_plate(int variant, dim3 grid, cudaStream_t st, const PlateParams& pp, const double* theta,
                         double* grad, long long ld, long long C, const double* gvals, long long ldg, double* partial,
                         int* flags, int want_grad) {
    g_specs[variant].launch(grid, st, pp, theta, grad, ld, C, gvals, ldg, partial, flags, want_grad);
}

}  // namespace jbugs
