// jbugs_specs.cuh -- table of plate-kernel instantiations and the run-time selection among them.
//
// Variant 0 is the interpreter (`DynSpec`): any plate of the plan class.  The other variants are
// compile-time specialisations of the SAME kernel body for the plate shapes of the configurations
// named in BASELINE.json; they are selected when the run-time shape is identical to the static one.
//
// The kernels are instantiated in separate translation units (spec_*.cu) so that they compile in
// parallel; this header only declares their launchers.
#pragma once

#include <cstring>

#include "jbugs_kernels.cuh"

namespace jbugs {

// chains per CTA for a batch of C chains: no idle lanes for small batches
static inline int plate_block(long long C) { return C <= 32 ? 32 : (C <= 64 ? 64 : 128); }

#define JBUGS_LAUNCH_PARAMS                                                                                     \
    long long n_tiles, size_t smem, cudaStream_t st, const PlateParams &pp, const double *theta, double *grad, \
        long long ld, long long C, const double *gvals, long long ldg, double *partial, int *flags, int want_grad, int pdl
#define JBUGS_LAUNCH_ARGS n_tiles, smem, st, pp, theta, grad, ld, C, gvals, ldg, partial, flags, want_grad, pdl

typedef cudaError_t (*plate_launch_fn)(JBUGS_LAUNCH_PARAMS);

// one-launch evaluation of a small single-plate model (fused_small_kernel)
#define JBUGS_FUSED_PARAMS                                                                                             \
    long long n_tiles, size_t smem, cudaStream_t st, const FusedParams &fq, const double *theta, double *grad, long long ld, \
        long long C, double *gvals, long long ldg, double *partial, int *flags, double *logp, int *status, int *any_bad,  \
        int want_grad, unsigned *tickets, int pdl
#define JBUGS_FUSED_ARGS n_tiles, smem, st, fq, theta, grad, ld, C, gvals, ldg, partial, flags, logp, status, any_bad, want_grad, tickets, pdl
typedef cudaError_t (*fused_launch_fn)(JBUGS_FUSED_PARAMS);
cudaError_t launch_fused_dyn(JBUGS_FUSED_PARAMS);
cudaError_t launch_fused_pumps(JBUGS_FUSED_PARAMS);
cudaError_t launch_fused_epil(JBUGS_FUSED_PARAMS);
cudaError_t launch_fused_epil_t4(JBUGS_FUSED_PARAMS);

// defined in spec_dyn.cu / spec_rats.cu / spec_glm.cu / spec_epil.cu
cudaError_t launch_dyn(JBUGS_LAUNCH_PARAMS);
cudaError_t launch_x(JBUGS_LAUNCH_PARAMS);  // expression-tape interpreter (XSpec)
cudaError_t launch_xw(JBUGS_LAUNCH_PARAMS); // ... for tapes that reference more than MAX_GREF_LINEAR global values (XWSpec)
cudaError_t launch_rats_t5(JBUGS_LAUNCH_PARAMS);
cudaError_t launch_rats(JBUGS_LAUNCH_PARAMS);
cudaError_t launch_seeds(JBUGS_LAUNCH_PARAMS);
cudaError_t launch_surgical(JBUGS_LAUNCH_PARAMS);
cudaError_t launch_pumps(JBUGS_LAUNCH_PARAMS);
cudaError_t launch_epil(JBUGS_LAUNCH_PARAMS);
cudaError_t launch_epil_t4(JBUGS_LAUNCH_PARAMS);

struct SpecEntry {
    const char* name;
    plate_launch_fn launch;
    bool (*matches)(const PlateShape&, long long T);
    int fused;  // FUSED_* code: tells the host whether the plate needs obs_const
    fused_launch_fn one_launch;  // whole evaluation of a small single-plate model in one kernel (nullptr: not built for this spec)
    int smpf_doubles;            // extra dynamic shared memory (doubles) of a spec with the shared-memory theta prefetch
};

#include "jbugs_static_specs.inc"

#ifdef JBUGS_SPEC_TABLE  // only the host translation unit owns the table
static bool match_never(const PlateShape&, long long) { return false; }

static const SpecEntry g_specs[] = {
    {"dynamic", &launch_dyn, &match_never, FUSED_NONE, &launch_fused_dyn, 0},
    {"rats_normal_identity_T5", &launch_rats_t5, &RatsDesc<5>::matches, RatsDesc<5>::FUSED, nullptr, 0},
    {"rats_normal_identity", &launch_rats, &RatsDesc<0>::matches, RatsDesc<0>::FUSED, nullptr, 0},
    {"seeds_binomial_logit", &launch_seeds, &SeedsDesc::matches, SeedsDesc::FUSED, nullptr, SeedsDesc::SMPF ? 2 * SeedsDesc::UNROLL * MAX_LOC * 128 : 0},
    {"surgical_binomial_logit", &launch_surgical, &SurgicalDesc::matches, SurgicalDesc::FUSED, nullptr, SurgicalDesc::SMPF ? 2 * SurgicalDesc::UNROLL * MAX_LOC * 128 : 0},
    {"epil_poisson_log_T4", &launch_epil_t4, &EpilDesc<4>::matches, EpilDesc<4>::FUSED, &launch_fused_epil_t4, 0},
    {"epil_poisson_log", &launch_epil, &EpilDesc<0>::matches, EpilDesc<0>::FUSED, &launch_fused_epil, 0},
    {"pumps_poisson_gamma", &launch_pumps, &PumpsDesc::matches, PumpsDesc::FUSED, &launch_fused_pumps, 0},
    {"expression_tape", &launch_x, &match_never, FUSED_NONE, nullptr, 0},
    {"expression_tape_wide", &launch_xw, &match_never, FUSED_NONE, nullptr, 0},
};
static const int g_num_specs = (int)(sizeof(g_specs) / sizeof(g_specs[0]));
static const int g_xspec = g_num_specs - 2;   // the tape interpreter: selected by the plate kind, never by shape
static const int g_xwspec = g_num_specs - 1;  // ... when the tape references more than MAX_GREF_LINEAR global values
static inline bool is_tape_variant(int v) { return v == g_xspec || v == g_xwspec; }

static int select_spec(const PlateShape& sh, long long T) {
    for (int k = 1; k < g_num_specs; ++k)
        if (g_specs[k].matches(sh, T)) return k;
    return 0;
}
static const char* spec_name(int variant) { return g_specs[variant].name; }
#endif

}  // namespace jbugs
