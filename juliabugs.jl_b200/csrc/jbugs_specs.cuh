// jbugs_specs.cuh -- table of plate-kernel instantiations and the run-time selection among them.
//
// Variant 0 is the interpreter (`DynSpec`): any plate of the plan class.  The other variants are
// compile-time specialisations of the SAME kernel body for the plate shapes of the configurations
// named in BASELINE.json; they are selected when the run-time shape is identical to the static one.
#pragma once

#include <cstring>

#include "jbugs_kernels.cuh"

namespace jbugs {

// chains per CTA for a batch of C chains: no idle lanes for small batches
static inline int plate_block(long long C) { return C <= 32 ? 32 : (C <= 64 ? 64 : 128); }

typedef cudaError_t (*plate_launch_fn)(long long n_tiles, size_t smem, cudaStream_t st, const PlateParams& pp,
                                       const double* theta, double* grad, long long ld, long long C,
                                       const double* gvals, long long ldg, double* partial, int* flags,
                                       int want_grad);

template <class S, int BLK>
static cudaError_t launch_plate_blk(long long n_tiles, size_t smem, cudaStream_t st, const PlateParams& pp,
                                    const double* theta, double* grad, long long ld, long long C, const double* gvals,
                                    long long ldg, double* partial, int* flags, int want_grad) {
    static size_t opted_in = 48 * 1024;  // dynamic shared memory above 48 KB needs an opt-in per kernel
    if (smem > opted_in) {
        cudaError_t e = cudaFuncSetAttribute(plate_kernel<S, BLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        opted_in = smem;
    }
    dim3 grid((unsigned)((C + BLK - 1) / BLK), (unsigned)n_tiles);
    plate_kernel<S, BLK><<<grid, BLK, smem, st>>>(pp, theta, grad, ld, C, gvals, ldg, partial, flags, want_grad);
    return cudaGetLastError();
}

template <class S>
static cudaError_t launch_plate_t(long long n_tiles, size_t smem, cudaStream_t st, const PlateParams& pp,
                                  const double* theta, double* grad, long long ld, long long C, const double* gvals,
                                  long long ldg, double* partial, int* flags, int want_grad) {
    switch (plate_block(C)) {
    case 32: return launch_plate_blk<S, 32>(n_tiles, smem, st, pp, theta, grad, ld, C, gvals, ldg, partial, flags, want_grad);
    case 64: return launch_plate_blk<S, 64>(n_tiles, smem, st, pp, theta, grad, ld, C, gvals, ldg, partial, flags, want_grad);
    default: return launch_plate_blk<S, 128>(n_tiles, smem, st, pp, theta, grad, ld, C, gvals, ldg, partial, flags, want_grad);
    }
}

struct SpecEntry {
    const char* name;
    plate_launch_fn launch;
    bool (*matches)(const PlateShape&, long long T);
    int fused;  // FUSED_* code: tells the host whether the plate needs obs_const
};

static bool match_never(const PlateShape&, long long) { return false; }

#include "jbugs_static_specs.inc"

static const SpecEntry g_specs[] = {
    {"dynamic", &launch_plate_t<DynSpec>, &match_never, FUSED_NONE},
    JBUGS_STATIC_SPEC_ENTRIES
};
static const int g_num_specs = (int)(sizeof(g_specs) / sizeof(g_specs[0]));

static int select_spec(const PlateShape& sh, long long T) {
    for (int k = 1; k < g_num_specs; ++k)
        if (g_specs[k].matches(sh, T)) return k;
    return 0;
}
static const char* spec_name(int variant) { return g_specs[variant].name; }

}  // namespace jbugs
