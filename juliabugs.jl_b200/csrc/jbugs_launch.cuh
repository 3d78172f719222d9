// jbugs_launch.cuh -- launcher template for plate_kernel<S, BLK, TMA>; included only by the spec_*.cu translation
// units, each of which instantiates it for its own specs (so the kernels compile in parallel).
#pragma once

#include "jbugs_specs.cuh"

namespace jbugs {

template <class S, int BLK, bool TMA, bool GRAD = true>
static cudaError_t launch_plate_blk(JBUGS_LAUNCH_PARAMS) {
    // dynamic shared memory above 48 KB needs an opt-in per kernel AND per device (a process may hold plans on several)
    static size_t opted_in[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& have = opted_in[dev & 63];
    if (smem > 48 * 1024 && smem > have) {
        cudaError_t e = cudaFuncSetAttribute(plate_kernel<S, BLK, TMA, GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        have = smem;
    }
    dim3 grid((unsigned)((C + BLK - 1) / BLK), (unsigned)n_tiles);
    if (pdl) {  // programmatic dependent launch: overlap this kernel's start with the tail of its predecessor
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(BLK);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, plate_kernel<S, BLK, TMA, GRAD>, pp, theta, grad, ld, C, gvals, ldg, partial, flags, want_grad);
    }
    plate_kernel<S, BLK, TMA, GRAD><<<grid, BLK, smem, st>>>(pp, theta, grad, ld, C, gvals, ldg, partial, flags, want_grad);
    return cudaGetLastError();
}

// one-launch evaluation of a small single-plate model (fused_small_kernel)
template <class S, int BLK>
static cudaError_t launch_fused_blk(JBUGS_FUSED_PARAMS) {
    static size_t opted_in[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& have = opted_in[dev & 63];
    if (smem > 48 * 1024 && smem > have) {
        cudaError_t e = cudaFuncSetAttribute(fused_small_kernel<S, BLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        have = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((C + BLK - 1) / BLK), (unsigned)n_tiles);
    cfg.blockDim = dim3(BLK);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, fused_small_kernel<S, BLK>, fq, theta, grad, ld, C, gvals, ldg, partial, flags, logp, status, any_bad,
                              want_grad, tickets);
}

}  // namespace jbugs
