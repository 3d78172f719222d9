// jbugs_obsconst.cuh -- data-only normalising constants of the fused observation paths.
//
// `Distributions.logpdf(Binomial(n, p), k)` carries -logbeta(k+1, n-k+1) - log(n+1) and
// `logpdf(Poisson(lambda), k)` carries -loggamma(k+1); neither depends on theta, so the fused
// kernels leave them out of the per-(obs, chain) work and this kernel adds them up ONCE per
// (plate, shard, data upload) in a fixed order (one CTA, strided partial sums, tree in shared
// memory).  It also validates the data the fused formulas assume (k integer, 0 <= k <= n):
// invalid data makes the host fall back to the interpreter kernel, which reproduces the
// reference's -Inf semantics.
#pragma once

#include "jbugs_kernels.cuh"

namespace jbugs {

__global__ void __launch_bounds__(1024)
obs_const_kernel(int fused, const double* __restrict__ y, const double* __restrict__ n, int y_ij, int n_kind,
                 double n_const, long long N, long long T, long long i0, long long i1, double* __restrict__ out /*[2]*/) {
    __shared__ double s_sum[1024];
    __shared__ int s_bad[1024];
    const long long count = (i1 - i0) * T;
    double acc = 0.0;
    int bad = 0;
    for (long long q = threadIdx.x; q < count; q += 1024) {
        const long long i = i0 + q / T, j = q % T;
        const double k = y_ij ? y[i + N * j] : y[i];
        if (fused == FUSED_POISSON_LOG || fused == FUSED_POISSON_LINK) {
            bad |= !(k >= 0.0) || k != floor(k);
            acc -= lgamma(k + 1.0);
        } else if (fused == FUSED_BINOMIAL_LOGIT || fused == FUSED_BINOMIAL_LINK) {
            const double nn = n_kind == JBUGS_ARG_DATA_IJ ? n[i + N * j] : (n_kind == JBUGS_ARG_DATA_I ? n[i] : n_const);
            bad |= !(nn >= 0.0) || !(k >= 0.0) || k > nn || k != floor(k);
            acc -= logbeta(k + 1.0, nn - k + 1.0) + log(nn + 1.0);
        } else if (fused == FUSED_BERNOULLI_LOGIT || fused == FUSED_BERNOULLI_LINK) {
            bad |= !(k == 0.0 || k == 1.0);
        }
    }
    s_sum[threadIdx.x] = acc;
    s_bad[threadIdx.x] = bad;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) {
            s_sum[threadIdx.x] += s_sum[threadIdx.x + w];
            s_bad[threadIdx.x] |= s_bad[threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = s_sum[0];
        out[1] = s_bad[0] ? 1.0 : 0.0;
    }
}

}  // namespace jbugs
