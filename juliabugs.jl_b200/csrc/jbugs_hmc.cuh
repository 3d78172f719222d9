// jbugs_hmc.cuh -- on-device batched Hamiltonian Monte Carlo around the batched evaluator.
//
// The reference samples with AdvancedHMC through `AbstractMCMC.sample(model, NUTS/HMC, n)`
// (JuliaBUGS/ext/JuliaBUGSAdvancedHMCExt.jl:40-66): every leapfrog step calls `logdensity_and_gradient` once, one
// chain per call, with theta on the host.  Here C chains advance in lock step, theta / momentum / gradient never
// leave HBM, and one iteration is  L x (leapfrog kernel + jbugs evaluator)  + a handful of per-chain kernels.
//
// All arrays are chain-fast [D][ld]; a thread owns one chain (column), exactly like the evaluator's kernels.
// Random numbers: Philox4x32-10 (counter = element / iteration / stream, key = seed): reproducible and independent
// of the launch geometry.
#pragma once

#include "jbugs_kernels.cuh"

namespace jbugs {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ double u01(unsigned hi, unsigned lo) {  // (0, 1)
    const unsigned long long b = (((unsigned long long)hi << 32) | lo) >> 11;
    return ((double)b + 0.5) * (1.0 / 9007199254740992.0);
}
// two independent standard normals for (element e, iteration it, stream s)
__device__ __forceinline__ double2 normal2(unsigned long long seed, unsigned long long e, unsigned it, unsigned s) {
    const uint4 r = philox4x32_10(make_uint4((unsigned)e, (unsigned)(e >> 32), it, s),
                                  make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const double u = u01(r.x, r.y), v = u01(r.z, r.w);
    const double rad = sqrt(-2.0 * log(u));
    double sn, cs;
    sincospi(2.0 * v, &sn, &cs);
    return make_double2(rad * cs, rad * sn);
}


// per-transition scalars live in device memory so that a captured CUDA graph of the trajectory can be replayed with a
// new step size / iteration counter (set by hmc_ctl_kernel just before the graph launch)
// per-chain step size and its dual-averaging state (Hoffman & Gelman 2014): every chain adapts to its own acceptance
// statistic, as the reference's independent chains do; the metric is pooled over chains
struct HmcDa {
    double *eps, *hbar, *logeps_bar, *mu;  // [ld] each
};

struct HmcCtl {
    double eps;           // multiplier of every chain's step size in this transition (the +-10 % jitter)
    unsigned it;
    unsigned pad;
    double mean_accept;   // mean acceptance probability of the last transition
    double accept_total;  // running sum of those means since the last reset
};

__global__ void __launch_bounds__(BLOCK) hmc_da_init_kernel(HmcDa da, long long C, double eps0) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    da.eps[c] = eps0;
    da.mu[c] = log(10.0 * eps0);
    da.hbar[c] = 0.0;
    da.logeps_bar[c] = log(eps0);
}
// m = number of updates since the last restart (1, 2, ...); final_ != 0: leave the averaged step size in place
__global__ void __launch_bounds__(BLOCK)
hmc_da_update_kernel(HmcDa da, const double* __restrict__ accept_prob, long long C, double m, int final_) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    const double delta = 0.8, gamma = 0.05, t0 = 10.0, kappa = 0.75;
    double a = accept_prob[c];
    if (!(a == a)) a = 0.0;
    const double hbar = (1.0 - 1.0 / (m + t0)) * da.hbar[c] + (delta - a) / (m + t0);
    const double logeps = da.mu[c] - sqrt(m) / gamma * hbar;
    const double eta = pow(m, -kappa);
    const double lb = eta * logeps + (1.0 - eta) * da.logeps_bar[c];
    da.hbar[c] = hbar;
    da.logeps_bar[c] = lb;
    const double e = exp(final_ ? lb : logeps);
    da.eps[c] = fmin(fmax(e, 1e-10), 1e4);
}
// new metric window: continue from the averaged step size
__global__ void __launch_bounds__(BLOCK) hmc_da_restart_kernel(HmcDa da, long long C) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    const double e = fmin(fmax(exp(da.logeps_bar[c]), 1e-10), 1e4);
    da.eps[c] = e;
    da.mu[c] = log(10.0 * e);
    da.hbar[c] = 0.0;
    da.logeps_bar[c] = log(e);
}

__global__ void hmc_ctl_kernel(HmcCtl* ctl, double eps, unsigned it, int reset_total) {
    ctl->eps = eps;
    ctl->it = it;
    if (reset_total) ctl->accept_total = 0.0;
}

// q[d][c] = theta0[d] + jitter * z
__global__ void __launch_bounds__(BLOCK)
hmc_init_kernel(double* __restrict__ q, long long ld, long long D, long long C, const double* __restrict__ theta0,
                double jitter, unsigned long long seed, long long HMC_ROWS) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    const long long d0 = (long long)blockIdx.y * HMC_ROWS, d1 = d0 + HMC_ROWS < D ? d0 + HMC_ROWS : D;
    for (long long d = d0; d < d1; ++d) q[d * ld + c] = theta0[d] + jitter * normal2(seed, d * C + c, 0u, 1u).x;
}

// chains whose start has no finite log density (or a DomainError) are drawn again closer to theta0: jitter / 2^attempt.
// `n_bad` counts them (set to 0 by the host before the launch; every row tile of a bad chain adds 1).
__global__ void __launch_bounds__(BLOCK)
hmc_reinit_kernel(double* __restrict__ q, long long ld, long long D, long long C, const double* __restrict__ theta0,
                  double jitter, unsigned long long seed, unsigned attempt, const double* __restrict__ lp,
                  const int* __restrict__ status, int* __restrict__ n_bad, long long HMC_ROWS) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    const double l = lp[c];
    if (l == l && fabs(l) != dev_inf() && status[c] == 0) return;
    if (blockIdx.y == 0) atomicAdd(n_bad, 1);
    const long long d0 = (long long)blockIdx.y * HMC_ROWS, d1 = d0 + HMC_ROWS < D ? d0 + HMC_ROWS : D;
    for (long long d = d0; d < d1; ++d) q[d * ld + c] = theta0[d] + jitter * normal2(seed, d * C + c, attempt, 5u).x;
}

// number of chains whose log density is not finite (or that were flagged DomainError)
__global__ void __launch_bounds__(BLOCK)
hmc_count_bad_kernel(const double* __restrict__ lp, const int* __restrict__ status, long long C, int* __restrict__ n_bad) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    const double l = lp[c];
    if (!(l == l && fabs(l) != dev_inf() && status[c] == 0)) atomicAdd(n_bad, 1);
}

// fresh momentum p ~ N(0, M), M = diag(1 / var);  kinetic partial sums 0.5 * sum p^2 var.
// One Philox call gives the normals of rows (d, d + 1) with d counted from the tile start: HMC_ROWS must be even
// (jbugs_hmc_create rounds it up) so that every tile starts on an even row and no two rows share a counter.
__global__ void __launch_bounds__(BLOCK)
hmc_refresh_kernel(double* __restrict__ p, long long ld, long long D, long long C, const double* __restrict__ var,
                   unsigned long long seed, const HmcCtl* __restrict__ ctl, double* __restrict__ kin_partial, long long ldg,
                   long long HMC_ROWS) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    const unsigned it = ctl->it;
    const long long d0 = (long long)blockIdx.y * HMC_ROWS, d1 = d0 + HMC_ROWS < D ? d0 + HMC_ROWS : D;
    double k = 0.0;
    for (long long d = d0; d < d1; d += 2) {
        const double2 z = normal2(seed, (d >> 1) * C + c, it, 2u);
        const double v0 = var[d];
        p[d * ld + c] = z.x * rsqrt(v0);
        k = fma(z.x, z.x, k);  // p^2 var = z^2
        if (d + 1 < d1) {
            p[(d + 1) * ld + c] = z.y * rsqrt(var[d + 1]);
            k = fma(z.y, z.y, k);
        }
    }
    kin_partial[(long long)blockIdx.y * ldg + c] = 0.5 * k;
}

// one leapfrog update:  p += wp * eps * g;  q_out = q_in + wq * eps * var * p     (wq = 0: momentum only)
// with `kin_partial` != null the kinetic energy of the updated momentum is accumulated per (row tile, chain)
__global__ void __launch_bounds__(BLOCK)
hmc_leap_kernel(const double* q_in, double* q_out /* may alias q_in */, double* __restrict__ p,
                const double* __restrict__ g, long long ld, long long D, long long C, const double* __restrict__ var,
                const HmcCtl* __restrict__ ctl, const double* __restrict__ eps_c, double wp, double wq,
                double* __restrict__ kin_partial, long long ldg, long long HMC_ROWS) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    const double eps = ctl->eps * eps_c[c];
    const long long d0 = (long long)blockIdx.y * HMC_ROWS, d1 = d0 + HMC_ROWS < D ? d0 + HMC_ROWS : D;
    double k = 0.0;
    // four rows at a time, every load issued before the first use (the index is clamped, not predicated)
    for (long long d = d0; d < d1; d += 4) {
        double gv[4], pv[4], qv[4], vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long dd = d + u < d1 ? d + u : d1 - 1, e = dd * ld + c;
            gv[u] = g[e];
            pv[u] = p[e];
            vv[u] = var[dd];
            qv[u] = wq != 0.0 ? q_in[e] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (d + u < d1) {
                const long long e = (d + u) * ld + c;
                const double pn = fma(wp * eps, gv[u], pv[u]);
                p[e] = pn;
                if (wq != 0.0) q_out[e] = fma(wq * eps * vv[u], pn, qv[u]);
                k = fma(pn * pn, vv[u], k);
            }
    }
    if (kin_partial) kin_partial[(long long)blockIdx.y * ldg + c] = 0.5 * k;
}

// per chain: sum the kinetic partials (ascending tiles); dst[c] = that sum
__global__ void __launch_bounds__(BLOCK)
hmc_colsum_kernel(const double* __restrict__ partial, int n_tiles, long long ldg, long long C, double* __restrict__ dst) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    double s = 0.0;
    for (int t0 = 0; t0 < n_tiles; t0 += 8) {  // ascending order, eight loads in flight
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = partial[(long long)(t0 + u < n_tiles ? t0 + u : n_tiles - 1) * ldg + c];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (t0 + u < n_tiles) s += v[u];
    }
    dst[c] = s;
}

// Metropolis test per chain.  H = -logp + K.  Non-finite proposals (divergences, DomainError) are rejected.
__global__ void __launch_bounds__(BLOCK)
hmc_accept_kernel(const double* lp0 /* == lp_cur */, const double* __restrict__ k0, const double* __restrict__ lp1,
                  const double* __restrict__ k1, const int* __restrict__ status1, long long C, unsigned long long seed,
                  const HmcCtl* __restrict__ ctl, int* __restrict__ accept, double* __restrict__ accept_prob, double* lp_cur) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    const unsigned it = ctl->it;
    const double h0 = -lp0[c] + k0[c], h1 = -lp1[c] + k1[c];
    double a = exp(fmin(0.0, h0 - h1));
    if (!(h1 == h1) || !(a == a) || status1[c] != 0 || fabs(h1) == dev_inf()) a = 0.0;
    const uint4 r = philox4x32_10(make_uint4((unsigned)c, (unsigned)(c >> 32), it, 3u), make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const int acc = u01(r.x, r.y) < a;
    accept[c] = acc;
    accept_prob[c] = a;
    if (acc) lp_cur[c] = lp1[c];
}

// mean acceptance probability over the chains (one CTA; fixed-order strided sums + shared-memory tree)
__global__ void __launch_bounds__(256)
hmc_mean_accept_kernel(const double* __restrict__ accept_prob, long long C, HmcCtl* __restrict__ ctl) {
    __shared__ double sh[256];
    double s = 0.0;
    for (long long c = threadIdx.x; c < C; c += 256) s += accept_prob[c];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double m = sh[0] / (double)C;
        ctl->mean_accept = m;
        ctl->accept_total += m;
    }
}

// accepted chains take the proposal:  q_cur, g_cur <- q_prop, g_prop
__global__ void __launch_bounds__(BLOCK)
hmc_select_kernel(double* __restrict__ q_cur, double* __restrict__ g_cur, const double* __restrict__ q_prop,
                  const double* __restrict__ g_prop, long long ld, long long D, long long C, const int* __restrict__ accept,
                  long long HMC_ROWS) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C || !accept[c]) return;
    const long long d0 = (long long)blockIdx.y * HMC_ROWS, d1 = d0 + HMC_ROWS < D ? d0 + HMC_ROWS : D;
    for (long long d = d0; d < d1; ++d) {
        const long long e = d * ld + c;
        q_cur[e] = q_prop[e];
        g_cur[e] = g_prop[e];
    }
}

// monitored rows of the current state -> out[m][c]
__global__ void __launch_bounds__(BLOCK)
hmc_monitor_kernel(const double* __restrict__ q, long long ld, long long C, const long long* __restrict__ rows, int n_rows,
                   double* __restrict__ out, long long ldo) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    for (int m = blockIdx.y; m < n_rows; m += gridDim.y) out[(long long)m * ldo + c] = q[rows[m] * ld + c];
}

// running first and second moments of every parameter, pooled over chains (one warp per row; fixed-order butterfly)
__global__ void __launch_bounds__(256)
hmc_moments_kernel(const double* __restrict__ q, long long ld, long long D, long long C, double* __restrict__ s1,
                   double* __restrict__ s2) {
    const long long d = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (d >= D) return;
    const int lane = threadIdx.x & 31;
    double a = 0.0, b = 0.0;
    for (long long c = lane; c < C; c += 32) {
        const double v = q[d * ld + c];
        a += v;
        b = fma(v, v, b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
        s1[d] += a;
        s2[d] += b;
    }
}

// var[d] from the pooled moments (regularised towards 1 like Stan's windowed adaptation), then reset the moments
__global__ void __launch_bounds__(BLOCK)
hmc_update_metric_kernel(double* __restrict__ var, double* __restrict__ s1, double* __restrict__ s2, long long D, double n) {
    const long long d = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (d >= D) return;
    const double mean = s1[d] / n;
    double v = (s2[d] - n * mean * mean) / (n - 1.0);
    v = (n / (n + 5.0)) * v + 1e-3 * (5.0 / (n + 5.0));
    var[d] = (v > 0.0 && v == v && fabs(v) != dev_inf()) ? v : 1.0;
    s1[d] = 0.0;
    s2[d] = 0.0;
}

__global__ void __launch_bounds__(BLOCK)
hmc_fill_kernel(double* __restrict__ x, long long n, double v) {
    const long long i = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (i < n) x[i] = v;
}

}  // namespace jbugs
