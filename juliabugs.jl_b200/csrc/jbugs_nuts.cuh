// jbugs_nuts.cuh -- batched No-U-Turn sampler on top of the HMC state (the reference samples with
// `NUTS(0.8)` through AdvancedHMC, JuliaBUGS/ext/JuliaBUGSAdvancedHMCExt.jl:40-66).
//
// Multinomial NUTS with the generalised U-turn criterion, built iteratively (no recursion): a doubling of depth j
// appends 2^j leapfrog leaves to one end of the trajectory; sub-tree U-turns are checked against O(depth) momentum
// checkpoints indexed by the bits of the leaf number (the scheme of Phan et al., "Composable effects for flexible and
// accelerated probabilistic programming in NumPyro", 2019, restated).  All chains of the batch run the same leaf at the
// same time, each in its own direction; a chain whose tree has stopped idles (dir = 0) until every chain has stopped.
//
// Arrays are chain-fast [D][ld]; the per-chain scalars live in NutsChain columns.
#pragma once

#include "jbugs_hmc.cuh"

namespace jbugs {

constexpr int NUTS_MAX_DEPTH = 12;

struct NutsState {
    // trajectory ends, tree / sub-tree proposals, momentum sums, checkpoints: [D][ld] each
    double *qL, *pL, *gL, *qR, *pR, *gR, *qT, *gT, *qS, *gS, *rhoT, *rhoS;
    double* ck_p;    // [NUTS_MAX_DEPTH][D][ld]
    double* ck_rho;  // [NUTS_MAX_DEPTH][D][ld]
    // per chain [ld]
    double *H0, *logwT, *logwS, *lpT, *lpS, *sum_acc, *n_leaf;
    int *dir, *active, *takeS, *takeT, *valid, *depth;
    double* partial;  // [n_tiles][1 + 2 * NUTS_MAX_DEPTH][ld]
    int* any_active;  // 1
    int* leaf;        // 1: index of the leaf being built inside the current sub-tree
    int* n_div;       // 1: divergent transitions so far
    int* cur_depth;   // 1: depth of the doubling under construction (read by the captured leaf sequence)
};

__device__ __forceinline__ double log_add_exp(double a, double b) {
    if (a == -dev_inf()) return b;
    if (b == -dev_inf()) return a;
    const double m = a > b ? a : b;
    return m + log1p(exp(-fabs(a - b)));
}

// start of a transition: the tree is the single point (q_cur, p)
__global__ void __launch_bounds__(BLOCK)
nuts_begin_rows_kernel(NutsState S, const double* __restrict__ q, const double* __restrict__ p, const double* __restrict__ g,
                       long long ld, long long D, long long C, long long ROWS) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    const long long d0 = (long long)blockIdx.y * ROWS, d1 = d0 + ROWS < D ? d0 + ROWS : D;
    for (long long d = d0; d < d1; ++d) {
        const long long e = d * ld + c;
        const double qv = q[e], pv = p[e], gv = g[e];
        S.qL[e] = qv; S.qR[e] = qv; S.qT[e] = qv;
        S.pL[e] = pv; S.pR[e] = pv; S.rhoT[e] = pv;
        S.gL[e] = gv; S.gR[e] = gv; S.gT[e] = gv;
    }
}
__global__ void __launch_bounds__(BLOCK)
nuts_begin_chain_kernel(NutsState S, const double* __restrict__ lp_cur, const double* __restrict__ k0, long long C,
                        int reset_stats) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c == 0) {
        *S.any_active = 1;
        if (reset_stats) *S.n_div = 0;  // first transition of a run: divergences are counted per run
    }
    if (c >= C) return;
    S.H0[c] = -lp_cur[c] + k0[c];
    S.logwT[c] = 0.0;  // weight of the initial point: exp(H0 - H0)
    S.lpT[c] = lp_cur[c];
    S.sum_acc[c] = 0.0;
    S.n_leaf[c] = 0.0;
    S.active[c] = (lp_cur[c] == lp_cur[c] && fabs(lp_cur[c]) != dev_inf()) ? 1 : 0;
    S.depth[c] = 0;
}

// start of a doubling: direction per chain, working state <- the end that grows, empty sub-tree
__global__ void __launch_bounds__(BLOCK)
nuts_dir_kernel(NutsState S, long long C, unsigned long long seed, const HmcCtl* __restrict__ ctl, int depth) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c == 0) {
        *S.leaf = 0;
        *S.any_active = 0;
        *S.cur_depth = depth;
    }
    if (c >= C) return;
    const uint4 r = philox4x32_10(make_uint4((unsigned)c, (unsigned)(c >> 32), ctl->it, 16u + (unsigned)depth),
                                  make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    S.dir[c] = S.active[c] ? ((r.x & 1u) ? 1 : -1) : 0;
    S.logwS[c] = -dev_inf();
    S.valid[c] = S.active[c];  // cleared when the sub-tree turns or diverges
    S.takeS[c] = 0;
    if (S.active[c]) S.depth[c] = depth + 1;
}
__global__ void __launch_bounds__(BLOCK)
nuts_load_end_kernel(NutsState S, double* __restrict__ qw, double* __restrict__ pw, double* __restrict__ gw, long long ld,
                     long long D, long long C, long long ROWS) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    const int dir = S.dir[c];
    if (dir == 0) return;
    const long long d0 = (long long)blockIdx.y * ROWS, d1 = d0 + ROWS < D ? d0 + ROWS : D;
    for (long long d = d0; d < d1; ++d) {
        const long long e = d * ld + c;
        qw[e] = dir > 0 ? S.qR[e] : S.qL[e];
        pw[e] = dir > 0 ? S.pR[e] : S.pL[e];
        gw[e] = dir > 0 ? S.gR[e] : S.gL[e];
        S.rhoS[e] = 0.0;
    }
}

// first half of a leapfrog step in the chain's own direction:  p += dir eps/2 g ;  q += dir eps var p
__global__ void __launch_bounds__(BLOCK)
nuts_leap_a_kernel(NutsState S, double* __restrict__ q, double* __restrict__ p, const double* __restrict__ g, long long ld,
                   long long D, long long C, const double* __restrict__ var, const HmcCtl* __restrict__ ctl,
                   const double* __restrict__ eps_c, long long ROWS) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    const int dir = S.valid[c] ? S.dir[c] : 0;
    if (dir == 0) return;
    const double eps = ctl->eps * eps_c[c] * (double)dir;
    const long long d0 = (long long)blockIdx.y * ROWS, d1 = d0 + ROWS < D ? d0 + ROWS : D;
    for (long long d = d0; d < d1; d += 4) {
        double gv[4], pv[4], qv[4], vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long dd = d + u < d1 ? d + u : d1 - 1, e = dd * ld + c;
            gv[u] = g[e]; pv[u] = p[e]; qv[u] = q[e]; vv[u] = var[dd];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (d + u < d1) {
                const long long e = (d + u) * ld + c;
                const double pn = fma(0.5 * eps, gv[u], pv[u]);
                p[e] = pn;
                q[e] = fma(eps * vv[u], pn, qv[u]);
            }
    }
}

// second half after the evaluation: p += dir eps/2 g; then the leaf's bookkeeping over this CTA's rows:
// kinetic energy, rhoS += p, checkpoint (even leaves) or sub-tree U-turn dot products (odd leaves)
__global__ void __launch_bounds__(BLOCK)
nuts_leaf_rows_kernel(NutsState S, double* __restrict__ p, const double* __restrict__ g, long long ld, long long D, long long C,
                      const double* __restrict__ var, const HmcCtl* __restrict__ ctl, const double* __restrict__ eps_c,
                      long long ROWS, int n_rows_partial) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    const int dir = S.valid[c] ? S.dir[c] : 0;
    const int n = *S.leaf;
    const int idx_max = __popc((unsigned)n >> 1);
    const int n_sub = (n & 1) ? __ffs(~(unsigned)n) - 1 : 0;  // trailing ones of an odd leaf number
    const int idx_min = idx_max - n_sub + 1;
    double* part = S.partial + (long long)blockIdx.y * n_rows_partial * ld + c;
    if (dir == 0) {
        part[0] = 0.0;
        return;
    }
    const double eps = ctl->eps * eps_c[c] * (double)dir;
    const long long d0 = (long long)blockIdx.y * ROWS, d1 = d0 + ROWS < D ? d0 + ROWS : D;
    const long long plane = D * ld;
    double k = 0.0;
    double da[NUTS_MAX_DEPTH], db[NUTS_MAX_DEPTH];
#pragma unroll
    for (int i = 0; i < NUTS_MAX_DEPTH; ++i) da[i] = db[i] = 0.0;
    for (long long d = d0; d < d1; ++d) {
        const long long e = d * ld + c;
        const double v = var[d];
        const double pn = fma(0.5 * eps, g[e], p[e]);
        p[e] = pn;
        k = fma(pn * pn, v, k);
        const double rho = S.rhoS[e] + pn;
        S.rhoS[e] = rho;
        if ((n & 1) == 0) {
            S.ck_p[idx_max * plane + e] = pn;
            S.ck_rho[idx_max * plane + e] = rho;
        } else {
#pragma unroll
            for (int i = 0; i < NUTS_MAX_DEPTH; ++i)
                if (i >= idx_min && i <= idx_max) {
                    const double cp = S.ck_p[i * plane + e];
                    const double sub = rho - S.ck_rho[i * plane + e] + cp;  // momentum sum of the sub-tree
                    da[i] = fma(sub * v, cp, da[i]);
                    db[i] = fma(sub * v, pn, db[i]);
                }
        }
    }
    part[0] = 0.5 * k;
    if (n & 1)
#pragma unroll
        for (int i = 0; i < NUTS_MAX_DEPTH; ++i)
            if (i >= idx_min && i <= idx_max) {
                part[(long long)(1 + 2 * i) * ld] = da[i];
                part[(long long)(2 + 2 * i) * ld] = db[i];
            }
}

// per chain: energy of the leaf, divergence, multinomial update of the sub-tree proposal, sub-tree U-turn
__global__ void __launch_bounds__(BLOCK)
nuts_leaf_chain_kernel(NutsState S, const double* __restrict__ lp_w, const int* __restrict__ status_w, long long ld, long long C,
                       int n_tiles, int n_rows_partial, unsigned long long seed, const HmcCtl* __restrict__ ctl) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    const int depth = *S.cur_depth;
    S.takeS[c] = 0;
    const int dir = S.valid[c] ? S.dir[c] : 0;
    if (dir == 0) return;
    const int n = *S.leaf;
    const int idx_max = __popc((unsigned)n >> 1);
    const int n_sub = (n & 1) ? __ffs(~(unsigned)n) - 1 : 0;
    const int idx_min = idx_max - n_sub + 1;
    const double* part = S.partial + c;
    const long long tstride = (long long)n_rows_partial * ld;
    double k = 0.0;
    for (int t = 0; t < n_tiles; ++t) k += part[t * tstride];
    const double H = -lp_w[c] + k;
    double delta = S.H0[c] - H;  // log weight of the leaf relative to the initial point
    const bool bad = !(delta == delta) || status_w[c] != 0 || fabs(H) == dev_inf();
    if (bad) delta = -dev_inf();
    S.sum_acc[c] += bad ? 0.0 : fmin(1.0, exp(delta));
    S.n_leaf[c] += 1.0;
    if (bad || delta < -1000.0) {  // divergent: the sub-tree is discarded and the tree stops
        S.valid[c] = 0;
        atomicAdd(S.n_div, 1);
        return;
    }
    const double lw = log_add_exp(S.logwS[c], delta);
    const uint4 r = philox4x32_10(make_uint4((unsigned)c, (unsigned)(c >> 32), ctl->it, 64u + (unsigned)depth * 4096u + (unsigned)n),
                                  make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    if (u01(r.x, r.y) < exp(delta - lw)) {
        S.takeS[c] = 1;
        S.lpS[c] = lp_w[c];
    }
    S.logwS[c] = lw;
    if (n & 1) {
        bool turning = false;
        for (int i = idx_min; i <= idx_max; ++i) {
            double a = 0.0, b = 0.0;
            for (int t = 0; t < n_tiles; ++t) {
                a += part[t * tstride + (long long)(1 + 2 * i) * ld];
                b += part[t * tstride + (long long)(2 + 2 * i) * ld];
            }
            turning |= (a <= 0.0) || (b <= 0.0);
        }
        if (turning) S.valid[c] = 0;
    }
}

// the leaf becomes the sub-tree's proposal where the multinomial draw said so; the last CTA row bumps the leaf counter
__global__ void __launch_bounds__(BLOCK)
nuts_take_leaf_kernel(NutsState S, const double* __restrict__ qw, const double* __restrict__ gw, long long ld, long long D,
                      long long C, long long ROWS) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C || !S.takeS[c]) return;
    const long long d0 = (long long)blockIdx.y * ROWS, d1 = d0 + ROWS < D ? d0 + ROWS : D;
    for (long long d = d0; d < d1; ++d) {
        const long long e = d * ld + c;
        S.qS[e] = qw[e];
        S.gS[e] = gw[e];
    }
}
__global__ void nuts_next_leaf_kernel(NutsState S) { *S.leaf += 1; }

// end of a doubling, per chain: a valid sub-tree joins the tree (biased progressive sampling of the proposal)
__global__ void __launch_bounds__(BLOCK)
nuts_merge_chain_kernel(NutsState S, long long C, unsigned long long seed, const HmcCtl* __restrict__ ctl, int depth) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    S.takeT[c] = 0;
    if (!S.active[c]) return;
    if (!S.valid[c]) {  // the new sub-tree turned or diverged: stop, keep the old tree
        S.active[c] = 0;
        return;
    }
    const uint4 r = philox4x32_10(make_uint4((unsigned)c, (unsigned)(c >> 32), ctl->it, 40u + (unsigned)depth),
                                  make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    if (u01(r.x, r.y) < exp(fmin(0.0, S.logwS[c] - S.logwT[c]))) {
        S.takeT[c] = 1;
        S.lpT[c] = S.lpS[c];
    }
    S.logwT[c] = log_add_exp(S.logwT[c], S.logwS[c]);
}
// rows: rhoT += rhoS, the grown end <- working state, tree proposal <- sub-tree proposal; whole-tree U-turn dot products
__global__ void __launch_bounds__(BLOCK)
nuts_merge_rows_kernel(NutsState S, const double* __restrict__ qw, const double* __restrict__ pw, const double* __restrict__ gw,
                       long long ld, long long D, long long C, const double* __restrict__ var, long long ROWS, int n_rows_partial) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    double* part = S.partial + (long long)blockIdx.y * n_rows_partial * ld + c;
    if (!S.active[c] || !S.valid[c]) {
        part[0] = 1.0;
        part[ld] = 1.0;
        return;
    }
    const int dir = S.dir[c], take = S.takeT[c];
    const long long d0 = (long long)blockIdx.y * ROWS, d1 = d0 + ROWS < D ? d0 + ROWS : D;
    double a = 0.0, b = 0.0;
    for (long long d = d0; d < d1; ++d) {
        const long long e = d * ld + c;
        const double rho = S.rhoT[e] + S.rhoS[e];
        S.rhoT[e] = rho;
        const double pwv = pw[e];
        if (dir > 0) { S.qR[e] = qw[e]; S.pR[e] = pwv; S.gR[e] = gw[e]; }
        else { S.qL[e] = qw[e]; S.pL[e] = pwv; S.gL[e] = gw[e]; }
        if (take) { S.qT[e] = S.qS[e]; S.gT[e] = S.gS[e]; }
        const double v = var[d];
        a = fma(rho * v, dir > 0 ? S.pL[e] : pwv, a);
        b = fma(rho * v, dir > 0 ? pwv : S.pR[e], b);
    }
    part[0] = a;
    part[ld] = b;
}
__global__ void __launch_bounds__(BLOCK)
nuts_merge_finish_kernel(NutsState S, long long ld, long long C, int n_tiles, int n_rows_partial) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C || !S.active[c]) return;
    const double* part = S.partial + c;
    const long long tstride = (long long)n_rows_partial * ld;
    double a = 0.0, b = 0.0;
    for (int t = 0; t < n_tiles; ++t) {
        a += part[t * tstride];
        b += part[t * tstride + ld];
    }
    if (a <= 0.0 || b <= 0.0) S.active[c] = 0;  // the whole trajectory makes a U-turn: this was the last doubling
    if (S.active[c]) *S.any_active = 1;
}

// end of the transition: state <- tree proposal; acceptance statistic = mean over the leaves of min(1, exp(H0 - H))
__global__ void __launch_bounds__(BLOCK)
nuts_finish_rows_kernel(NutsState S, double* __restrict__ q, double* __restrict__ g, long long ld, long long D, long long C,
                        long long ROWS) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    const long long d0 = (long long)blockIdx.y * ROWS, d1 = d0 + ROWS < D ? d0 + ROWS : D;
    for (long long d = d0; d < d1; ++d) {
        const long long e = d * ld + c;
        q[e] = S.qT[e];
        g[e] = S.gT[e];
    }
}
__global__ void __launch_bounds__(BLOCK)
nuts_finish_chain_kernel(NutsState S, double* __restrict__ lp_cur, double* __restrict__ accept_prob, long long C) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    lp_cur[c] = S.lpT[c];
    accept_prob[c] = S.n_leaf[c] > 0.0 ? S.sum_acc[c] / S.n_leaf[c] : 0.0;
}

}  // namespace jbugs
