// jbugs_gq.cuh -- generated quantities for S draws at once (include/jbugs_gq.h).
//
// thread = draw (column s of the chain-fast draws matrix), blockIdx.y = a tile of the statement's cells.  Theta rows and
// Q rows are read / written as coalesced row segments exactly like the evaluator's kernels; the tape, the per-cell row
// numbers and data values are warp-uniform (broadcast loads).  Forward values only: no adjoints here.
#pragma once

#include "../../include/jbugs_gq.h"
#include "jbugs_hmc.cuh"

namespace jbugs {

struct GqParams {
    const jbugs_gq_op_rec* ops;  // the statement's ops (device)
    int n_ops;
    long long n_cells, out_row0, cells_per_tile;
    unsigned family;
    int n_args;
    int arg_slot[3];
    const long long* ipool;
    const double* dpool;
    const int* tr;          // [D]
    const double *lb, *ub;  // [D]
    const int* in_row;      // [D]: input row of each theta index
    unsigned long long seed;
};

// counter-based stream of uniforms for one (row, draw): Philox4x32-10 blocks of four 32-bit words
struct GqRng {
    uint2 key;
    unsigned row_lo, row_hi, draw, n;
    uint4 buf;
    int have;
    __device__ GqRng(unsigned long long seed, unsigned long long row, unsigned long long draw_)
        : key(make_uint2((unsigned)seed, (unsigned)(seed >> 32))), row_lo((unsigned)row), row_hi((unsigned)(row >> 32) ^ (unsigned)(draw_ >> 32) * 0x9E3779B9u),
          draw((unsigned)draw_), n(0), have(0) {}
    __device__ double uniform() {  // (0, 1), 53 bits
        if (have < 2) {
            buf = philox4x32_10(make_uint4(row_lo, row_hi, draw, n++), key);
            have = 4;
        }
        const double u = have == 4 ? u01(buf.x, buf.y) : u01(buf.z, buf.w);
        have -= 2;
        return u;
    }
    __device__ double normal() {
        const double u = uniform(), v = uniform();
        double sn, cs;
        sincospi(2.0 * v, &sn, &cs);
        return sqrt(-2.0 * log(u)) * cs;
    }
    __device__ double exponential() { return -log(uniform()); }
    // Marsaglia & Tsang (2000); shape < 1 by the boost gamma(a + 1) * U^(1/a)
    __device__ double gamma(double a) {
        if (!(a > 0.0)) return dev_nan();
        double boost = 1.0;
        if (a < 1.0) {
            boost = pow(uniform(), 1.0 / a);
            a += 1.0;
        }
        const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
        for (int it = 0; it < 256; ++it) {
            const double x = normal();
            double v = 1.0 + c * x;
            if (v <= 0.0) continue;
            v = v * v * v;
            const double u = uniform();
            if (u < 1.0 - 0.0331 * (x * x) * (x * x) || log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return boost * d * v;
        }
        return boost * d;
    }
    // Poisson: multiplication of uniforms for small means, Hoermann's PTRS (1993) transformed rejection above 10
    __device__ double poisson(double lam) {
        if (!(lam >= 0.0)) return dev_nan();
        if (lam == 0.0) return 0.0;
        if (lam < 10.0) {
            const double L = exp(-lam);
            double p = uniform();
            int k = 0;
            while (p > L && k < 1000) {
                p *= uniform();
                ++k;
            }
            return (double)k;
        }
        const double slam = sqrt(lam), loglam = log(lam);
        const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b, inv_alpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
        for (int it = 0; it < 1000; ++it) {
            const double u = uniform() - 0.5, v = uniform();
            const double us = 0.5 - fabs(u);
            const double k = floor((2.0 * a / us + b) * u + lam + 0.43);
            if (us >= 0.07 && v <= vr) return k;
            if (k < 0.0 || (us < 0.013 && v > us)) continue;
            if (log(v) + log(inv_alpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + 1.0)) return k;
        }
        return floor(lam);
    }
    // Binomial(n, p): inversion by sequential search from 0 for n * min(p, 1-p) < 30, otherwise the sum of a normal-sized
    // split: Binomial(n, p) = Binomial(m, p) + Binomial(n - m, p) is exact, so large n is cut into pieces of mean < 30
    __device__ double binomial(double nd, double p) {
        if (!(p >= 0.0 && p <= 1.0) || !(nd >= 0.0)) return dev_nan();
        const bool flip = p > 0.5;
        const double q = flip ? 1.0 - p : p;
        long long n = (long long)nd;
        double total = 0.0;
        if (q == 0.0) return flip ? nd : 0.0;
        const long long piece = (long long)fmax(1.0, floor(30.0 / q));
        while (n > 0) {
            const long long m = n < piece ? n : piece;
            // inversion: P(X = 0) = (1 - q)^m, P(X = k + 1) = P(X = k) * (m - k) / (k + 1) * q / (1 - q)
            const double r = q / (1.0 - q);
            double pk = exp((double)m * log1p(-q)), cdf = pk;
            const double u = uniform();
            long long k = 0;
            while (u > cdf && k < m) {
                pk *= r * (double)(m - k) / (double)(k + 1);
                cdf += pk;
                ++k;
            }
            total += (double)k;
            n -= m;
        }
        return flip ? nd - total : total;
    }
};

// forward value of a unary / binary op of a generated-quantity tape
__device__ __forceinline__ double gq_unary(unsigned op, double x) {
    switch (op) {
    case JBUGS_OP_IDENTITY: return x;
    case JBUGS_OP_EXP: return exp(x);
    case JBUGS_OP_LOGISTIC: return logistic(x);
    case JBUGS_OP_CEXPEXP: return -expm1(-exp(x));
    case JBUGS_OP_PHI: return 0.5 * erfc(-x * kInvSqrt2);
    case JBUGS_OP_LOG: return log(x);
    case JBUGS_OP_SQRT: return sqrt(x);
    case JBUGS_OP_NEG: return -x;
    case JBUGS_GQ_OP_ABS: return fabs(x);
    case JBUGS_GQ_OP_STEP: return x >= 0.0 ? 1.0 : 0.0;
    case JBUGS_GQ_OP_LOGIT: return log(x / (1.0 - x));
    case JBUGS_GQ_OP_CLOGLOG: return log(-log1p(-x));
    case JBUGS_GQ_OP_PROBIT: return normcdfinv(x);
    case JBUGS_GQ_OP_SIN: return sin(x);
    case JBUGS_GQ_OP_COS: return cos(x);
    case JBUGS_GQ_OP_ROUND: return round(x);
    case JBUGS_GQ_OP_TRUNC: return trunc(x);
    case JBUGS_GQ_OP_LOGGAM: return lgamma(x);
    case JBUGS_GQ_OP_LOGFACT: return lgamma(x + 1.0);
    default: return dev_nan();
    }
}
__device__ __forceinline__ double gq_binary(unsigned op, double a, double b) {
    switch (op) {
    case JBUGS_OP_ADD: return a + b;
    case JBUGS_OP_SUB: return a - b;
    case JBUGS_OP_MUL: return a * b;
    case JBUGS_OP_DIV: return a / b;
    case JBUGS_OP_POW: return pow(a, b);
    case JBUGS_GQ_OP_EQUALS: return a == b ? 1.0 : 0.0;
    case JBUGS_GQ_OP_MAX: return fmax(a, b);
    case JBUGS_GQ_OP_MIN: return fmin(a, b);
    default: return dev_nan();
    }
}

// a draw from `family` in the BUGS parameterisation (BUGSPrimitives/distributions.jl; the same table as jbugs/gq.py)
__device__ double gq_sample(unsigned fam, double a0, double a1, double a2, GqRng& g) {
    switch (fam) {
    case JBUGS_FAM_NORMAL: return a0 + g.normal() / sqrt(a1);
    case JBUGS_FAM_GAMMA: return g.gamma(a0) / a1;
    case JBUGS_FAM_EXPONENTIAL: return g.exponential() / a0;
    case JBUGS_FAM_UNIFORM: return a0 + (a1 - a0) * g.uniform();
    case JBUGS_FAM_BETA: { const double x = g.gamma(a0), y = g.gamma(a1); return x / (x + y); }
    case JBUGS_FAM_LOGNORMAL: return exp(a0 + g.normal() / sqrt(a1));
    case JBUGS_FAM_BERNOULLI: return g.uniform() < a0 ? 1.0 : 0.0;
    case JBUGS_FAM_BINOMIAL: return g.binomial(a1, a0);
    case JBUGS_FAM_POISSON: return g.poisson(a0);
    case JBUGS_FAM_LOGISTIC: { const double u = g.uniform(); return a0 + log(u / (1.0 - u)) / sqrt(a1); }
    case JBUGS_FAM_LAPLACE: { const double u = g.uniform() - 0.5; return a0 - (u < 0.0 ? -1.0 : 1.0) * log1p(-2.0 * fabs(u)) / sqrt(a1); }
    case JBUGS_FAM_WEIBULL: return pow(g.exponential(), 1.0 / a0) / a1;
    case JBUGS_FAM_NEGBIN: return g.poisson(g.gamma(a1) * (1.0 - a0) / a0);  // gamma-Poisson mixture: NegativeBinomial(r, p)
    case JBUGS_FAM_CHISQ: return 2.0 * g.gamma(0.5 * a0);
    case JBUGS_FAM_T: return a0 + g.normal() / sqrt(2.0 * g.gamma(0.5 * a2) / a2) / sqrt(a1);
    case JBUGS_FAM_BETABIN: { const double x = g.gamma(a0), y = g.gamma(a1); return g.binomial(a2, x / (x + y)); }
    case JBUGS_FAM_GEOMETRIC: return a0 >= 1.0 ? 0.0 : floor(log(g.uniform()) / log1p(-a0));  // failures before the first success
    default: return dev_nan();
    }
}

__global__ void __launch_bounds__(BLOCK)
gq_stmt_kernel(const GqParams P, const double* __restrict__ theta, long long ld, long long S, double* __restrict__ Q, long long ldq) {
    const long long s = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (s >= S) return;
    const long long c0 = (long long)blockIdx.y * P.cells_per_tile;
    const long long c1 = c0 + P.cells_per_tile < P.n_cells ? c0 + P.cells_per_tile : P.n_cells;
    double v[JBUGS_GQ_MAX_OPS];
    for (long long cell = c0; cell < c1; ++cell) {
        for (int q = 0; q < P.n_ops; ++q) {
            const jbugs_gq_op_rec o = P.ops[q];
            if (o.op == JBUGS_OP_LEAF) {
                long long row;
                switch (o.leaf_kind) {
                case JBUGS_GQ_CONST: v[q] = o.value; break;
                case JBUGS_GQ_THETA:
                case JBUGS_GQ_THETA_CELL:
                    row = o.leaf_kind == JBUGS_GQ_THETA ? o.offset : P.ipool[o.offset + cell];
                    v[q] = transform_eval(P.tr[row], P.lb[row], P.ub[row], theta[(long long)P.in_row[row] * ld + s]).x;
                    break;
                case JBUGS_GQ_Q:
                case JBUGS_GQ_Q_CELL:
                    row = o.leaf_kind == JBUGS_GQ_Q ? o.offset : P.ipool[o.offset + cell];
                    v[q] = Q[row * ldq + s];
                    break;
                case JBUGS_GQ_DATA_CELL: v[q] = P.dpool[o.offset + cell]; break;
                default: v[q] = dev_nan(); break;
                }
            } else if (o.op < 16 || (o.op >= 40 && o.op < 60)) {
                v[q] = gq_unary(o.op, v[o.a]);
            } else {
                v[q] = gq_binary(o.op, v[o.a], v[o.b]);
            }
        }
        const long long row = P.out_row0 + cell;
        double r;
        if (P.family == JBUGS_GQ_LOGICAL) {
            r = v[P.n_ops - 1];
        } else {
            GqRng g(P.seed, (unsigned long long)row, (unsigned long long)s);
            r = gq_sample(P.family, P.n_args > 0 ? v[P.arg_slot[0]] : 0.0, P.n_args > 1 ? v[P.arg_slot[1]] : 0.0,
                          P.n_args > 2 ? v[P.arg_slot[2]] : 0.0, g);
        }
        Q[row * ldq + s] = r;
    }
}

}  // namespace jbugs
