/*
 * jbugs_trunc.h -- probability mass of a univariate prior with literal arguments on an interval.
 *
 * `x ~ dist(args) T(lower, upper)` (BUGS) = `truncated(dist(args), lower, upper)` (JuliaBUGS,
 * JuliaBUGS/src/parser/bugs_parser.jl:470-500; Distributions.truncated): inside the bounds the log density is the
 * inner one minus log(cdf(upper) - cdf(lower)).  In a plan the bounds travel as the support of the parameter
 * (jbugs_global.lb / ub, which also pick the bijector); a prior whose support is narrower than the natural support of
 * its family is truncated, and -- its arguments being literals -- the normaliser is a constant that the evaluators add
 * once (libjbugs_cuda: folded into the host-side constant of the prior; the C restatement: at evaluation).
 *
 * Plain C, header-only; parameterisations as in jbugs_plan.h (BUGSPrimitives/distributions.jl).
 */
#ifndef JBUGS_TRUNC_H
#define JBUGS_TRUNC_H

#include <math.h>

#include "jbugs_plan.h"

/* regularised incomplete gamma functions P(a, x) and Q(a, x) = 1 - P: series for x < a + 1, Lentz continued fraction
   above (both converge to double precision in a few dozen terms for the shapes priors use) */
static inline void jbugs_gamma_pq(double a, double x, double* p, double* q) {
    if (!(x > 0.0)) { *p = 0.0; *q = 1.0; return; }
    if (isinf(x)) { *p = 1.0; *q = 0.0; return; }
    const double lg = lgamma(a);
    if (x < a + 1.0) {
        double ap = a, del = 1.0 / a, sum = del;
        for (int n = 0; n < 2000; ++n) {
            ap += 1.0; del *= x / ap; sum += del;
            if (fabs(del) < fabs(sum) * 1e-17) break;
        }
        *p = sum * exp(-x + a * log(x) - lg);
        *q = 1.0 - *p;
    } else {
        const double tiny = 1e-300;
        double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
        for (int i = 1; i < 2000; ++i) {
            const double an = -i * (i - a);
            b += 2.0;
            d = an * d + b; if (fabs(d) < tiny) d = tiny;
            c = b + an / c; if (fabs(c) < tiny) c = tiny;
            d = 1.0 / d;
            const double del = d * c;
            h *= del;
            if (fabs(del - 1.0) < 1e-16) break;
        }
        *q = exp(-x + a * log(x) - lg) * h;
        *p = 1.0 - *q;
    }
}

/* regularised incomplete beta function I = I_x(a, b) and its complement J = 1 - I: Lentz continued fraction, entered from
   the side on which it converges fast (x below the mean: I directly; above: J directly through I_{1-x}(b, a)), so the
   smaller of the two is the one computed without cancellation */
static inline double jbugs_betacf(double a, double b, double x) {
    const double tiny = 1e-300, qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m < 5000; ++m) {
        const double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c; if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return h;
}

static inline void jbugs_beta_ij(double a, double b, double x, double* i, double* j) {
    if (!(x > 0.0)) { *i = 0.0; *j = 1.0; return; }
    if (!(x < 1.0)) { *i = 1.0; *j = 0.0; return; }
    const double bt = exp(lgamma(a + b) - lgamma(a) - lgamma(b) + a * log(x) + b * log1p(-x));
    if (x < (a + 1.0) / (a + b + 2.0)) {
        *i = bt * jbugs_betacf(a, b, x) / a;
        *j = 1.0 - *i;
    } else {
        *j = bt * jbugs_betacf(b, a, 1.0 - x) / b;
        *i = 1.0 - *j;
    }
}

/* natural support of a family with these arguments; 0 when the family is one whose truncation is provided */
static inline int jbugs_natural_support(unsigned fam, double a0, double a1, double* lo, double* hi) {
    switch (fam) {
    case JBUGS_FAM_NORMAL: case JBUGS_FAM_LOGISTIC: case JBUGS_FAM_LAPLACE: case JBUGS_FAM_T:
        *lo = -INFINITY; *hi = INFINITY; return 0;
    case JBUGS_FAM_BETA: *lo = 0.0; *hi = 1.0; return 0;
    case JBUGS_FAM_GAMMA: case JBUGS_FAM_EXPONENTIAL: case JBUGS_FAM_LOGNORMAL: case JBUGS_FAM_WEIBULL: case JBUGS_FAM_CHISQ:
        *lo = 0.0; *hi = INFINITY; return 0;
    case JBUGS_FAM_UNIFORM: *lo = a0; *hi = a1; return 0;
    default: return 1;
    }
}

/* lower tail F(x) and upper tail S(x); a2: the third argument of the family (dt's degrees of freedom), else unused */
static inline void jbugs_tails(unsigned fam, double a0, double a1, double a2, double x, double* F, double* S) {
    switch (fam) {
    case JBUGS_FAM_NORMAL: { /* dnorm(mu, tau) */
        const double z = (x - a0) * sqrt(a1) / sqrt(2.0);
        *F = 0.5 * erfc(-z); *S = 0.5 * erfc(z); return;
    }
    case JBUGS_FAM_LOGNORMAL: { /* dlnorm(mu, tau) */
        if (!(x > 0.0)) { *F = 0.0; *S = 1.0; return; }
        const double z = (log(x) - a0) * sqrt(a1) / sqrt(2.0);
        *F = 0.5 * erfc(-z); *S = 0.5 * erfc(z); return;
    }
    case JBUGS_FAM_LOGISTIC: { /* dlogis(mu, tau): scale 1 / sqrt(tau) */
        const double z = (x - a0) * sqrt(a1);
        *F = 1.0 / (1.0 + exp(-z)); *S = 1.0 / (1.0 + exp(z)); return;
    }
    case JBUGS_FAM_LAPLACE: { /* ddexp(mu, tau): scale 1 / sqrt(tau) */
        const double z = (x - a0) * sqrt(a1);
        if (z < 0.0) { *F = 0.5 * exp(z); *S = 1.0 - *F; } else { *S = 0.5 * exp(-z); *F = 1.0 - *S; }
        return;
    }
    case JBUGS_FAM_EXPONENTIAL: /* dexp(lambda) */
        if (!(x > 0.0)) { *F = 0.0; *S = 1.0; return; }
        *S = exp(-a0 * x); *F = -expm1(-a0 * x); return;
    case JBUGS_FAM_WEIBULL: { /* dweib(a, b) = Weibull(shape a, scale 1 / b) */
        if (!(x > 0.0)) { *F = 0.0; *S = 1.0; return; }
        const double t = pow(a1 * x, a0);
        *S = exp(-t); *F = -expm1(-t); return;
    }
    case JBUGS_FAM_GAMMA: jbugs_gamma_pq(a0, a1 * x, F, S); return;        /* dgamma(shape, rate) */
    case JBUGS_FAM_CHISQ: jbugs_gamma_pq(0.5 * a0, 0.5 * x, F, S); return;  /* dchisqr(k) */
    case JBUGS_FAM_BETA: jbugs_beta_ij(a0, a1, x, F, S); return;            /* dbeta(a, b) */
    case JBUGS_FAM_T: { /* dt(mu, tau, nu): P(|T| > |z|) = I_{nu / (nu + z^2)}(nu / 2, 1 / 2), z = (x - mu) sqrt(tau) */
        const double z = (x - a0) * sqrt(a1);
        double i, j;
        if (isinf(z)) { *F = z > 0.0 ? 1.0 : 0.0; *S = 1.0 - *F; return; }
        jbugs_beta_ij(0.5 * a2, 0.5, a2 / (a2 + z * z), &i, &j);
        if (z > 0.0) { *S = 0.5 * i; *F = 1.0 - *S; } else { *F = 0.5 * i; *S = 1.0 - *F; }
        return;
    }
    case JBUGS_FAM_UNIFORM: {
        const double u = x <= a0 ? 0.0 : (x >= a1 ? 1.0 : (x - a0) / (a1 - a0));
        *F = u; *S = 1.0 - u; return;
    }
    default: *F = NAN; *S = NAN; return;
    }
}

/* 1 when [lb, ub] is narrower than the natural support of the family (a truncated prior), 0 when not, -1 when the bounds
   are narrower but the family's truncation is not provided */
static inline int jbugs_is_truncated(unsigned fam, double a0, double a1, double lb, double ub) {
    double lo, hi;
    if (!(lb < ub)) return 0;
    if (jbugs_natural_support(fam, a0, a1, &lo, &hi)) return 0;
    return (lb > lo || ub < hi) ? 1 : 0;
}

/* log(cdf(ub) - cdf(lb)), from the tails on the side where the difference does not cancel */
static inline double jbugs_trunc_logmass(unsigned fam, double a0, double a1, double a2, double lb, double ub) {
    double Fl = 0.0, Sl = 1.0, Fu = 1.0, Su = 0.0;
    if (isfinite(lb)) jbugs_tails(fam, a0, a1, a2, lb, &Fl, &Sl);
    if (isfinite(ub)) jbugs_tails(fam, a0, a1, a2, ub, &Fu, &Su);
    const double mass = Fl > 0.5 ? Sl - Su : Fu - Fl;
    return log(mass);
}

#endif /* JBUGS_TRUNC_H */
