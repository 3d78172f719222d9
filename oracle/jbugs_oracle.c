/*
 * TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
 *
 * jbugs_oracle.c : scalar CPU restatement of the reference's *generated* evaluator
 * `__compute_log_density__` (JuliaBUGS/src/source_gen.jl:670-774): real `for` loops over the
 * plates, one theta at a time, Float64 throughout, `logpdf` per statement instance
 * (`:738-741` observe, `:747-774` parameters incl. inverse bijector + logabsdetjac), same
 * DomainError -> (-Inf, NaN) mapping as JuliaBUGS/src/model/logdensityproblems.jl:22-26,226-229.
 * The gradient is the analytic reverse sweep of that same straight-line program (what the
 * reference obtains from its AD backend, logdensityproblems.jl:219-232).
 *
 * The arithmetic follows the third-party packages the reference delegates to (NOT vendored under
 * /root/reference, only compat ranges in JuliaBUGS/Project.toml:53-84): Distributions 0.25 /
 * StatsFuns 1.x (`normlogpdf`, `gammalogpdf`, `betalogpdf`, `binomlogpdf`, `poislogpdf`),
 * LogExpFunctions 0.3 (`logistic` incl. its saturation bounds, `cexpexp`), Bijectors 0.13-0.16
 * (support-based log / scaled-logit maps).
 *
 * It interprets the same flat plan (include/jbugs_plan.h) the CUDA library consumes; the
 * plan/lowering itself is validated against the per-node graph oracle (oracle/graph_oracle.py),
 * which is pinned to the reference's golden constants (tests/test_oracle_goldens.py).
 *
 * Used as: the checker in tests/, smoke(), and the `cpu_baseline` / `--impl reference` leg of
 * bench.py (OpenMP over chains; one chain per thread exactly like MCMCThreads runs the reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/jbugs_plan.h"
#include "../include/jbugs_trunc.h"

#define LOG2PI 1.8378770664093454835606594728112
#define LOGISTIC_LOWER (-744.4400719213812)
#define LOGISTIC_UPPER (36.7368005696771)

typedef struct jbo_plan {
    jbugs_plan_header h;
    jbugs_global* globals;
    jbugs_gop* gtape;
    jbugs_data_desc* data;
    jbugs_plate* plates;
    jbugs_local* locals;
    jbugs_term* terms;
    jbugs_dense* dense;
    jbugs_xop* xops;
    const void** slots; /* borrowed pointers */
    int64_t* shard_i0;  /* per plate */
    int64_t* shard_i1;
    int n_gvals;
} jbo_plan;

/* ---------------------------------------------------------------- special functions */

static double digamma(double x) {
    /* asymptotic series after upward recurrence; |err| < 1e-15 for x > 0 */
    double r = 0.0;
    if (x <= 0.0) { /* reflection */
        if (x == floor(x)) return NAN;
        return digamma(1.0 - x) - M_PI / tan(M_PI * x);
    }
    while (x < 10.0) {
        r -= 1.0 / x;
        x += 1.0;
    }
    double f = 1.0 / (x * x);
    double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 +
               f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
    return r + log(x) - 0.5 / x + t;
}

static double logistic(double x) {
    /* LogExpFunctions.logistic for Float64 */
    if (x < LOGISTIC_LOWER) return 0.0;
    if (x > LOGISTIC_UPPER) return 1.0;
    double e = exp(x);
    return e / (1.0 + e);
}

static double normcdf(double x) { return 0.5 * erfc(-x * M_SQRT1_2); }
static double normpdf(double x) { return exp(-0.5 * x * x) / sqrt(2.0 * M_PI); }
static double xlogy(double x, double y) { return x == 0.0 ? 0.0 : x * log(y); }
static double xlog1py(double x, double y) { return x == 0.0 ? 0.0 : x * log1p(y); }
static double logbeta(double a, double b) { return lgamma(a) + lgamma(b) - lgamma(a + b); }

/* unary op: value and derivative */
static int unary(uint32_t op, double x, double* v, double* d) {
    switch (op) {
    case JBUGS_OP_IDENTITY: *v = x; *d = 1.0; return 0;
    case JBUGS_OP_EXP: *v = exp(x); *d = *v; return 0;
    case JBUGS_OP_LOGISTIC: *v = logistic(x); *d = *v * (1.0 - *v); return 0;
    case JBUGS_OP_CEXPEXP: *v = -expm1(-exp(x)); *d = exp(x - exp(x)); return 0;
    case JBUGS_OP_PHI: *v = normcdf(x); *d = normpdf(x); return 0;
    case JBUGS_OP_LOG: if (x < 0) return 1; *v = log(x); *d = 1.0 / x; return 0;
    case JBUGS_OP_SQRT: if (x < 0) return 1; *v = sqrt(x); *d = 0.5 / *v; return 0;
    case JBUGS_OP_NEG: *v = -x; *d = -1.0; return 0;
    default: return 2;
    }
}

/* log density of `fam` at x with arguments a[0..2]; partials dx, da[]; returns 1 on DomainError */
static int fam_logpdf(uint32_t fam, double x, const double* a, double* lp, double* dx, double* da) {
    da[0] = da[1] = da[2] = 0.0;
    *dx = 0.0;
    switch (fam) {
    case JBUGS_FAM_NORMAL: { /* distributions.jl:11-18 */
        double mu = a[0], tau = a[1];
        if (tau < 0.0) return 1;
        double sigma = sqrt(1.0 / tau);
        double z = (x - mu) / sigma;
        *lp = -(z * z + LOG2PI) / 2.0 - log(sigma);
        double r = x - mu;
        *dx = -tau * r;
        da[0] = tau * r;
        da[1] = 0.5 / tau - 0.5 * r * r;
        return 0;
    }
    case JBUGS_FAM_GAMMA: { /* distributions.jl:200-203 -> Gamma(a, 1/b) */
        double k = a[0], b = a[1];
        double theta = 1.0 / b;
        if (!(k > 0.0) || !(theta > 0.0)) return 1;
        if (x < 0.0) { *lp = -INFINITY; return 0; }
        double xt = x / theta;
        *lp = -lgamma(k) + xlogy(k - 1.0, xt) - log(theta) - xt;
        *dx = (k - 1.0) / x - b;
        da[0] = -digamma(k) + log(xt);
        da[1] = k / b - x;
        return 0;
    }
    case JBUGS_FAM_EXPONENTIAL: { /* :186-188 -> Exponential(1/lambda) */
        double theta = 1.0 / a[0];
        if (!(theta > 0.0)) return 1;
        if (x < 0.0) { *lp = -INFINITY; return 0; }
        double rate = 1.0 / theta;
        *lp = log(rate) - rate * x;
        *dx = -rate;
        da[0] = 1.0 / a[0] - x;
        return 0;
    }
    case JBUGS_FAM_UNIFORM: {
        if (!(a[0] < a[1])) return 1;
        if (x < a[0] || x > a[1]) { *lp = -INFINITY; return 0; }
        *lp = -log(a[1] - a[0]);
        da[0] = 1.0 / (a[1] - a[0]);
        da[1] = -1.0 / (a[1] - a[0]);
        return 0;
    }
    case JBUGS_FAM_BETA: {
        if (!(a[0] > 0.0) || !(a[1] > 0.0)) return 1;
        if (x < 0.0 || x > 1.0) { *lp = -INFINITY; return 0; }
        *lp = xlogy(a[0] - 1.0, x) + xlog1py(a[1] - 1.0, -x) - logbeta(a[0], a[1]);
        *dx = (a[0] - 1.0) / x - (a[1] - 1.0) / (1.0 - x);
        double pab = digamma(a[0] + a[1]);
        da[0] = log(x) - digamma(a[0]) + pab;
        da[1] = log1p(-x) - digamma(a[1]) + pab;
        return 0;
    }
    case JBUGS_FAM_LOGNORMAL: {
        double mu = a[0], tau = a[1];
        if (tau < 0.0) return 1;
        double sigma = sqrt(1.0 / tau);
        if (!(sigma > 0.0)) return 1;
        if (x <= 0.0) { *lp = -INFINITY; return 0; }
        double lx = log(x), z = (lx - mu) / sigma, r = lx - mu;
        *lp = -(z * z + LOG2PI) / 2.0 - log(sigma) - lx;
        *dx = -(tau * r + 1.0) / x;
        da[0] = tau * r;
        da[1] = 0.5 / tau - 0.5 * r * r;
        return 0;
    }
    case JBUGS_FAM_BERNOULLI: {
        double p = a[0];
        if (!(p >= 0.0 && p <= 1.0)) return 1;
        if (x == 0.0) { *lp = log(1.0 - p); da[0] = -1.0 / (1.0 - p); }
        else if (x == 1.0) { *lp = log(p); da[0] = 1.0 / p; }
        else *lp = -INFINITY;
        return 0;
    }
    case JBUGS_FAM_BINOMIAL: { /* :458-460 -> Binomial(n, p) */
        double p = a[0], n = a[1];
        if (!(n >= 0.0) || !(p >= 0.0 && p <= 1.0)) return 1;
        if (x < 0.0 || x > n || x != floor(x)) { *lp = -INFINITY; return 0; }
        /* StatsFuns.binomlogpdf = betalogpdf(k+1, n-k+1, p) - log(n+1) */
        *lp = xlogy(x, p) + xlog1py(n - x, -p) - logbeta(x + 1.0, n - x + 1.0) - log(n + 1.0);
        da[0] = (x == 0.0 ? 0.0 : x / p) - (n - x == 0.0 ? 0.0 : (n - x) / (1.0 - p));
        return 0;
    }
    case JBUGS_FAM_POISSON: {
        double lam = a[0];
        if (!(lam >= 0.0)) return 1;
        if (x < 0.0 || x != floor(x)) { *lp = -INFINITY; return 0; }
        *lp = xlogy(x, lam) - lam - lgamma(x + 1.0);
        da[0] = (x == 0.0 ? 0.0 : x / lam) - 1.0;
        return 0;
    }
    case JBUGS_FAM_FLAT: *lp = 0.0; return 0;
    case JBUGS_FAM_LOGISTIC: { /* :30-33 -> Logistic(mu, 1/sqrt(tau)); logpdf = u - 2 log1pexp(u) - log(s), u = -|z| */
        double mu = a[0], tau = a[1];
        if (tau < 0.0) return 1;
        double s = 1.0 / sqrt(tau);
        if (!(s > 0.0)) return 1;
        double z = (x - mu) / s, u = -fabs(z);
        *lp = u - 2.0 * log1p(exp(u)) - log(s);
        double w = 1.0 - 2.0 * logistic(z); /* d/dz = -tanh(z/2) */
        *dx = w / s;
        da[0] = -w / s;
        da[1] = w * (x - mu) / (2.0 * sqrt(tau)) + 0.5 / tau;
        return 0;
    }
    case JBUGS_FAM_LAPLACE: { /* :92-95 -> Laplace(mu, 1/sqrt(tau)); logpdf = -(|x - mu| / b + log(2 b)) */
        double mu = a[0], tau = a[1];
        if (tau < 0.0) return 1;
        double b = 1.0 / sqrt(tau);
        if (!(b > 0.0)) return 1;
        double r = x - mu, sg = r >= 0.0 ? 1.0 : -1.0;
        *lp = -(fabs(r) / b + log(2.0 * b));
        *dx = -sg / b;
        da[0] = sg / b;
        da[1] = -fabs(r) / (2.0 * sqrt(tau)) + 0.5 / tau;
        return 0;
    }
    case JBUGS_FAM_WEIBULL: { /* :233-235 -> Weibull(a, 1/b); logpdf = log(a/theta) + xlogy(a-1, z) - z^a, z = x/theta */
        double al = a[0], b = a[1], theta = 1.0 / b;
        if (!(al > 0.0) || !(theta > 0.0)) return 1;
        if (x < 0.0) { *lp = -INFINITY; return 0; }
        double z = x / theta, lz = log(z), za = exp(al * lz);
        *lp = log(al / theta) + xlogy(al - 1.0, z) - za;
        *dx = ((al - 1.0) - al * za) / x;
        da[0] = 1.0 / al + lz - za * lz;
        da[1] = al * (1.0 - za) / b;
        return 0;
    }
    case JBUGS_FAM_NEGBIN: { /* :516-518 -> NegativeBinomial(r, p) */
        double p = a[0], r = a[1];
        if (!(r > 0.0) || !(p > 0.0 && p <= 1.0)) return 1;
        if (x < 0.0 || x != floor(x)) { *lp = -INFINITY; return 0; }
        *lp = xlogy(r, p) + xlog1py(x, -p) - log(x + r) - logbeta(r, x + 1.0);
        da[0] = r / p - (x == 0.0 ? 0.0 : x / (1.0 - p));
        da[1] = log(p) - 1.0 / (x + r) - digamma(r) + digamma(r + x + 1.0);
        return 0;
    }
    case JBUGS_FAM_CHISQ: { /* :215-217 -> Chisq(k) = Gamma(k/2, 2) */
        double h = 0.5 * a[0];
        if (!(a[0] > 0.0)) return 1;
        if (x < 0.0) { *lp = -INFINITY; return 0; }
        *lp = -lgamma(h) - h * M_LN2 + xlogy(h - 1.0, x) - 0.5 * x;
        *dx = (h - 1.0) / x - 0.5;
        da[0] = 0.5 * (-digamma(h) - M_LN2 + log(x));
        return 0;
    }
    case JBUGS_FAM_T: { /* :73-80 -> TDistShiftedScaled(nu, mu, 1/sqrt(tau)): StatsFuns.tdistlogpdf(nu, z) - log(sigma) */
        double mu = a[0], tau = a[1], nu = a[2];
        if (tau < 0.0) return 1;
        double sigma = sqrt(1.0 / tau);
        if (!(nu > 0.0) || !(sigma > 0.0)) return 1;
        double z = (x - mu) / sigma;
        *lp = lgamma((nu + 1.0) / 2.0) - log(nu * M_PI) / 2.0 - lgamma(nu / 2.0) - (nu + 1.0) / 2.0 * log1p(z * z / nu) - log(sigma);
        double s = (nu + 1.0) * z / (nu + z * z);
        *dx = -s / sigma;
        da[0] = s / sigma;
        da[1] = (0.5 / tau) * (1.0 - s * z);
        da[2] = 0.5 * digamma((nu + 1.0) / 2.0) - 0.5 / nu - 0.5 * digamma(nu / 2.0) - 0.5 * log1p(z * z / nu)
                + (nu + 1.0) / 2.0 * (z * z / (nu * nu)) / (1.0 + z * z / nu);
        return 0;
    }
    case JBUGS_FAM_GEOMETRIC: { /* :500-502 -> Geometric(p): log(p) + xlog1py(x, -p), x = 0, 1, 2, ... */
        double p = a[0];
        if (!(p > 0.0 && p <= 1.0)) return 1;
        if (x < 0.0 || x != floor(x)) { *lp = -INFINITY; return 0; }
        *lp = log(p) + xlog1py(x, -p);
        da[0] = 1.0 / p - (x == 0.0 ? 0.0 : x / (1.0 - p));
        return 0;
    }
    case JBUGS_FAM_CATEGORICAL: { /* distributions.jl:472-474 -> Categorical(p): logpdf = log p[x]; a[0] = p[x].
                                     (the whole-vector check sum(p) == 1 is not part of the plan: the other components
                                     never enter the density) */
        double py = a[0];
        if (!(py >= 0.0 && py <= 1.0 + 1e-8)) return 1;
        *lp = log(py);
        da[0] = 1.0 / py;
        return 0;
    }
    case JBUGS_FAM_BETABIN: { /* :530-532 -> BetaBinomial(n, a, b): -log1p(n) - logbeta(k+1, n-k+1) + logbeta(k+a, n-k+b) - logbeta(a, b) */
        double al = a[0], be = a[1], n = a[2];
        if (!(n >= 0.0) || !(al >= 0.0) || !(be >= 0.0)) return 1;
        if (x < 0.0 || x > n || x != floor(x)) { *lp = -INFINITY; return 0; }
        *lp = -log1p(n) - logbeta(x + 1.0, n - x + 1.0) + logbeta(x + al, n - x + be) - logbeta(al, be);
        da[0] = digamma(x + al) - digamma(n + al + be) - digamma(al) + digamma(al + be);
        da[1] = digamma(n - x + be) - digamma(n + al + be) - digamma(be) + digamma(al + be);
        return 0;
    }
    default: return 2;
    }
}

/* slot y -> constrained value x, logjac, dx/dy, dlogjac/dy (Bijectors, evaluation.jl:243-251) */
static void transform(uint32_t tr, double lb, double ub, double y, double* x, double* lj, double* dxdy,
                      double* dljdy) {
    switch (tr) {
    case JBUGS_TR_LOG: { double e = exp(y); *x = lb + e; *lj = y; *dxdy = e; *dljdy = 1.0; return; }
    case JBUGS_TR_UPPER: { double e = exp(y); *x = ub - e; *lj = y; *dxdy = -e; *dljdy = 1.0; return; }
    case JBUGS_TR_LOGIT: {
        double s = logistic(y), w = ub - lb;
        *x = lb + w * s;
        *lj = log((*x - lb) * (ub - *x) / w);
        *dxdy = w * s * (1.0 - s);
        *dljdy = 1.0 - 2.0 * s;
        return;
    }
    default: *x = y; *lj = 0.0; *dxdy = 1.0; *dljdy = 0.0; return;
    }
}

/* ---------------------------------------------------------------- plan handling */

jbo_plan* jbo_plan_create(const void* blob, size_t nbytes) {
    if (nbytes < sizeof(jbugs_plan_header)) return NULL;
    jbo_plan* p = (jbo_plan*)calloc(1, sizeof(jbo_plan));
    memcpy(&p->h, blob, sizeof(jbugs_plan_header));
    if (memcmp(p->h.magic, JBUGS_PLAN_MAGIC, 8) != 0 || p->h.version != JBUGS_PLAN_VERSION) { free(p); return NULL; }
    size_t need = sizeof(jbugs_plan_header) + p->h.n_globals * sizeof(jbugs_global) + p->h.n_gtape * sizeof(jbugs_gop) +
                  p->h.n_data * sizeof(jbugs_data_desc) + p->h.n_plates * sizeof(jbugs_plate) +
                  p->h.n_locals * sizeof(jbugs_local) + p->h.n_terms * sizeof(jbugs_term) +
                  p->h.n_dense * sizeof(jbugs_dense) + p->h.n_xops * sizeof(jbugs_xop);
    if (need != nbytes) { free(p); return NULL; }
    const char* c = (const char*)blob + sizeof(jbugs_plan_header);
#define TAKE(field, type, n) \
    p->field = (type*)malloc((n) * sizeof(type) + 1); memcpy(p->field, c, (n) * sizeof(type)); c += (n) * sizeof(type)
    TAKE(globals, jbugs_global, p->h.n_globals);
    TAKE(gtape, jbugs_gop, p->h.n_gtape);
    TAKE(data, jbugs_data_desc, p->h.n_data);
    TAKE(plates, jbugs_plate, p->h.n_plates);
    TAKE(locals, jbugs_local, p->h.n_locals);
    TAKE(terms, jbugs_term, p->h.n_terms);
    TAKE(dense, jbugs_dense, p->h.n_dense);
    TAKE(xops, jbugs_xop, p->h.n_xops);
#undef TAKE
    p->slots = (const void**)calloc(p->h.n_data + 1, sizeof(void*));
    p->shard_i0 = (int64_t*)calloc(p->h.n_plates + 1, sizeof(int64_t));
    p->shard_i1 = (int64_t*)calloc(p->h.n_plates + 1, sizeof(int64_t));
    for (uint32_t k = 0; k < p->h.n_plates; ++k) p->shard_i1[k] = p->plates[k].N;
    p->n_gvals = (int)(p->h.n_globals + p->h.n_gtape);
    return p;
}

void jbo_plan_destroy(jbo_plan* p) {
    if (!p) return;
    free(p->globals); free(p->gtape); free(p->data); free(p->plates); free(p->locals); free(p->terms); free(p->dense); free(p->xops);
    free(p->slots); free(p->shard_i0); free(p->shard_i1); free(p);
}

int64_t jbo_dim(const jbo_plan* p) { return p->h.D; }

/* borrowed pointer: the caller keeps the array alive */
int jbo_plan_set_data(jbo_plan* p, int slot, const void* ptr, size_t nbytes) {
    if (slot < 0 || (uint32_t)slot >= p->h.n_data) return -1;
    if (nbytes != (size_t)p->data[slot].len * 8) return -1;
    p->slots[slot] = ptr;
    return 0;
}

int jbo_plan_set_shard(jbo_plan* p, int plate, int64_t i0, int64_t i1) {
    if (plate < 0 || (uint32_t)plate >= p->h.n_plates || i0 < 0 || i1 > p->plates[plate].N || i0 > i1) return -1;
    p->shard_i0[plate] = i0;
    p->shard_i1[plate] = i1;
    return 0;
}

/* ---------------------------------------------------------------- one chain */

#define MAXG 256
#define MAXL 16

typedef struct {
    const double* th; int64_t sth; /* theta[d * sth] */
    double* g; int64_t sg;         /* grad[d * sg]   */
} chain_io;

static inline double argval(const jbo_plan* p, const jbugs_arg* a, const double* gval, const double* lx,
                            int64_t i, int64_t j, int64_t N) {
    switch (a->kind) {
    case JBUGS_ARG_CONST: return a->value;
    case JBUGS_ARG_ONE: return 1.0;
    case JBUGS_ARG_GVAL: return gval[a->id];
    case JBUGS_ARG_GVEC_J: return gval[a->id + j];
    case JBUGS_ARG_LOCAL: return lx[a->id];
    case JBUGS_ARG_DATA_I: return ((const double*)p->slots[a->id])[i];
    case JBUGS_ARG_DATA_J: return ((const double*)p->slots[a->id])[j];
    case JBUGS_ARG_DATA_IJ: return ((const double*)p->slots[a->id])[i + N * j];
    default: return NAN;
    }
}

static inline void argadj(const jbugs_arg* a, double d, double* gadj, double* ladj) {
    if (a->kind == JBUGS_ARG_GVAL) gadj[a->id] += d;
    else if (a->kind == JBUGS_ARG_LOCAL) ladj[a->id] += d;
}

static inline int64_t local_slot(const jbo_plan* p, const jbugs_local* l, int64_t i, int64_t j, int64_t N) {
    if (l->idx_slot >= 0) {
        const int64_t* ix = (const int64_t*)p->slots[l->idx_slot];
        return l->level == JBUGS_LEVEL_OBS ? ix[i + N * j] : ix[i];
    }
    return l->theta_base + l->stride_i * i + l->stride_j * j;
}

/* returns 0 ok, 1 domain error */
static int eval_chain(const jbo_plan* p, chain_io io, int want_grad, double* logp_out) {
    double gval[MAXG], gadj[MAXG], gdxdy[MAXG], gdlj[MAXG];
    const int H = (int)p->h.n_globals, NG = p->n_gvals;
    double logp = 0.0;
    for (int k = 0; k < NG; ++k) gadj[k] = 0.0;
    /* globals: inverse bijector + logjac (source_gen.jl:747-768) */
    for (int h = 0; h < H; ++h) {
        const jbugs_global* g = &p->globals[h];
        double lj;
        transform(g->transform, g->lb, g->ub, io.th[g->theta_idx * io.sth], &gval[h], &lj, &gdxdy[h], &gdlj[h]);
        logp += lj;
    }
    /* scalar logical nodes of globals */
    double tv_d1[MAXG], tv_d2[MAXG];
    for (uint32_t k = 0; k < p->h.n_gtape; ++k) {
        const jbugs_gop* o = &p->gtape[k];
        double a = argval(p, &o->a, gval, NULL, 0, 0, 0), v, d1 = 0, d2 = 0;
        if (o->op < 16) {
            if (unary(o->op, a, &v, &d1)) return 1;
        } else {
            double b = argval(p, &o->b, gval, NULL, 0, 0, 0);
            switch (o->op) {
            case JBUGS_OP_ADD: v = a + b; d1 = 1; d2 = 1; break;
            case JBUGS_OP_SUB: v = a - b; d1 = 1; d2 = -1; break;
            case JBUGS_OP_MUL: v = a * b; d1 = b; d2 = a; break;
            case JBUGS_OP_DIV: v = a / b; d1 = 1.0 / b; d2 = -a / (b * b); break;
            case JBUGS_OP_POW:
                if (a < 0 && b != floor(b)) return 1;
                v = pow(a, b); d1 = b * pow(a, b - 1.0); d2 = (a > 0) ? v * log(a) : 0.0; break;
            default: return 1;
            }
        }
        gval[H + k] = v; tv_d1[k] = d1; tv_d2[k] = d2;
    }
    /* priors of the globals */
    for (int h = 0; h < H; ++h) {
        const jbugs_global* g = &p->globals[h];
        double a[3], da[3], lp, dx;
        for (int q = 0; q < 3; ++q) a[q] = argval(p, &g->args[q], gval, NULL, 0, 0, 0);
        if (fam_logpdf(g->family, gval[h], a, &lp, &dx, da)) return 1;
        logp += lp;
        gadj[h] += dx;
        for (int q = 0; q < 3; ++q) argadj(&g->args[q], da[q], gadj, NULL);
        /* truncated(dist(args), lb, ub) (the BUGS suffix T(lb, ub)): a global whose support bounds are narrower than
           the natural support of its family (include/jbugs_trunc.h); literal arguments only (checked at plan creation),
           so the normaliser log(cdf(ub) - cdf(lb)) is a constant.  Outside the bounds (possible in untransformed space
           only): -Inf */
        if (jbugs_is_truncated(g->family, a[0], a[1], g->lb, g->ub) == 1) {
            if (gval[h] < g->lb || gval[h] > g->ub) return 1;
            logp -= jbugs_trunc_logmass(g->family, a[0], a[1], a[2], g->lb, g->ub);
        }
    }
    /* plates */
    for (uint32_t pi = 0; pi < p->h.n_plates; ++pi) {
        const jbugs_plate* pl = &p->plates[pi];
        const jbugs_local* L = p->locals + pl->first_local;
        const jbugs_term* Tm = p->terms + pl->first_term;
        const int nl = (int)pl->n_locals, nt = (int)pl->n_terms;
        const int64_t N = pl->N, T = pl->T;
        double lx[MAXL], ladj[MAXL], ldx[MAXL], ldlj[MAXL];
        int64_t lslot[MAXL];
        for (int64_t i = p->shard_i0[pi]; i < p->shard_i1[pi]; ++i) {
            /* group-level locals */
            for (int l = 0; l < nl; ++l) {
                if (L[l].level != JBUGS_LEVEL_GROUP) continue;
                double lj;
                lslot[l] = local_slot(p, &L[l], i, 0, N);
                transform(L[l].transform, L[l].lb, L[l].ub, io.th[lslot[l] * io.sth], &lx[l], &lj, &ldx[l], &ldlj[l]);
                logp += lj;
                ladj[l] = 0.0;
            }
            for (int l = 0; l < nl; ++l) {
                if (L[l].level != JBUGS_LEVEL_GROUP) continue;
                double a[3], da[3], lp, dx;
                for (int q = 0; q < 3; ++q) a[q] = argval(p, &L[l].args[q], gval, lx, i, 0, N);
                if (fam_logpdf(L[l].family, lx[l], a, &lp, &dx, da)) return 1;
                logp += lp;
                ladj[l] += dx;
                for (int q = 0; q < 3; ++q) argadj(&L[l].args[q], da[q], gadj, ladj);
            }
            for (int64_t j = 0; j < T; ++j) {
                for (int l = 0; l < nl; ++l) {
                    if (L[l].level != JBUGS_LEVEL_OBS) continue;
                    double lj;
                    lslot[l] = local_slot(p, &L[l], i, j, N);
                    transform(L[l].transform, L[l].lb, L[l].ub, io.th[lslot[l] * io.sth], &lx[l], &lj, &ldx[l], &ldlj[l]);
                    logp += lj;
                    ladj[l] = 0.0;
                }
                for (int l = 0; l < nl; ++l) {
                    if (L[l].level != JBUGS_LEVEL_OBS) continue;
                    double a[3], da[3], lp, dx;
                    for (int q = 0; q < 3; ++q) a[q] = argval(p, &L[l].args[q], gval, lx, i, j, N);
                    if (fam_logpdf(L[l].family, lx[l], a, &lp, &dx, da)) return 1;
                    logp += lp;
                    ladj[l] += dx;
                    for (int q = 0; q < 3; ++q) argadj(&L[l].args[q], da[q], gadj, ladj);
                }
                const int has_w = pl->aux[2].kind == JBUGS_ARG_DATA_I || pl->aux[2].kind == JBUGS_ARG_DATA_IJ;
                if (pl->has_obs >= 2 && !(has_w && argval(p, &pl->aux[2], gval, lx, i, j, N) == 0.0)) {
                    /* expression plate: forward sweep over the tape (the composed node functions of the observation's
                       logical parents), observation, reverse sweep back to the leaves.  has_obs == 3: a finite discrete
                       latent per observation is summed out (evaluation.jl:739-961 restricted to one latent per plate
                       cell): the cell contributes logsumexp_k [log p(z = k) + log p(y | z = k)], its gradient is the
                       posterior-weighted sum of the branch gradients */
                    const jbugs_xop* X = p->xops + pl->first_term;
                    const int nx = (int)pl->n_terms;
                    const int K = pl->has_obs == 3 ? (int)pl->eta_arg : 1;
                    const int lps = pl->has_obs == 3 ? (int)pl->link[0] : -1;
                    double v[JBUGS_MAX_XOPS], w[JBUGS_MAX_XOPS], d1[JBUGS_MAX_XOPS], d2[JBUGS_MAX_XOPS];
                    double branch[JBUGS_MAX_STATES];
                    if (K < 1 || K > JBUGS_MAX_STATES) return 1;
                    double y = argval(p, &pl->y, gval, lx, i, j, N);
                    double lse = 0.0;
                    /* a partially observed latent: link[1] - 1 is the tape slot of a data leaf with the observed state
                       of the cell (0 = missing); an observed cell takes the branch of its value only */
                    const int zs = pl->has_obs == 3 ? (int)pl->link[1] - 1 : -1;
                    const int zo = zs >= 0 ? (int)argval(p, &X[zs].leaf, gval, lx, i, j, N) : 0;
                    for (int pass = 0; pass < 2; ++pass) {
                        for (int k = 0; k < K; ++k) {
                            if (zo > 0 && k != zo - 1) {
                                if (pass == 0) branch[k] = -INFINITY;
                                continue;
                            }
                            for (int q = 0; q < nx; ++q) {
                                const jbugs_xop* o = &X[q];
                                w[q] = 0.0; d1[q] = 0.0; d2[q] = 0.0;
                                if (o->op == JBUGS_OP_LEAF) { v[q] = argval(p, &o->leaf, gval, lx, i, j, N); continue; }
                                if (o->op == JBUGS_OP_SELECT) { v[q] = v[o->a] != 0.0 ? v[o->b] : v[o->c]; continue; }
                                if (o->op == JBUGS_OP_PICK_K) { v[q] = v[o->a + k]; continue; }
                                if (o->op == JBUGS_OP_GVEC_AT) {
                                    const int e = (int)v[o->a];
                                    if (e < 0 || e >= (int)o->leaf.value) return 1;  /* an index outside the vector */
                                    v[q] = gval[o->leaf.id + e];
                                    continue;
                                }
                                const double a = v[o->a];
                                if (o->op < 16) {
                                    if (unary(o->op, a, &v[q], &d1[q])) return 1;
                                    continue;
                                }
                                const double b = v[o->b];
                                switch (o->op) {
                                case JBUGS_OP_ADD: v[q] = a + b; d1[q] = 1; d2[q] = 1; break;
                                case JBUGS_OP_SUB: v[q] = a - b; d1[q] = 1; d2[q] = -1; break;
                                case JBUGS_OP_MUL: v[q] = a * b; d1[q] = b; d2[q] = a; break;
                                case JBUGS_OP_DIV: v[q] = a / b; d1[q] = 1.0 / b; d2[q] = -a / (b * b); break;
                                case JBUGS_OP_POW:
                                    if (a < 0 && b != floor(b)) return 1;
                                    v[q] = pow(a, b); d1[q] = b * pow(a, b - 1.0); d2[q] = (a > 0) ? v[q] * log(a) : 0.0; break;
                                default: return 1;
                                }
                            }
                            double a[3], da[3], lp, dx;
                            for (int q = 0; q < 3; ++q)
                                a[q] = pl->aux[q].kind == JBUGS_ARG_XVAL ? v[pl->aux[q].id] : argval(p, &pl->aux[q], gval, lx, i, j, N);
                            if (fam_logpdf(pl->family, y, a, &lp, &dx, da)) return 1;
                            if (pass == 0) {
                                branch[k] = lp + (lps >= 0 ? v[lps] : 0.0);
                                continue;
                            }
                            /* pass 1: reverse sweep of branch k, scaled by its posterior weight */
                            const double r = K == 1 ? 1.0 : exp(branch[k] - lse);
                            if (r == 0.0) continue;
                            if (lps >= 0) w[lps] += r;
                            for (int q = 0; q < 3; ++q) {
                                if (pl->aux[q].kind == JBUGS_ARG_XVAL) w[pl->aux[q].id] += r * da[q];
                                else if (q < 2) argadj(&pl->aux[q], r * da[q], gadj, ladj);
                            }
                            for (int q = nx - 1; q >= 0; --q) {
                                const jbugs_xop* o = &X[q];
                                if (w[q] == 0.0) continue;
                                if (o->op == JBUGS_OP_LEAF) {
                                    if (o->leaf.kind == JBUGS_ARG_GVEC_J) gadj[o->leaf.id + j] += w[q];
                                    else argadj(&o->leaf, w[q], gadj, ladj);
                                } else if (o->op == JBUGS_OP_SELECT) {
                                    w[v[o->a] != 0.0 ? o->b : o->c] += w[q];
                                } else if (o->op == JBUGS_OP_PICK_K) {
                                    w[o->a + k] += w[q];
                                } else if (o->op == JBUGS_OP_GVEC_AT) {
                                    gadj[o->leaf.id + (int)v[o->a]] += w[q];
                                } else {
                                    w[o->a] += w[q] * d1[q];
                                    if (o->op >= 16) w[o->b] += w[q] * d2[q];
                                }
                            }
                        }
                        if (pass == 0) {
                            double mx = branch[0];
                            for (int k = 1; k < K; ++k) mx = branch[k] > mx ? branch[k] : mx;
                            if (K == 1 || !(mx > -INFINITY)) lse = mx;
                            else {
                                double sum = 0.0;
                                for (int k = 0; k < K; ++k) sum += exp(branch[k] - mx);
                                lse = mx + log(sum);
                            }
                            logp += lse;
                            if (!(lse > -INFINITY)) break; /* no branch has support: nothing to differentiate */
                        }
                    }
                } else if (pl->has_obs && !(has_w && argval(p, &pl->aux[2], gval, lx, i, j, N) == 0.0)) {
                    /* linear predictor, link tape, observation (source_gen.jl:738-741); cells with weight 0 are the
                       padding of a plate whose observations were gathered by a data index */
                    double eta = 0.0;
                    for (int t = 0; t < nt; ++t)
                        eta += argval(p, &Tm[t].coef, gval, lx, i, j, N) * argval(p, &Tm[t].cov, gval, lx, i, j, N);
                    double v = eta, dv = 1.0;
                    for (uint32_t q = 0; q < pl->n_link; ++q) {
                        double nv, d;
                        if (unary(pl->link[q], v, &nv, &d)) return 1;
                        v = nv; dv *= d;
                    }
                    double a[3], da[3], lp, dx;
                    for (int q = 0; q < 3; ++q)
                        a[q] = (q == (int)pl->eta_arg) ? v : argval(p, &pl->aux[q], gval, lx, i, j, N);
                    double y = argval(p, &pl->y, gval, lx, i, j, N);
                    if (fam_logpdf(pl->family, y, a, &lp, &dx, da)) return 1;
                    logp += lp;
                    double deta = da[pl->eta_arg] * dv;
                    for (int q = 0; q < 3; ++q)
                        if (q != (int)pl->eta_arg) argadj(&pl->aux[q], da[q], gadj, ladj);
                    for (int t = 0; t < nt; ++t)
                        argadj(&Tm[t].coef, deta * argval(p, &Tm[t].cov, gval, lx, i, j, N), gadj, ladj);
                }
                if (want_grad)
                    for (int l = 0; l < nl; ++l)
                        if (L[l].level == JBUGS_LEVEL_OBS) io.g[lslot[l] * io.sg] = ladj[l] * ldx[l] + ldlj[l];
            }
            if (want_grad)
                for (int l = 0; l < nl; ++l)
                    if (L[l].level == JBUGS_LEVEL_GROUP) io.g[lslot[l] * io.sg] = ladj[l] * ldx[l] + ldlj[l];
        }
    }
    /* dense predictors: eta_i = inprod(X[i, :], beta) (functions.jl:33-35), one logpdf node per observation */
    for (uint32_t di = 0; di < p->h.n_dense; ++di) {
        const jbugs_dense* dn = &p->dense[di];
        const double* X = (const double*)p->slots[dn->x_slot];
        const double* Y = (const double*)p->slots[dn->y_slot];
        const double* Nn = dn->n_slot >= 0 ? (const double*)p->slots[dn->n_slot] : NULL;
        for (int64_t i = 0; i < dn->N; ++i) {
            double eta = 0.0;
            for (int64_t j = 0; j < dn->P; ++j)
                eta += X[i + dn->N * j] * io.th[(dn->theta_base + dn->theta_stride * j) * io.sth];
            double v, dv, a[3] = {0, 0, 0}, da[3], lp, dx;
            if (unary(dn->link, eta, &v, &dv)) return 1;
            a[0] = v;
            a[1] = Nn ? Nn[i] : 0.0;
            if (fam_logpdf(dn->family, Y[i], a, &lp, &dx, da)) return 1;
            logp += lp;
            if (want_grad) {
                const double deta = da[0] * dv;
                for (int64_t j = 0; j < dn->P; ++j)
                    io.g[(dn->theta_base + dn->theta_stride * j) * io.sg] += deta * X[i + dn->N * j];
            }
        }
    }
    /* reverse sweep over the gtape, then through the bijectors of the globals */
    for (int k = (int)p->h.n_gtape - 1; k >= 0; --k) {
        const jbugs_gop* o = &p->gtape[k];
        double ad = gadj[H + k];
        argadj(&o->a, ad * tv_d1[k], gadj, NULL);
        if (o->op >= 16) argadj(&o->b, ad * tv_d2[k], gadj, NULL);
    }
    if (want_grad)
        for (int h = 0; h < H; ++h) io.g[p->globals[h].theta_idx * io.sg] = gadj[h] * gdxdy[h] + gdlj[h];
    *logp_out = logp;
    return 0;
}

/* theta/grad: layout 0 = CHAIN_FAST (d * ld + c), 1 = PARAM_FAST (c * ld + d) */
int jbo_eval_batch(const jbo_plan* p, const double* theta, int64_t ld, int64_t C, int layout, double* logp,
                   double* grad, int32_t* status, int want_grad, int nthreads) {
    if (p->n_gvals > MAXG) return -2;
    for (uint32_t k = 0; k < p->h.n_plates; ++k)
        if (p->plates[k].n_locals > MAXL) return -2;
    for (uint32_t k = 0; k < p->h.n_data; ++k)
        if (!p->slots[k]) return -5;
    const int64_t D = p->h.D;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int64_t c = 0; c < C; ++c) {
        chain_io io;
        if (layout == 0) { io.th = theta + c; io.sth = ld; io.g = grad ? grad + c : NULL; io.sg = ld; }
        else { io.th = theta + c * ld; io.sth = 1; io.g = grad ? grad + c * ld : NULL; io.sg = 1; }
        double lp = 0.0;
        int st = eval_chain(p, io, want_grad && grad, &lp);
        /* NaN anywhere in theta poisons logp; map like the boundary does */
        if (st) {
            logp[c] = -INFINITY;
            if (want_grad && grad)
                for (int64_t d = 0; d < D; ++d) io.g[d * io.sg] = NAN;
        } else {
            logp[c] = lp;
        }
        if (status) status[c] = st ? 1 : 0;
    }
    return 0;
}

int jbo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
