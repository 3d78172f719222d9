// jbugs_device.cuh -- fp64 device math for the plate kernels: distribution families with their
// analytic partials, bijector transforms, inverse-link ops.
//
// What each block replaces in the reference (paths relative to /root/reference/JuliaBUGS):
//   families   src/BUGSPrimitives/distributions.jl (:11-18 dnorm, :186-188 dexp, :200-203 dgamma,
//              :343-345 dunif, :357-359 dbeta, :442-444 dbern, :458-460 dbin, :486-488 dpois)
//              + Distributions.logpdf (third party) + the AD backend's reverse rules
//   transforms Bijectors.with_logabsdet_jacobian at src/model/evaluation.jl:243-251
//   link ops   src/BUGSPrimitives/functions.jl:15-26,87-89 via the link rewrite
//              src/parser/bugs_macro.jl:159-182
// Everything is templated on compile-time family/op codes where the caller knows them (static
// plate specs) and falls back to a warp-uniform switch otherwise.
#pragma once

#ifdef __CUDACC_RTC__
// run-time specialisation (jbugs_jit.inc) compiles this header with NVRTC: no host headers, the device math and the
// fixed-width integer types are built in or defined here; the headers come embedded in libjbugs_cuda (flat names)
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
#define JBUGS_PLAN_NO_STDINT 1
#include "jbugs_plan.h"
#else
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/jbugs_plan.h"
#endif

namespace jbugs {

constexpr double kLog2Pi = 1.8378770664093454835606594728112;
constexpr double kLogisticLower = -744.4400719213812;  // LogExpFunctions._logistic_bounds(Float64)
constexpr double kLogisticUpper = 36.7368005696771;
constexpr double kInvSqrt2 = 0.70710678118654752440;
constexpr double kInvSqrt2Pi = 0.39894228040143267794;

__device__ __forceinline__ double dev_inf() { return __longlong_as_double(0x7ff0000000000000LL); }
__device__ __forceinline__ double dev_nan() { return __longlong_as_double(0x7ff8000000000000LL); }

__device__ __forceinline__ double digamma(double x) {
    // upward recurrence to x >= 10 then the asymptotic series (|err| < 1e-15 for x > 0)
    double r = 0.0;
    if (x <= 0.0) {
        if (x == floor(x)) return dev_nan();
        const double pi = 3.14159265358979323846;
        r = -pi / tan(pi * x);
        x = 1.0 - x;
    }
    while (x < 10.0) {
        r -= 1.0 / x;
        x += 1.0;
    }
    const double f = 1.0 / (x * x);
    const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 +
                     f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
    return r + log(x) - 0.5 / x + t;
}

__device__ __forceinline__ double logistic(double x) {
    const double e = exp(x);
    const double p = e / (1.0 + e);
    return x < kLogisticLower ? 0.0 : (x > kLogisticUpper ? 1.0 : p);
}

__device__ __forceinline__ double xlogy(double x, double y) { return x == 0.0 ? 0.0 : x * log(y); }
__device__ __forceinline__ double xlog1py(double x, double y) { return x == 0.0 ? 0.0 : x * log1p(y); }
__device__ __forceinline__ double logbeta(double a, double b) { return lgamma(a) + lgamma(b) - lgamma(a + b); }

// ---- unary ops: value v and derivative d; returns true on DomainError
__device__ __forceinline__ bool unary_op(int op, double x, double& v, double& d) {
    switch (op) {
    case JBUGS_OP_EXP: v = exp(x); d = v; return false;
    case JBUGS_OP_LOGISTIC: v = logistic(x); d = v * (1.0 - v); return false;
    case JBUGS_OP_CEXPEXP: { const double e = exp(x); v = -expm1(-e); d = exp(x - e); return false; }
    case JBUGS_OP_PHI: v = 0.5 * erfc(-x * kInvSqrt2); d = kInvSqrt2Pi * exp(-0.5 * x * x); return false;
    case JBUGS_OP_LOG: v = log(x); d = 1.0 / x; return x < 0.0;
    case JBUGS_OP_SQRT: v = sqrt(x); d = 0.5 / v; return x < 0.0;
    case JBUGS_OP_NEG: v = -x; d = -1.0; return false;
    default: v = x; d = 1.0; return false;
    }
}

// ---- families: log density at x given a0,a1 ; partials wrt x, a0, a1. Returns true on DomainError.
// `need_norm` == false lets observation plates skip theta-independent normalising constants?  No:
// the reference's logp includes them (golden constants), so they are always computed.
struct FamOut {
    double lp, dx, da0, da1;
};

// a2: third argument of the three-argument families (dt's degrees of freedom), a constant: it has no adjoint
__device__ inline bool fam_eval(int fam, double x, double a0, double a1, FamOut& o, double a2 = 0.0) {
    o.dx = 0.0; o.da0 = 0.0; o.da1 = 0.0; o.lp = 0.0;
    switch (fam) {
    case JBUGS_FAM_NORMAL: {
        const double sigma = sqrt(1.0 / a1);
        const double r = x - a0;
        const double z = r / sigma;
        o.lp = -(z * z + kLog2Pi) * 0.5 - log(sigma);
        o.dx = -a1 * r; o.da0 = a1 * r; o.da1 = 0.5 / a1 - 0.5 * r * r;
        return a1 < 0.0;
    }
    case JBUGS_FAM_GAMMA: {
        const double theta = 1.0 / a1;
        const bool bad = !(a0 > 0.0) || !(theta > 0.0);
        const double xt = x / theta;
        o.lp = x < 0.0 ? -dev_inf() : (-lgamma(a0) + xlogy(a0 - 1.0, xt) - log(theta) - xt);
        o.dx = (a0 - 1.0) / x - a1;
        o.da0 = -digamma(a0) + log(xt);
        o.da1 = a0 / a1 - x;
        return bad;
    }
    case JBUGS_FAM_EXPONENTIAL: {
        const double theta = 1.0 / a0;
        const double rate = 1.0 / theta;
        o.lp = x < 0.0 ? -dev_inf() : (log(rate) - rate * x);
        o.dx = -rate; o.da0 = 1.0 / a0 - x;
        return !(theta > 0.0);
    }
    case JBUGS_FAM_UNIFORM: {
        const double w = a1 - a0;
        o.lp = (x < a0 || x > a1) ? -dev_inf() : -log(w);
        o.da0 = 1.0 / w; o.da1 = -1.0 / w;
        return !(a0 < a1);
    }
    case JBUGS_FAM_BETA: {
        const bool out = (x < 0.0 || x > 1.0);
        o.lp = out ? -dev_inf() : (xlogy(a0 - 1.0, x) + xlog1py(a1 - 1.0, -x) - logbeta(a0, a1));
        o.dx = (a0 - 1.0) / x - (a1 - 1.0) / (1.0 - x);
        const double pab = digamma(a0 + a1);
        o.da0 = log(x) - digamma(a0) + pab;
        o.da1 = log1p(-x) - digamma(a1) + pab;
        return !(a0 > 0.0) || !(a1 > 0.0);
    }
    case JBUGS_FAM_LOGNORMAL: {
        const double sigma = sqrt(1.0 / a1);
        const double lx = log(x), r = lx - a0, z = r / sigma;
        o.lp = x <= 0.0 ? -dev_inf() : (-(z * z + kLog2Pi) * 0.5 - log(sigma) - lx);
        o.dx = -(a1 * r + 1.0) / x; o.da0 = a1 * r; o.da1 = 0.5 / a1 - 0.5 * r * r;
        return a1 < 0.0 || !(sigma > 0.0);
    }
    case JBUGS_FAM_BERNOULLI: {
        const double q = 1.0 - a0;
        if (x == 0.0) { o.lp = log(q); o.da0 = -1.0 / q; }
        else if (x == 1.0) { o.lp = log(a0); o.da0 = 1.0 / a0; }
        else o.lp = -dev_inf();
        return !(a0 >= 0.0 && a0 <= 1.0);
    }
    case JBUGS_FAM_BINOMIAL: {
        const double p = a0, n = a1, m = n - x;
        const bool out = (x < 0.0 || x > n || x != floor(x));
        o.lp = out ? -dev_inf()
                   : (xlogy(x, p) + xlog1py(m, -p) - logbeta(x + 1.0, m + 1.0) - log(n + 1.0));
        o.da0 = (x == 0.0 ? 0.0 : x / p) - (m == 0.0 ? 0.0 : m / (1.0 - p));
        return !(n >= 0.0) || !(p >= 0.0 && p <= 1.0);
    }
    case JBUGS_FAM_POISSON: {
        const bool out = (x < 0.0 || x != floor(x));
        o.lp = out ? -dev_inf() : (xlogy(x, a0) - a0 - lgamma(x + 1.0));
        o.da0 = (x == 0.0 ? 0.0 : x / a0) - 1.0;
        return !(a0 >= 0.0);
    }
    case JBUGS_FAM_LOGISTIC: {  // Logistic(mu, 1/sqrt(tau)): u - 2 log1pexp(u) - log(s), u = -|z|
        const double rt = sqrt(a1), s = 1.0 / rt;
        const double z = (x - a0) * rt, u = -fabs(z);
        o.lp = u - 2.0 * log1p(exp(u)) - log(s);
        const double w = 1.0 - 2.0 * logistic(z);  // -tanh(z / 2)
        o.dx = w * rt; o.da0 = -w * rt; o.da1 = w * (x - a0) / (2.0 * rt) + 0.5 / a1;
        return a1 < 0.0 || !(s > 0.0);
    }
    case JBUGS_FAM_LAPLACE: {  // Laplace(mu, 1/sqrt(tau)): -(|x - mu| / b + log(2 b))
        const double rt = sqrt(a1), b = 1.0 / rt;
        const double r = x - a0, sg = r >= 0.0 ? 1.0 : -1.0;
        o.lp = -(fabs(r) / b + log(2.0 * b));
        o.dx = -sg * rt; o.da0 = sg * rt; o.da1 = -fabs(r) / (2.0 * rt) + 0.5 / a1;
        return a1 < 0.0 || !(b > 0.0);
    }
    case JBUGS_FAM_WEIBULL: {  // Weibull(a, 1/b): log(a / theta) + xlogy(a - 1, z) - z^a, z = x / theta
        const double theta = 1.0 / a1;
        const double z = x / theta, lz = log(z), za = exp(a0 * lz);
        o.lp = x < 0.0 ? -dev_inf() : (log(a0 / theta) + xlogy(a0 - 1.0, z) - za);
        o.dx = ((a0 - 1.0) - a0 * za) / x;
        o.da0 = 1.0 / a0 + lz - za * lz;
        o.da1 = a0 * (1.0 - za) / a1;
        return !(a0 > 0.0) || !(theta > 0.0);
    }
    case JBUGS_FAM_NEGBIN: {  // NegativeBinomial(r = a1, p = a0)
        const bool out = (x < 0.0 || x != floor(x));
        o.lp = out ? -dev_inf() : (xlogy(a1, a0) + xlog1py(x, -a0) - log(x + a1) - logbeta(a1, x + 1.0));
        o.da0 = a1 / a0 - (x == 0.0 ? 0.0 : x / (1.0 - a0));
        o.da1 = log(a0) - 1.0 / (x + a1) - digamma(a1) + digamma(a1 + x + 1.0);
        return !(a1 > 0.0) || !(a0 > 0.0 && a0 <= 1.0);
    }
    case JBUGS_FAM_CHISQ: {  // Chisq(k) = Gamma(k / 2, 2)
        const double h = 0.5 * a0;
        o.lp = x < 0.0 ? -dev_inf() : (-lgamma(h) - h * 0.69314718055994530942 + xlogy(h - 1.0, x) - 0.5 * x);
        o.dx = (h - 1.0) / x - 0.5;
        o.da0 = 0.5 * (-digamma(h) - 0.69314718055994530942 + log(x));
        return !(a0 > 0.0);
    }
    case JBUGS_FAM_T: {  // TDistShiftedScaled(nu = a2, mu = a0, sigma = 1/sqrt(tau)): tdistlogpdf(nu, z) - log(sigma)
        const double nu = a2, rt = sqrt(a1), z = (x - a0) * rt, w = z * z / nu;
        o.lp = lgamma(0.5 * (nu + 1.0)) - 0.5 * log(nu * 3.14159265358979323846) - lgamma(0.5 * nu) - 0.5 * (nu + 1.0) * log1p(w) + 0.5 * log(a1);
        const double s = (nu + 1.0) * z / (nu + z * z);  // -d/dz of the log-density
        o.dx = -s * rt; o.da0 = s * rt;
        o.da1 = (0.5 / a1) * (1.0 - s * z);
        return a1 < 0.0 || !(nu > 0.0) || !(rt > 0.0);
    }
    case JBUGS_FAM_GEOMETRIC: {  // Geometric(p = a0): log(p) + xlog1py(x, -p)
        const bool out = (x < 0.0 || x != floor(x));
        o.lp = out ? -dev_inf() : (log(a0) + xlog1py(x, -a0));
        o.da0 = 1.0 / a0 - (x == 0.0 ? 0.0 : x / (1.0 - a0));
        return !(a0 > 0.0 && a0 <= 1.0);
    }
    case JBUGS_FAM_CATEGORICAL: {  // Categorical(p): a0 = p[x], the probability of the observed category
        o.lp = log(a0);
        o.da0 = 1.0 / a0;
        return !(a0 >= 0.0 && a0 <= 1.0 + 1e-8);
    }
    case JBUGS_FAM_BETABIN: {  // BetaBinomial(n = a2, alpha = a0, beta = a1)
        const double n = a2, m = n - x;
        const bool out = (x < 0.0 || x > n || x != floor(x));
        const double xs = out ? 0.0 : x, ms = out ? n : m;
        o.lp = out ? -dev_inf() : (-log1p(n) - logbeta(xs + 1.0, ms + 1.0) + logbeta(xs + a0, ms + a1) - logbeta(a0, a1));
        const double dsum = digamma(a0 + a1) - digamma(n + a0 + a1);
        o.da0 = digamma(xs + a0) - digamma(a0) + dsum;
        o.da1 = digamma(ms + a1) - digamma(a1) + dsum;
        return !(n >= 0.0) || !(a0 >= 0.0) || !(a1 >= 0.0);
    }
    default: return false;  // FLAT
    }
}

// ---- logit-link building block: e = exp(-|eta|), 1/(1+e), log1p(e) -- branch free, coefficients in constant memory
// (the CUDA math library's exp / log1p cost ~230 issue slots per observation in the binomial plate kernel, half of
// them UMOV/IMAD constant materialisation and divergence bookkeeping; this version is ~50 FP64 instructions whose
// coefficient operands come from the constant bank, and the observations of an unrolled block interleave freely).
// Accuracy (the same double arithmetic replayed on the host against 50-digit values, 20000 points over [1e-12, 700]):
// relative error <= 4.4e-16 for exp(-|eta|), 4.0e-16 for 1/(1+e), 6.8e-16 for log1p(e).
// Polynomials: Chebyshev interpolants converted to the monomial basis in 60-digit arithmetic (highest degree first).
//   exp(r),  |r| <= ln2/2:           degree 10, max relative error 4.1e-16 (a degree-12 Taylor polynomial gives 1.4e-16)
//   atanh(s)/s in z = s^2 <= 1/9:    degree  9, max relative error 2.0e-16 (the Taylor series needs 16 terms for that)
static __constant__ double kExpPoly[11] = {
    2.762635737968284e-07,
    2.764018096215644e-06,
    2.4801504346665615e-05,
    0.0001984117027004143,
    0.001388888893248886,
    0.008333333385668096,
    0.04166666666657314,
    0.16666666666554403,
    0.5000000000000006,
    1.0000000000000067,
    1.0};
static __constant__ double kAtanhPoly[10] = {
    0.08890381913056651,
    0.04940394524382393,
    0.06795652742789947,
    0.07681838124445817,
    0.09091429900236955,
    0.11111095430673142,
    0.1428571455582858,
    0.199999999976438,
    0.33333333333341286,
    1.0};

struct SoftplusParts {
    double e;    // exp(-|eta|)
    double inv;  // 1 / (1 + e)
    double l1p;  // log(1 + e)
};

// 2 * (atanh(s) / s) coefficients: l1p = s * poly(s^2) with the factor 2 folded in (exact: a power of two)
static __constant__ double kAtanhPoly2[10] = {
    2 * 0.08890381913056651,
    2 * 0.04940394524382393,
    2 * 0.06795652742789947,
    2 * 0.07681838124445817,
    2 * 0.09091429900236955,
    2 * 0.11111095430673142,
    2 * 0.1428571455582858,
    2 * 0.199999999976438,
    2 * 0.33333333333341286,
    2.0};

// approximate reciprocal of a double from its upper word only (SASS MUFU.RCP64H, ~20 significant bits): the seed of the
// cubic correction below.  One instruction, against eleven for the float round trip with its IEEE slow-path call.
__device__ __forceinline__ double rcp_seed(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    return r;
}

// Branch-free and free of FP64 compares / selects: everything that is not arithmetic runs on the integer pipe
// (the plate kernels that call this are bound by the FP64 pipe: 16 lanes per SM sub-partition, a warp-wide FP64
// instruction holds it for two cycles, so every FP64 instruction saved is two cycles per observation-warp).
//   hi_abs : upper word of |eta|, clamped by the caller's integer min to the upper word of 700.0 (e < 1e-304 beyond:
//            irrelevant next to |eta|; the exponent field below must not wrap).  NaN is clamped as well: the caller's
//            own use of eta keeps it alive.
// polynomial evaluation order: Horner (fewest instructions, one dependent chain of `degree` fused multiply-adds) or
// Estrin (three more multiplies, depth 4): the plate kernels interleave several observations per thread, which hides a
// Horner chain; the build picks with -DJB_ESTRIN=1 (measured: see DESIGN section 3)
#ifndef JB_ESTRIN
#define JB_ESTRIN 0
#endif
__device__ __forceinline__ double exp_poly(double r) {  // degree 10, kExpPoly highest degree first
#if JB_ESTRIN
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double a0 = fma(kExpPoly[9], r, kExpPoly[10]), a1 = fma(kExpPoly[7], r, kExpPoly[8]),
                 a2 = fma(kExpPoly[5], r, kExpPoly[6]), a3 = fma(kExpPoly[3], r, kExpPoly[4]),
                 a4 = fma(kExpPoly[1], r, kExpPoly[2]);
    const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(kExpPoly[0], r2, a4);
    return fma(b2, r8, fma(b1, r4, b0));
#else
    double p = kExpPoly[0];
#pragma unroll
    for (int n = 1; n < 11; ++n) p = fma(p, r, kExpPoly[n]);
    return p;
#endif
}
__device__ __forceinline__ double atanh_poly2(double z) {  // degree 9 in z = s^2, kAtanhPoly2 highest degree first
#if JB_ESTRIN
    const double z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
    const double e0 = fma(kAtanhPoly2[8], z, kAtanhPoly2[9]), e1 = fma(kAtanhPoly2[6], z, kAtanhPoly2[7]),
                 e2 = fma(kAtanhPoly2[4], z, kAtanhPoly2[5]), e3 = fma(kAtanhPoly2[2], z, kAtanhPoly2[3]),
                 e4 = fma(kAtanhPoly2[0], z, kAtanhPoly2[1]);
    const double f0 = fma(e1, z2, e0), f1 = fma(e3, z2, e2);
    return fma(e4, z8, fma(f1, z4, f0));
#else
    double q = kAtanhPoly2[0];
#pragma unroll
    for (int n = 1; n < 10; ++n) q = fma(q, z, kAtanhPoly2[n]);
    return q;
#endif
}

__device__ __forceinline__ SoftplusParts softplus_parts(double ax /* |eta| (<= ~700, see above) */) {
    SoftplusParts o;
    // exp(-a), a in [0, 700]: -a = k ln2 + r, |r| <= ln2/2, degree-10 polynomial, scaling by 2^k on the integer pipe
    // (the sign of a lives in the operand modifiers of the fma's: no separate negation on the FP64 pipe)
    const double magic = 6755399441055744.0;  // 1.5 * 2^52: the low word of (t + magic) is rint(t)
    const double t = fma(ax, -1.4426950408889634074, magic);
    const int k = __double2loint(t);
    const double kf = t - magic;
    double r = fma(kf, -6.93147180369123816490e-01, -ax);
    r = fma(kf, -1.90821492927058770002e-10, r);
    const double p = exp_poly(r);
    // p in [0.70, 1.42], k >= -1010: adding k to the exponent field keeps the number normal -- no FP64 multiply
    o.e = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    // one reciprocal serves both 1/(1+e) and e/(2+e): d = (1+e)(2+e) in (2, 6]
    const double u1 = 1.0 + o.e, u2 = 2.0 + o.e, d = u1 * u2;
    // ~20-bit seed, then one cubic step rc (1 + h + h^2), h = 1 - d rc < 2^-19: error h^3 < 1e-17 (three DFMA)
    double rc = rcp_seed(d);
    const double h = fma(-d, rc, 1.0);
    rc = fma(rc, fma(h, h, h), rc);
    o.inv = rc * u2;
    // log1p(e) = 2 atanh(s), s = e / (2 + e) in (0, 1/3]
    const double s = o.e * (rc * u1);
    o.l1p = s * atanh_poly2(s * s);
    return o;
}

// |eta| with its upper word clamped to that of 700.0 (integer min: for non-negative doubles the bit patterns order like
// the values), and the upper word itself for range tests on the integer pipe
constexpr int kHi700 = 0x4085E000;        // 700.0 = 0x4085E00000000000
constexpr int kHiLogisticUpper = 0x40425E4F;  // upper word of kLogisticUpper = 36.7368005696771
__device__ __forceinline__ double abs_clamped(double eta, int& hi_abs) {
    hi_abs = __double2hiint(eta) & 0x7fffffff;
    return __hiloint2double(min(hi_abs, kHi700), __double2loint(eta));
}

// k ~ dbin(logistic(eta), n)  (dbern: n = 1), the theta-dependent part of the log-density and d/d eta:
//   k log p + m log(1 - p),  m = n - k,  log p = -(max(-eta, 0) + l1p),  log(1 - p) = -(max(eta, 0) + l1p),
//   max(x, 0) = (|x| + x) / 2   =>   = -[ n (|eta| / 2 + l1p) + (n / 2 - k) eta ]          (no select)
//   p = 1/2 + copysign(1 / (1 + e) - 1/2, eta)   =>   d/d eta = k - n p = -[(n / 2 - k) + n copysign(...)]
// The data-only constant -logbeta(k+1, m+1) - log(n+1) is added once per plate (obs_const).  `sat` collects (integer
// pipe) whether |eta| reached the range where the reference's logistic saturates to exactly 0 / 1; the caller then runs
// logit_tail_is_inf on the block -- rarely -- to reproduce the reference's -Inf there.
__device__ __forceinline__ void binomial_logit(double eta, double k, double n, double& lp, double& deta, int& sat) {
    int hi_abs;
    const double ax = abs_clamped(eta, hi_abs);
    const SoftplusParts sp = softplus_parts(ax);
    const double hd = fma(0.5, n, -k);
    lp = fma(-n, fma(0.5, fabs(eta), sp.l1p), lp);
    lp = fma(-hd, eta, lp);
    const double cq = copysign(sp.inv - 0.5, eta);
    deta = -fma(n, cq, hd);
    sat |= (hi_abs >= kHiLogisticUpper) ? 1 : 0;
}
// LogExpFunctions.logistic returns exactly 1 above kLogisticUpper and exactly 0 below kLogisticLower: the binomial
// log-density is then -Inf unless the corresponding count is zero
__device__ __forceinline__ bool logit_tail_is_inf(double eta, double k, double n) {
    return (eta > kLogisticUpper && n - k != 0.0) || (eta < kLogisticLower && k != 0.0);
}

// ---- bijectors: slot y -> value x, logjac lj, dx/dy, d(lj)/dy
struct TrOut {
    double x, lj, dxdy, dljdy;
};

__device__ __forceinline__ TrOut transform_eval(int tr, double lb, double ub, double y) {
    TrOut o;
    switch (tr) {
    case JBUGS_TR_LOG: { const double e = exp(y); o.x = lb + e; o.lj = y; o.dxdy = e; o.dljdy = 1.0; break; }
    case JBUGS_TR_UPPER: { const double e = exp(y); o.x = ub - e; o.lj = y; o.dxdy = -e; o.dljdy = 1.0; break; }
    case JBUGS_TR_LOGIT: {
        const double s = logistic(y), w = ub - lb;
        o.x = lb + w * s;
        o.lj = log((o.x - lb) * (ub - o.x) / w);
        o.dxdy = w * s * (1.0 - s);
        o.dljdy = 1.0 - 2.0 * s;
        break;
    }
    default: o.x = y; o.lj = 0.0; o.dxdy = 1.0; o.dljdy = 0.0; break;
    }
    return o;
}

}  // namespace jbugs
