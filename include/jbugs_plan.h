/*
 * jbugs_plan.h -- the flat "kernel plan" a BUGS model is lowered to.
 *
 * One plan describes everything `evaluate_with_values!!`
 * (JuliaBUGS/src/model/evaluation.jl:217-275) walks node by node, regrouped at
 * statement/plate granularity the way the reference's own generated evaluator is
 * (JuliaBUGS/src/source_gen.jl:459-552, 615-666, 670-774):
 *
 *   globals  scalar model parameters (theta slot, bijector, prior)       -- evaluation.jl:240-260
 *   gtape    scalar logical nodes that depend only on globals            -- evaluation.jl:237-239
 *   plates   a loop nest (group i in [0,N), inner j in [0,T)) holding
 *              locals : per-group / per-observation parameters with their priors
 *              terms  : the linear predictor  eta = sum_t coef_t * cov_t (logical parents inlined)
 *              link   : tape of unary inverse-link ops applied to eta     -- bugs_macro.jl:159-182
 *              obs    : the observed stochastic statement (family, y, aux args) -- evaluation.jl:264-266
 *
 * The blob is a header followed by packed arrays of the records below, in this order:
 *   header | globals[n_globals] | gtape[n_gtape] | data[n_data] | plates[n_plates]
 *          | locals[n_locals] | terms[n_terms] | dense[n_dense] | xops[n_xops]
 * All integers little-endian, all reals IEEE-754 binary64 (the reference computes in Float64,
 * indices in Int64).  Arrays of observations/covariates are NOT in the blob: they are uploaded
 * per data slot (jbugs_plan_upload_data) so that `set_observed_values!`
 * (JuliaBUGS/src/model/abstractppl.jl:872-888) maps to a re-upload, not a re-plan.
 *
 * Index conventions: zero-based.  A data slot of rank 2 is stored column-major (Julia), i.e.
 * element (i, j) of an N x T array is at i + N * j.
 */
#ifndef JBUGS_PLAN_H
#define JBUGS_PLAN_H

#ifndef JBUGS_PLAN_NO_STDINT /* (set by the NVRTC build of the device headers, which defines the integer types itself) */
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define JBUGS_PLAN_MAGIC "JBUGSPL1"
#define JBUGS_PLAN_VERSION 1u

/* header.flags */
#define JBUGS_FLAG_TRANSFORMED 1u /* theta lives in unconstrained space (model.transformed) */

/* ---- argument references ------------------------------------------------------------------ */
enum jbugs_arg_kind {
    JBUGS_ARG_CONST = 0,   /* value                                                   */
    JBUGS_ARG_GVAL = 1,    /* id = global value slot (a global parameter's constrained
                              value, or a gtape result)                               */
    JBUGS_ARG_LOCAL = 2,   /* id = local index inside the plate (constrained value)   */
    JBUGS_ARG_DATA_I = 3,  /* id = data slot, indexed by the group index i            */
    JBUGS_ARG_DATA_J = 4,  /* id = data slot, indexed by the inner index j            */
    JBUGS_ARG_DATA_IJ = 5, /* id = data slot (N x T column-major), indexed by (i, j)  */
    JBUGS_ARG_ONE = 6,     /* the constant 1 (covariate of an intercept-like term)    */
    JBUGS_ARG_XVAL = 7,    /* id = value slot of the plate's expression tape          */
    JBUGS_ARG_GVEC_J = 8   /* id = global value slot of element 0 of a (short) parameter
                              vector indexed by the inner index: gval[id + j]         */
};

typedef struct jbugs_arg {
    uint32_t kind;
    int32_t id;
    double value;
} jbugs_arg; /* 16 bytes */

/* ---- distribution families (BUGSPrimitives/distributions.jl parameterisation) ------------- */
enum jbugs_family {
    JBUGS_FAM_NORMAL = 0,      /* dnorm(mu, tau)   :11-18   */
    JBUGS_FAM_GAMMA = 1,       /* dgamma(a, b)     :200-203 */
    JBUGS_FAM_EXPONENTIAL = 2, /* dexp(lambda)     :186-188 */
    JBUGS_FAM_UNIFORM = 3,     /* dunif(a, b)      :343-345 */
    JBUGS_FAM_BETA = 4,        /* dbeta(a, b)      :357-359 */
    JBUGS_FAM_LOGNORMAL = 5,   /* dlnorm(mu, tau)           */
    JBUGS_FAM_BERNOULLI = 6,   /* dbern(p)         :442-444 */
    JBUGS_FAM_BINOMIAL = 7,    /* dbin(p, n)       :458-460 */
    JBUGS_FAM_POISSON = 8,     /* dpois(lambda)    :486-488 */
    JBUGS_FAM_FLAT = 9,        /* dflat()                   */
    JBUGS_FAM_LOGISTIC = 10,   /* dlogis(mu, tau)  :30-33   Logistic(mu, 1/sqrt(tau)) */
    JBUGS_FAM_LAPLACE = 11,    /* ddexp(mu, tau)   :92-95   Laplace(mu, 1/sqrt(tau))  */
    JBUGS_FAM_WEIBULL = 12,    /* dweib(a, b)      :233-235 Weibull(a, 1/b)           */
    JBUGS_FAM_NEGBIN = 13,     /* dnegbin(p, r)    :516-518 NegativeBinomial(r, p)    */
    JBUGS_FAM_CHISQ = 14,      /* dchisqr(k)       :215-217 Chisq(k)                  */
    JBUGS_FAM_T = 15,          /* dt(mu, tau, nu)  :73-80   TDistShiftedScaled(nu, mu, 1/sqrt(tau)); nu a constant */
    JBUGS_FAM_GEOMETRIC = 16,  /* dgeom(p)         :500-502 Geometric(p): failures before the first success */
    JBUGS_FAM_CATEGORICAL = 17,/* dcat(p[1:K])     :472-474 Categorical(p): the plate's argument is p[y], the
                                  probability of the observed category (log density = log p[y]) */
    JBUGS_FAM_BETABIN = 18,    /* dbetabin(a, b, n) :530-532 BetaBinomial(n, a, b); n a constant */
    JBUGS_FAM_COUNT = 19
};

/* ---- constrained <-> unconstrained maps (Bijectors, call site evaluation.jl:243-257) ------- */
enum jbugs_transform {
    JBUGS_TR_IDENTITY = 0, /* x = y,                  logjac 0                          */
    JBUGS_TR_LOG = 1,      /* x = lb + exp(y),        logjac y                          */
    JBUGS_TR_LOGIT = 2,    /* x = lb + (ub-lb) s(y),  logjac log((x-lb)(ub-x)/(ub-lb))  */
    JBUGS_TR_UPPER = 3     /* x = ub - exp(y),        logjac y                          */
};

/* ---- unary ops of the link tape and of the global tape ------------------------------------ */
enum jbugs_op {
    JBUGS_OP_IDENTITY = 0,
    JBUGS_OP_EXP = 1,      /* log link      */
    JBUGS_OP_LOGISTIC = 2, /* logit link    */
    JBUGS_OP_CEXPEXP = 3,  /* cloglog link  */
    JBUGS_OP_PHI = 4,      /* probit link   */
    JBUGS_OP_LOG = 5,
    JBUGS_OP_SQRT = 6,
    JBUGS_OP_NEG = 7,
    /* binary ops (gtape only) */
    JBUGS_OP_ADD = 16,
    JBUGS_OP_SUB = 17,
    JBUGS_OP_MUL = 18,
    JBUGS_OP_DIV = 19,
    JBUGS_OP_POW = 20,
    /* expression tapes only */
    JBUGS_OP_LEAF = 32,  /* value of `leaf`: CONST | ONE | GVAL | GVEC_J | LOCAL | DATA_I | DATA_J | DATA_IJ */
    JBUGS_OP_SELECT = 33,/* v[a] != 0 ? v[b] : v[c]  (a: a 0/1 data leaf; no adjoint flows into a)          */
    JBUGS_OP_PICK_K = 34,/* v[a + k], k = the state of the plate's marginalised discrete latent (has_obs == 3):
                            slots a .. a + K - 1 hold the K candidates (`mu[z[i]]` -> mu[1], ..., mu[K])         */
    JBUGS_OP_GVEC_AT = 35/* gval[leaf.id + (int) v[a]]: an element of a short parameter vector whose elements are the
                            consecutive global values leaf.id .. leaf.id + leaf.value - 1, picked by the 0-based index
                            in tape slot a (a data leaf): the frailty `b[pair[i]]` of LeukFr inside an expression plate.
                            The adjoint goes to the picked element; none flows into the index                       */
};

/* ---- expression tape of a plate whose predictor is not linear in the parameters -----------------------------
 * (dugongs `alpha - beta * pow(gamma, x[i])`, lsat `logistic(beta * theta[j] - alpha[k])`, the cumulative-logit
 * probabilities behind Bones' `dcat`).  The node functions of the reference (compiler_pass.jl:751-814) composed along
 * the chain of logical parents of ONE observation, in static single assignment form: op k produces value slot k from
 * slots a, b, c < k.  The forward sweep fills the slots, the reverse sweep carries adjoints back to the leaves. */
#define JBUGS_MAX_XOPS 64
#define JBUGS_MAX_STATES 16 /* states of a discrete latent that is summed out per observation (has_obs == 3) */
typedef struct jbugs_xop {
    uint32_t op;       /* jbugs_op */
    uint16_t a, b, c;  /* operand slots */
    uint16_t pad;
    uint32_t pad2;
    jbugs_arg leaf;    /* JBUGS_OP_LEAF */
} jbugs_xop;           /* 32 bytes */

typedef struct jbugs_global {
    int64_t theta_idx;  /* slot in theta                                             */
    uint32_t transform; /* jbugs_transform (IDENTITY for every slot when !transformed) */
    uint32_t family;    /* prior                                                      */
    double lb, ub;      /* support bounds used by the transform.  A prior whose bounds are narrower than the
                           natural support of its family is truncated(dist(args), lb, ub) -- BUGS
                           `dist(args) T(lb, ub)`, bugs_parser.jl:470-500: literal arguments only, the
                           normaliser log(cdf(ub) - cdf(lb)) is folded into the prior's constant
                           (jbugs_trunc.h lists the families)                                        */
    jbugs_arg args[3];  /* prior arguments: CONST | GVAL (of an earlier global/gtape) */
} jbugs_global;         /* 80 bytes */

typedef struct jbugs_gop {
    uint32_t op;  /* jbugs_op */
    uint32_t pad;
    jbugs_arg a;  /* CONST | GVAL */
    jbugs_arg b;  /* CONST | GVAL (binary ops) */
} jbugs_gop;      /* 40 bytes; result goes to gval slot n_globals + k */

typedef struct jbugs_data_desc {
    int64_t len;   /* number of elements             */
    int64_t dim0;  /* leading dimension (rank 2)     */
    uint32_t rank; /* 1 or 2                         */
    uint32_t dtype;/* 0 = float64                    */
} jbugs_data_desc; /* 24 bytes */

enum jbugs_level { JBUGS_LEVEL_GROUP = 0, JBUGS_LEVEL_OBS = 1 };

typedef struct jbugs_local {
    uint32_t level;     /* GROUP: one per i;  OBS: one per (i, j)                        */
    uint32_t transform;
    int64_t theta_base; /* theta slot = theta_base + stride_i * i + stride_j * j         */
    int64_t stride_i;
    int64_t stride_j;
    int32_t idx_slot;   /* >= 0: explicit int64 index array (data slot) replaces the map */
    uint32_t family;    /* prior                                                         */
    double lb, ub;
    jbugs_arg args[3];  /* CONST | GVAL | DATA_I | DATA_J | DATA_IJ                       */
} jbugs_local;          /* 104 bytes */

typedef struct jbugs_term {
    jbugs_arg coef; /* CONST | GVAL | LOCAL          */
    jbugs_arg cov;  /* ONE | CONST | DATA_I | DATA_J | DATA_IJ */
} jbugs_term;       /* 32 bytes */

#define JBUGS_MAX_LINK 4

typedef struct jbugs_plate {
    int64_t N; /* groups */
    int64_t T; /* observations per group */
    uint32_t first_local, n_locals;
    uint32_t first_term, n_terms;
    uint32_t n_link;
    uint32_t link[JBUGS_MAX_LINK]; /* applied to eta in order */
    uint32_t family;               /* observation family */
    uint32_t eta_arg;              /* which family argument receives link(eta) */
    uint32_t has_obs;              /* 0: prior-only plate (locals without a likelihood); 1: linear predictor (terms +
                                      link); 2: expression plate -- first_term / n_terms address xops[] instead of
                                      terms[], n_link = 0, and the family arguments that depend on parameters are
                                      XVAL references into the tape; 3: expression plate with a finite discrete
                                      latent per observation that is summed out (auto-marginalisation,
                                      evaluation.jl:739-961): eta_arg holds the number of states K, link[0] the tape
                                      slot of the latent's log prior log p(z = k), link[1] - 1 the tape slot of a data leaf
                                      with the observed state of a partially observed latent (0 = missing; link[1] == 0:
                                      no cell is observed); the cell contributes
                                      logsumexp_k [ log p(z = k) + log p(y | z = k) ] */
    jbugs_arg y;                   /* DATA_I | DATA_IJ */
    jbugs_arg aux[3];              /* the family's arguments; aux[eta_arg] is ignored (has_obs == 1) */
} jbugs_plate;                     /* 128 bytes */

/* Dense linear predictor:  y[i] ~ dbern|dbin(logistic(inprod(X[i, 1:P], beta[1:P])) [, n[i]])
 * (`inprod`: JuliaBUGS/src/BUGSPrimitives/functions.jl:33-35).  beta's prior is an ordinary prior-only plate
 * (has_obs = 0) over j in [0, P); this record adds the likelihood and its gradient w.r.t. beta, evaluated as two
 * FP64 tensor-core GEMMs over all chains. */
typedef struct jbugs_dense {
    int64_t N, P;                      /* observations, coefficients                       */
    int64_t theta_base, theta_stride;  /* slot of beta[j] = theta_base + theta_stride * j  */
    int32_t x_slot, y_slot, n_slot;    /* X: N x P column-major; n_slot < 0: Bernoulli     */
    uint32_t family, link, pad;        /* BERNOULLI | BINOMIAL, JBUGS_OP_LOGISTIC          */
} jbugs_dense;                         /* 56 bytes */

typedef struct jbugs_plan_header {
    char magic[8];
    uint32_t version;
    uint32_t flags;
    int64_t D;
    uint32_t n_globals;
    uint32_t n_gtape;
    uint32_t n_data;
    uint32_t n_plates;
    uint32_t n_locals;
    uint32_t n_terms;
    uint32_t n_dense; /* jbugs_dense records appended after terms[] */
    uint32_t n_xops;  /* jbugs_xop records appended after dense[] */
    uint64_t reserved;
} jbugs_plan_header; /* 64 bytes */

#ifdef __cplusplus
}
#endif
#endif /* JBUGS_PLAN_H */
