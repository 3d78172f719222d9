// jbugs_kernels.cuh -- the batched forward+reverse plate kernels (sm_100a).
//
// Replaces the reference's node loop `evaluate_with_values!!`
// (/root/reference/JuliaBUGS/src/model/evaluation.jl:217-275) + its AD backend
// (src/model/logdensityproblems.jl:219-232) for a whole batch of chains.
//
// Mapping (native CHAIN_FAST layout, theta[d * ld + c]):
//   * one thread  <-> one chain (column c); a warp reads 32 consecutive chains of one parameter
//     row: every global access is a fully coalesced 256-byte row segment whatever the theta order
//     (UseGraph's reversed-interleaved AoS or the generated mode's order -- SURVEY 8a/A15);
//   * one CTA     <-> BLOCK chains x one tile of plate groups; the group loop is sequential
//     inside the thread, so logp and the gradients of the shared (global) parameters
//     accumulate in registers in a FIXED order -- no shuffles, no float atomics;
//   * observation / covariate arrays are indexed by the group only, i.e. warp-uniform: they are
//     staged once per CTA chunk into shared memory (bulk async copy when aligned) and every
//     lane reads them as broadcasts;
//   * per (tile, chain) partial sums go to `partial[tile][1 + n_gref][ldg]`; the finalize kernel
//     adds the tiles in ascending order, then the priors / bijectors of the globals.
//
// One body serves two kinds of instantiation: `DynSpec` reads the plate "shape" (families,
// link tape, argument kinds ...) from kernel parameters at run time; static specs supply the
// same accessors as constexpr so the compiler folds every switch and keeps all per-chain state
// in registers (the shape of the hot configurations in BASELINE.json).
#pragma once

#include "jbugs_device.cuh"

namespace jbugs {

constexpr int MAX_GT = 8;                        // terms whose coefficient is a global value
constexpr int MAX_LOC = 4;                       // local parameters of a plate
constexpr int MAX_TERMS = MAX_GT + MAX_LOC + 1;  // + optional data offset
constexpr int MAX_GREF = 24;                     // distinct global values a plate references (expression plates, compiled kernels)
constexpr int MAX_GREF_LINEAR = 12;              // ... a linear-predictor plate run by the interpreter (DynSpec keeps 12 in registers)
constexpr int MAX_GVALS = 64;                    // globals + gtape results of a plan
constexpr int BLOCK = 128;                       // chains per CTA
constexpr int CHAIN_OK = 0, CHAIN_DOMAIN = 1;    // JBUGS_CHAIN_OK / JBUGS_CHAIN_DOMAIN of jbugs_cuda.h (checked in jbugs_cuda.cu)

struct ArgS {  // structural half of an argument
    int kind;  // jbugs_arg_kind
    int id;    // GVAL: plate-local gref index; LOCAL: local index
    int soff;  // DATA_*: offset (in doubles) of the stream inside the staged chunk / J region.  Part of the SHAPE: a static
               // spec with a static inner extent knows its stream layout at compile time, so shared-memory addresses of
               // the data become immediates (one base register + constants instead of one address computation per stream)
};
struct ArgV {  // value half
    const double* ptr;  // DATA_*: device array (host-side bookkeeping / obs_const kernel)
    double value;       // CONST
    int soff;           // DATA_*: offset (in doubles) of the stream inside the staged chunk / J region
    int pad;
};

// Shared-memory staging of the plate's data (observations, covariates, data-valued prior arguments).
// All of it is indexed by the group (and inner) index only, i.e. identical for every chain of the CTA:
// it is copied once per chunk of STAGE_G groups with cp.async (double buffered) and read back as
// warp-wide broadcasts.  A stream is one distinct (data slot, index kind) pair.
constexpr int STAGE_G = 32;      // groups per staged chunk
constexpr int MAX_STREAMS = 20;
struct Stream {
    const double* ptr;
    int kind;  // DATA_I | DATA_J | DATA_IJ
    int soff;
};

struct LocalS {
    int level, transform, family, has_idx;
    ArgS args[3];
    int mu_zero;  // args[0] is the literal constant 0.0 (part of the shape: static specs may assume it)
};
struct LocalV {
    long long base, si, sj;
    long long row0, step_i, step_j;  // base * ld, si * ld, sj * ld: filled per launch (element offsets)
    const long long* idx;
    double lb, ub;
    ArgV args[3];
};

// one op of an expression tape on the device (jbugs_xop with its leaf resolved: plate-local gref index, stream offset)
struct XOpDev {
    int op;
    unsigned short a, b, c, pad;
    int kind, id, soff, pad2;  // leaf
    double value;
};

struct PlateShape {
    int kg, nl, has_off, n_gref;
    int family, eta_arg, has_obs, n_link;
    int link[JBUGS_MAX_LINK];
    ArgS cov[MAX_TERMS];
    ArgS y, aux[3];
    LocalS loc[MAX_LOC];
    int n_x;  // > 0: expression plate (has_obs == 2): the predictor is the tape PlateVals::xtape, aux[q] may be XVAL
    int n_states, lp_slot, zobs_slot;  // n_states > 0 (has_obs == 3 in the record): one discrete latent per observation is summed out
                            // over its n_states states; tape slot lp_slot holds log p(z = k), zobs_slot (or -1) the data
                            // leaf with the observed state of a partially observed latent (0 = missing)
};

struct PlateVals {
    long long N, T, i0, i1, tile;
    ArgV cov[MAX_TERMS];
    ArgV y, aux[3];
    LocalV loc[MAX_LOC];
    int gref[MAX_GREF];
    double obs_const;  // sum over the shard of the data-only normalising constants (fused paths)
    int n_streams;     // staged data streams
    int per_group;     // doubles per group in a staged chunk (I streams: 1, IJ streams: T)
    int j_doubles;     // size of the J region (J streams: T each), rounded up to even
    int bulk_ok;       // streams are 16-byte aligned for every even group index: chunks may use the TMA path
    Stream streams[MAX_STREAMS];
    const XOpDev* xtape;  // expression plates: sh.n_x ops in device memory (read-only, identical for every thread)
};

struct PlateParams {
    PlateShape sh;
    PlateVals v;
};

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream is still running; it must not touch the predecessor's results before
// `pdl_wait()`.  `pdl_launch_dependents()` lets the successor's CTAs be scheduled early, so that its launch latency and
// its own prologue (kernel parameters, data staging) overlap the predecessor's tail.  Both are no-ops in a kernel that
// was launched the ordinary way.  The evaluator is three dependent launches; on small models the gaps between them are a
// visible share of the 25-40 us an evaluation takes.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// Spec accessors.  DynSpec forwards to the run-time shape.
struct DynSpec {
    static constexpr bool STATIC = false;
    static constexpr int FUSED = 0;
    static constexpr int UNROLL = 1;
    // occupancy of the interpreter (round 2 A/B, profiles/r2_interp_ab.txt): 4 CTAs of 128 per SM at 128 registers with
    // ~600 bytes of spills beat 2 CTAs at 254 registers -- Rats variant 10^6 x 5 x 1024: 468 -> 299 ms, Epil 8192 chains
    // 392 -> 260 us, Pumps 47.6 -> 49.3 us; 6 and 8 CTAs lose again (309 / 322 ms)
#ifndef JB_DYN_MINB
#define JB_DYN_MINB 4
#endif
    static constexpr int MIN_BLOCKS = JB_DYN_MINB;
    static constexpr int T_STATIC = 0;  // inner extent read at run time
    static constexpr bool PREFETCH = false;
    static constexpr bool OBS_LOCALS = true;  // unknown at compile time
    static constexpr bool SOFF = false;       // stream offsets come from the kernel parameters
    static constexpr bool XTAPE = false;      // (XSpec: the instantiation that interprets expression tapes)
    static constexpr bool XGEN = false;       // (a run-time generated description whose tape is compiled code)
    static constexpr bool SMPF = false;       // (shared-memory theta prefetch: static specs only)
    // bounds of the unrolled loops over locals / global-coefficient terms (the arrays keep their MAX_* sizes: the kernel
    // parameter layout is one for all instantiations).  A static spec knows its counts: the front end then emits nl / kg
    // copies of each loop body instead of MAX_LOC / MAX_GT guarded ones for the optimiser to throw away.
    static constexpr int NL_MAX = MAX_LOC, KG_MAX = MAX_GT;
    // entries of ChainState::gv / ga this instantiation touches: the interpreter of linear-predictor plates keeps 12 (the
    // select chains of pick / bump and the register file are sized by it), expression plates and compiled kernels 24
    static constexpr int NG_MAX = MAX_GREF_LINEAR;
    static constexpr int L2PF = 0;            // no L2 prefetch of theta rows
    __device__ __forceinline__ bool prior_fused(int) const { return false; }
    __device__ __forceinline__ bool prior_mu_zero(int) const { return false; }
    const PlateShape& s;
    __device__ __forceinline__ explicit DynSpec(const PlateShape& sh) : s(sh) {}
    __device__ __forceinline__ int kg() const { return s.kg; }
    __device__ __forceinline__ int nl() const { return s.nl; }
    __device__ __forceinline__ int has_off() const { return s.has_off; }
    __device__ __forceinline__ int n_gref() const { return s.n_gref; }
    __device__ __forceinline__ int family() const { return s.family; }
    __device__ __forceinline__ int eta_arg() const { return s.eta_arg; }
    __device__ __forceinline__ int has_obs() const { return s.has_obs; }
    __device__ __forceinline__ int n_link() const { return s.n_link; }
    __device__ __forceinline__ int link(int q) const { return s.link[q]; }
    __device__ __forceinline__ ArgS cov(int t) const { return s.cov[t]; }
    __device__ __forceinline__ ArgS y() const { return s.y; }
    __device__ __forceinline__ ArgS aux(int q) const { return s.aux[q]; }
    // aux[2] holds an optional 0/1 observation weight (padded cells of a plate gathered by a data index)
    __device__ __forceinline__ bool has_w() const { return s.aux[2].kind == JBUGS_ARG_DATA_I || s.aux[2].kind == JBUGS_ARG_DATA_IJ; }
    __device__ __forceinline__ int loc_level(int l) const { return s.loc[l].level; }
    __device__ __forceinline__ int loc_transform(int l) const { return s.loc[l].transform; }
    __device__ __forceinline__ int loc_family(int l) const { return s.loc[l].family; }
    __device__ __forceinline__ int loc_has_idx(int l) const { return s.loc[l].has_idx; }
    __device__ __forceinline__ ArgS loc_arg(int l, int q) const { return s.loc[l].args[q]; }
};

// The interpreter for expression plates: DynSpec plus the tape sweep in observe() (its own instantiation, so that the
// per-thread value / adjoint arrays of the tape -- 2 KB of local memory -- are not part of every interpreter launch)
struct XSpec : DynSpec {
    static constexpr bool XTAPE = true;
    static constexpr int MIN_BLOCKS = 2;  // (the tape's value / adjoint arrays live in local memory anyway: keep the registers)
    __device__ __forceinline__ explicit XSpec(const PlateShape& sh) : DynSpec(sh) {}
    __device__ __forceinline__ int n_x() const { return s.n_x; }
    __device__ __forceinline__ int n_states() const { return s.n_states; }
    __device__ __forceinline__ int lp_slot() const { return s.lp_slot; }
    __device__ __forceinline__ int zobs_slot() const { return s.zobs_slot; }
};

// ... and the same for tapes that reference more than MAX_GREF_LINEAR global values (LeukFr's 21 frailties, a mixture of
// eight components): the wide select chains and the larger per-chain state cost the narrow interpreter a quarter of its speed
struct XWSpec : XSpec {
    static constexpr int NG_MAX = MAX_GREF;
    __device__ __forceinline__ explicit XWSpec(const PlateShape& sh) : XSpec(sh) {}
};

// fused observation paths (family x link pairs with a hand-derived forward+reverse)
enum { FUSED_NONE = 0, FUSED_NORMAL_IDENTITY = 1, FUSED_BINOMIAL_LOGIT = 2, FUSED_POISSON_LOG = 3,
       FUSED_BERNOULLI_LOGIT = 4,
       // any inverse-link tape (probit, cloglog, identity, ...): the kernel keeps the theta-dependent part of the
       // log-density, the data-only normaliser goes to obs_const exactly as for the logit / log versions
       FUSED_BINOMIAL_LINK = 5, FUSED_BERNOULLI_LINK = 6, FUSED_POISSON_LINK = 7 };

// does the description carry generated code for an expression tape (`XGEN`, `observe_gen`)?  Only run-time generated ones do.
template <class D, class = void>
struct desc_xgen { static constexpr bool value = false; };
template <class D>
struct desc_xgen<D, decltype((void)D::XGEN)> { static constexpr bool value = D::XGEN; };

// does the description ask for the shared-memory theta prefetch (`SMPF`)?
template <class D, class = void>
struct desc_smpf { static constexpr bool value = false; };
template <class D>
struct desc_smpf<D, decltype((void)D::SMPF)> { static constexpr bool value = D::SMPF; };

// A static spec is a struct with `static constexpr PlateShape SH` ; see specs in jbugs_cuda.cu.
// the shape of a description as ONE constant object in device memory: every accessor of StaticSpec is a load at a
// constant address that the optimiser folds (calling the constexpr sh() per access instead made NVVM inline and
// scalarise a 500-byte struct thousands of times: 47 s of optimiser time for one kernel under NVRTC against 0.1 s in
// the front end)
template <class Desc>
__device__ constexpr PlateShape g_shape_of = Desc::sh();

template <class Desc>
struct StaticSpec {
    using D = Desc;
    static constexpr int NL_MAX = Desc::sh().nl, KG_MAX = Desc::sh().kg;
    static constexpr int NG_MAX = MAX_GREF;   // (indices are compile-time constants here: unused entries disappear)
    static constexpr bool XGEN = desc_xgen<Desc>::value;
    // FP64-bound plates: the theta rows of the NEXT block of UNROLL groups travel global -> shared with cp.async while
    // this block is evaluated (no registers held across the block, unlike PREFETCH); group-level locals with affine
    // slot maps only.  The launcher adds smpf_doubles() of dynamic shared memory behind the staging buffers.
    static constexpr bool SMPF = desc_smpf<Desc>::value;
    static constexpr bool STATIC = true;
    static constexpr int FUSED = Desc::FUSED;
    static constexpr int UNROLL = Desc::UNROLL;
    static constexpr int MIN_BLOCKS = Desc::MIN_BLOCKS;
    static constexpr int T_STATIC = Desc::T_STATIC;  // > 0: inner loop fully unrolled
    static constexpr bool PREFETCH = Desc::PREFETCH; // theta loads of the next U-block issued before this one is evaluated
    static constexpr bool obs_locals_() {
        for (int l = 0; l < Desc::sh().nl; ++l)
            if (Desc::sh().loc[l].level == JBUGS_LEVEL_OBS) return true;
        return false;
    }
    static constexpr bool OBS_LOCALS = obs_locals_();  // the plate has observation-level (i, j) locals
    // static inner extent: the layout of the staged streams follows from the shape alone (canon_soffs, mirrored by
    // build_plate and checked by shape_equal), so stream offsets are compile-time constants
    static constexpr bool SOFF = Desc::T_STATIC > 0;
    static constexpr bool XTAPE = false;
    static constexpr int L2PF = Desc::L2PF;           // groups ahead whose theta rows are prefetched into L2 (0 = off)
    // a dnorm prior whose arguments do not vary over the plate keeps sufficient statistics
    __device__ __forceinline__ constexpr bool prior_fused(int l) const {
        return g_shape_of<Desc>.loc[l].family == JBUGS_FAM_NORMAL &&
               (g_shape_of<Desc>.loc[l].args[0].kind == JBUGS_ARG_GVAL || g_shape_of<Desc>.loc[l].args[0].kind == JBUGS_ARG_CONST) &&
               (g_shape_of<Desc>.loc[l].args[1].kind == JBUGS_ARG_GVAL || g_shape_of<Desc>.loc[l].args[1].kind == JBUGS_ARG_CONST);
    }
    // ... and its mean is the literal 0 (the shape records it: LocalS::mu_zero, set by build_plate from the plan)
    __device__ __forceinline__ constexpr bool prior_mu_zero(int l) const { return g_shape_of<Desc>.loc[l].mu_zero != 0; }
    __device__ __forceinline__ explicit StaticSpec(const PlateShape&) {}
    __device__ __forceinline__ constexpr int kg() const { return g_shape_of<Desc>.kg; }
    __device__ __forceinline__ constexpr int nl() const { return g_shape_of<Desc>.nl; }
    __device__ __forceinline__ constexpr int has_off() const { return g_shape_of<Desc>.has_off; }
    __device__ __forceinline__ constexpr int n_gref() const { return g_shape_of<Desc>.n_gref; }
    __device__ __forceinline__ constexpr int family() const { return g_shape_of<Desc>.family; }
    __device__ __forceinline__ constexpr int eta_arg() const { return g_shape_of<Desc>.eta_arg; }
    __device__ __forceinline__ constexpr int has_obs() const { return g_shape_of<Desc>.has_obs; }
    __device__ __forceinline__ constexpr int n_link() const { return g_shape_of<Desc>.n_link; }
    __device__ __forceinline__ constexpr int link(int q) const { return g_shape_of<Desc>.link[q]; }
    __device__ __forceinline__ constexpr ArgS cov(int t) const { return g_shape_of<Desc>.cov[t]; }
    __device__ __forceinline__ constexpr ArgS y() const { return g_shape_of<Desc>.y; }
    __device__ __forceinline__ constexpr ArgS aux(int q) const { return g_shape_of<Desc>.aux[q]; }
    __device__ __forceinline__ constexpr bool has_w() const {  // (only a run-time generated description has a weight)
        return g_shape_of<Desc>.aux[2].kind == JBUGS_ARG_DATA_I || g_shape_of<Desc>.aux[2].kind == JBUGS_ARG_DATA_IJ;
    }
    __device__ __forceinline__ constexpr int loc_level(int l) const { return g_shape_of<Desc>.loc[l].level; }
    __device__ __forceinline__ constexpr int loc_transform(int l) const { return g_shape_of<Desc>.loc[l].transform; }
    __device__ __forceinline__ constexpr int loc_family(int l) const { return g_shape_of<Desc>.loc[l].family; }
    __device__ __forceinline__ constexpr int loc_has_idx(int l) const { return g_shape_of<Desc>.loc[l].has_idx; }
    __device__ __forceinline__ constexpr ArgS loc_arg(int l, int q) const { return g_shape_of<Desc>.loc[l].args[q]; }
};

// ------------------------------------------------------------------------------------------------
// register-array helpers: every index is a compile-time constant after unrolling, so the arrays
// stay in registers even when the *selected* element is a run-time value.
template <int N>
__device__ __forceinline__ double pick(const double (&a)[N], int id) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) v = (k == id) ? a[k] : v;
    return v;
}
template <int N>
__device__ __forceinline__ void bump(double (&a)[N], int id, double d) {
#pragma unroll
    for (int k = 0; k < N; ++k) a[k] += (k == id) ? d : 0.0;
}
// the same over the first M entries only (M = S::NG_MAX: the entries an instantiation uses)
template <int M, int N>
__device__ __forceinline__ double pick_n(const double (&a)[N], int id) {
    static_assert(M <= N, "pick_n");
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) v = (k == id) ? a[k] : v;
    return v;
}
template <int M, int N>
__device__ __forceinline__ void bump_n(double (&a)[N], int id, double d) {
    static_assert(M <= N, "bump_n");
#pragma unroll
    for (int k = 0; k < M; ++k) a[k] += (k == id) ? d : 0.0;
}

struct ChainState {
    double gv[MAX_GREF];  // referenced global values
    double ga[MAX_GREF];  // their adjoints
    double lx[MAX_LOC];   // constrained local values
    double la[MAX_LOC];   // their adjoints
    // dgamma(a, b) priors of locals whose arguments do not vary over the plate: lgamma(a), digamma(a), log(b) are
    // computed once per chain instead of once per group
    double pre_lg[MAX_LOC], pre_dg[MAX_LOC], pre_lb[MAX_LOC];
    int sat;  // logit paths: an |eta| of the current block reached the range where the reference's logistic saturates
};

__device__ __forceinline__ bool arg_invariant(ArgS s) {
    return s.kind == JBUGS_ARG_GVAL || s.kind == JBUGS_ARG_CONST || s.kind == JBUGS_ARG_ONE;
}

// where the staged data of the current chunk lives (shared memory)
struct StageView {
    const double* chunk;  // [per_group][STAGE_G]
    const double* jreg;   // J region
};

// g = group index inside the staged chunk, j = inner index.  SS: take the stream offset from the (compile-time) shape
template <bool SS = false, int NG = MAX_GREF>
__device__ __forceinline__ double arg_value(ArgS s, const ArgV& v, const ChainState& st, const StageView& sv, int g, int j) {
    const int soff = SS ? s.soff : v.soff;
    switch (s.kind) {
    case JBUGS_ARG_CONST: return v.value;
    case JBUGS_ARG_ONE: return 1.0;
    case JBUGS_ARG_GVAL: return pick_n<NG>(st.gv, s.id);
    case JBUGS_ARG_LOCAL: return pick(st.lx, s.id);
    case JBUGS_ARG_DATA_I: return sv.chunk[soff + g];
    case JBUGS_ARG_DATA_J: return sv.jreg[soff + j];
    case JBUGS_ARG_DATA_IJ: return sv.chunk[soff + j * STAGE_G + g];
    default: return 0.0;
    }
}

template <int NG = MAX_GREF>
__device__ __forceinline__ void arg_adjoint(ArgS s, ChainState& st, double d) {
    if (s.kind == JBUGS_ARG_GVAL) bump_n<NG>(st.ga, s.id, d);
    else if (s.kind == JBUGS_ARG_LOCAL) bump(st.la, s.id, d);
}

// ------------------------------------------------------------------------------------------------
// prologue: constrained values of the globals and the gtape, per chain -> gvals[k][c]
struct GlobalsParams {
    int n_globals, n_gtape, transformed;
    jbugs_global g[MAX_GVALS];
    // priors whose arguments are literal constants: the argument-only part of the log-density is computed once on
    // the host (cst); fast = family code + 1 when that shortcut applies, else 0
    double cst[MAX_GVALS];
    unsigned char fast[MAX_GVALS];
};
struct GTapeParams {
    jbugs_gop op[MAX_GVALS];
};

__device__ __forceinline__ double garg(const jbugs_arg& a, const double* gv) {
    return a.kind == JBUGS_ARG_GVAL ? gv[a.id] : (a.kind == JBUGS_ARG_ONE ? 1.0 : a.value);
}

__device__ __forceinline__ bool gtape_forward(const jbugs_gop& o, const double* gv, double& v, double& d1, double& d2) {
    const double a = garg(o.a, gv);
    d2 = 0.0;
    if (o.op < 16) return unary_op((int)o.op, a, v, d1);
    const double b = garg(o.b, gv);
    switch (o.op) {
    case JBUGS_OP_ADD: v = a + b; d1 = 1.0; d2 = 1.0; return false;
    case JBUGS_OP_SUB: v = a - b; d1 = 1.0; d2 = -1.0; return false;
    case JBUGS_OP_MUL: v = a * b; d1 = b; d2 = a; return false;
    case JBUGS_OP_DIV: v = a / b; d1 = 1.0 / b; d2 = -a / (b * b); return false;
    default:  // POW
        v = pow(a, b); d1 = b * pow(a, b - 1.0); d2 = a > 0.0 ? v * log(a) : 0.0;
        return a < 0.0 && b != floor(b);
    }
}

// Thread-per-chain helper code runs with very few warps per SM, so a load whose result is consumed right away
// costs a full memory round trip each time (the warp issues in order).  `batched_load` issues eight independent
// loads before the first use; the index is clamped instead of predicated so that the loads stay unconditional.
template <typename AddrFn>
__device__ __forceinline__ void batched_load(double* __restrict__ dst, int n, AddrFn addr) {
#pragma unroll 1
    for (int k0 = 0; k0 < n; k0 += 8) {
        double t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = *addr(min(k0 + u, n - 1));
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (k0 + u < n) dst[k0 + u] = t[u];
    }
}

// constrained globals + gtape of chain c -> gvals[k][c].  ONE_LAUNCH: part of fused_small_kernel, where every CTA of a
// chain block runs it for its own chains (identical values to identical addresses) and a DomainError is OR-ed into
// the flag instead of stored, because the other CTAs' plate tiles raise the same flag concurrently.
template <bool ONE_LAUNCH>
__device__ __forceinline__ void prologue_body(const GlobalsParams& G, const GTapeParams& TP, const double* __restrict__ theta,
                                              long long ld, long long c, double* __restrict__ gvals, long long ldg,
                                              int* __restrict__ flags) {
    double gv[MAX_GVALS];
    bool bad = false;
    batched_load(gv, G.n_globals, [&](int h) { return theta + G.g[h].theta_idx * ld + c; });
    for (int h = 0; h < G.n_globals; ++h)
        gv[h] = transform_eval((int)G.g[h].transform, G.g[h].lb, G.g[h].ub, gv[h]).x;
    for (int k = 0; k < G.n_gtape; ++k) {
        double v, d1, d2;
        bad |= gtape_forward(TP.op[k], gv, v, d1, d2);
        gv[G.n_globals + k] = v;
    }
    for (int k = 0; k < G.n_globals + G.n_gtape; ++k) gvals[k * ldg + c] = gv[k];
    if (ONE_LAUNCH) {
        if (bad) atomicOr(flags + c, 1);
    } else {
        flags[c] = bad ? 1 : 0;
    }
}

#ifdef JBUGS_HOST_TU  // the non-template kernels are compiled once, in the host translation unit
__global__ void __launch_bounds__(BLOCK)
prologue_kernel(const __grid_constant__ GlobalsParams G, const __grid_constant__ GTapeParams TP,
                const double* __restrict__ theta, long long ld, long long C, double* __restrict__ gvals,
                long long ldg, int* __restrict__ flags, int* __restrict__ any_bad) {
    pdl_launch_dependents();  // the plate kernel may get going (it waits before it reads gvals)
    pdl_wait();               // theta may come from the kernel before (a leapfrog update); gvals / flags are still being read
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (c == 0) *any_bad = 0;
    prologue_body<false>(G, TP, theta, ld, c, gvals, ldg, flags);
}
#endif  // JBUGS_HOST_TU

}  // namespace jbugs

#include "jbugs_plate.cuh"

namespace jbugs {

// ------------------------------------------------------------------------------------------------
// finalize: add the tiles in ascending order, the priors of the globals, reverse the gtape and the
// bijectors; write logp, status and the gradient rows of the globals.
struct PlateFinal {
    int n_tiles, n_gref;
    long long offset;  // into the partial buffer (in units of ldg rows)
    int gref[MAX_GREF];
    double obs_const;
    int replicated;  // observation-sharded runs: the plate is evaluated identically on every rank -> counted once
};
struct FinalParams {
    int n_plates;
    long long poison_rows;  // > 0: this kernel also writes the NaN gradient of a flagged chain (small models); 0: poison_kernel does
    PlateFinal p[8];
};

// mode 0: single GPU, everything in one pass.
// mode 1: observation-sharded, first half: write this rank's reduced partials red[0] = logp part,
//         red[1 + k] = adjoint of global value k, red[1 + NG] = DomainError flag; the host all-reduces `red`.
// mode 2: second half: read the all-reduced `red`, add the priors of the globals ONCE, finish.
// ONE_LAUNCH: the last CTA of a chain block inside fused_small_kernel; partials and flags were written by other CTAs
// of the same launch, so they are read past L1 (__ldcg).
template <bool ONE_LAUNCH>
__device__ __forceinline__ void finalize_body(const GlobalsParams& G, const GTapeParams& TP, const FinalParams& F,
                                              const double* __restrict__ theta, double* __restrict__ grad, long long ld,
                                              long long c, const double* __restrict__ gvals, long long ldg,
                                              const double* partial, int* flags, double* __restrict__ logp_out,
                                              int* __restrict__ status_out, int* __restrict__ any_bad, int want_grad,
                                              int mode, double* __restrict__ red) {
    const int H = G.n_globals, NG = G.n_globals + G.n_gtape;
    double gv[MAX_GVALS], ga[MAX_GVALS], th[MAX_GVALS];
    batched_load(gv, NG, [&](int k) { return gvals + k * ldg + c; });
    if (mode != 1) batched_load(th, H, [&](int h) { return theta + G.g[h].theta_idx * ld + c; });
    for (int k = 0; k < NG; ++k) ga[k] = 0.0;
    double lp = 0.0;
    bool bad = (ONE_LAUNCH ? __ldcg(flags + c) : flags[c]) != 0;
    if (mode == 2) {
        lp = red[c];
        batched_load(ga, NG, [&](int k) { return red + (long long)(1 + k) * ldg + c; });
        bad = red[(long long)(1 + NG) * ldg + c] != 0.0;
    }
    for (int p = 0; p < F.n_plates; ++p) {
        const PlateFinal& pf = F.p[p];
        // sharded plates enter the all-reduce (mode 1); replicated ones are added once, after it (mode 2)
        if ((mode == 1 && pf.replicated) || (mode == 2 && !pf.replicated)) continue;
        const long long stride = (long long)(1 + pf.n_gref) * ldg;
        const double* base = partial + pf.offset * ldg + c;
        const int nt = pf.n_tiles;
        // row 0 = log-density partials, row 1 + k = adjoint of the k-th referenced global value; tiles are added
        // in ascending order, sixteen loads in flight at a time
#pragma unroll 1
        for (int row = 0; row <= pf.n_gref; ++row) {
            const double* src = base + (long long)row * ldg;
            double s = 0.0;
#pragma unroll 1
            for (int t0 = 0; t0 < nt; t0 += 16) {
                double v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) v[u] = ONE_LAUNCH ? __ldcg(src + min(t0 + u, nt - 1) * stride) : src[min(t0 + u, nt - 1) * stride];
#pragma unroll
                for (int u = 0; u < 16; ++u)
                    if (t0 + u < nt) s += v[u];
            }
            if (row == 0) lp += s + pf.obs_const;
            else ga[pf.gref[row - 1]] += s;
        }
    }
    if (mode == 1) {
        red[c] = lp;
        for (int k = 0; k < NG; ++k) red[(long long)(1 + k) * ldg + c] = ga[k];
        red[(long long)(1 + NG) * ldg + c] = bad ? 1.0 : 0.0;
        return;
    }
    if (!ONE_LAUNCH) flags[c] = bad ? 1 : 0;  // (mode 2: the flag now reflects every rank's shard)
    for (int h = 0; h < H; ++h) {
        const jbugs_global& g = G.g[h];
        if (G.fast[h] == JBUGS_FAM_NORMAL + 1) {  // dnorm(m, t), literals: cst = (log t - log 2 pi) / 2
            const double r = gv[h] - g.args[0].value;
            lp += fma(-0.5 * g.args[1].value * r, r, G.cst[h]);
            ga[h] -= g.args[1].value * r;
            continue;
        }
        if (G.fast[h] == JBUGS_FAM_GAMMA + 1) {  // dgamma(a, b), literals: cst = a log b - lgamma(a)
            const double x = gv[h], a = g.args[0].value, b = g.args[1].value;
            lp += x < 0.0 ? -dev_inf() : (G.cst[h] + xlogy(a - 1.0, x) - b * x);
            ga[h] += (a - 1.0) / x - b;
            continue;
        }
        FamOut o;
        bad |= fam_eval((int)g.family, gv[h], garg(g.args[0], gv), garg(g.args[1], gv), o, garg(g.args[2], gv));
        lp += o.lp + G.cst[h];  // (cst: 0, or minus the log mass of a truncated prior with literal arguments)
        ga[h] += o.dx;
        if (g.args[0].kind == JBUGS_ARG_GVAL) ga[g.args[0].id] += o.da0;
        if (g.args[1].kind == JBUGS_ARG_GVAL) ga[g.args[1].id] += o.da1;
    }
    // reverse sweep over the gtape (recompute the local partials)
    for (int k = G.n_gtape - 1; k >= 0; --k) {
        double v, d1, d2;
        gtape_forward(TP.op[k], gv, v, d1, d2);
        const double ad = ga[H + k];
        if (TP.op[k].a.kind == JBUGS_ARG_GVAL) ga[TP.op[k].a.id] += ad * d1;
        if (TP.op[k].op >= 16 && TP.op[k].b.kind == JBUGS_ARG_GVAL) ga[TP.op[k].b.id] += ad * d2;
    }
    for (int h = 0; h < H; ++h) {
        const jbugs_global& g = G.g[h];
        const TrOut t = transform_eval((int)g.transform, g.lb, g.ub, th[h]);
        lp += t.lj;
        if (want_grad) grad[g.theta_idx * ld + c] = ga[h] * t.dxdy + t.dljdy;
    }
    if (ONE_LAUNCH) flags[c] = 0;  // (the one-launch kernel owns its flag row: clean for the next evaluation)
    else if (bad) flags[c] = 1;
    logp_out[c] = bad ? -dev_inf() : lp;
    if (status_out) status_out[c] = bad ? CHAIN_DOMAIN : CHAIN_OK;
    if (bad) atomicOr(any_bad, 1);
    if (bad && want_grad)  // DomainError -> NaN gradient (logdensityproblems.jl:226-229); every plate kernel has finished
        for (long long d = 0; d < F.poison_rows; ++d) grad[d * ld + c] = dev_nan();
}

// ------------------------------------------------------------------------------------------------
// One launch per evaluation for small single-plate models (BASELINE config 5: a few hundred nodes x thousands of
// chains, latency-bound).  Three dependent launches -- prologue, plate, finalize -- cost more in launch latency and
// drain/fill gaps than the arithmetic itself.  Here every CTA (chain block x group tile) recomputes the prologue for
// its own chains, evaluates its tile, publishes its partials and takes a ticket per chain block; the CTA that draws the
// last ticket adds the tiles in ascending order (the same fixed order as finalize_kernel) and finishes the chains.
struct FusedParams {
    GlobalsParams G;
    GTapeParams TP;
    FinalParams F;
    PlateParams P;
};

template <class S, int BLK>
__global__ void __launch_bounds__(BLK, (S::MIN_BLOCKS * (128 / BLK) > 24 ? 24 : S::MIN_BLOCKS * (128 / BLK)))
fused_small_kernel(const __grid_constant__ FusedParams Q, const double* __restrict__ theta, double* __restrict__ grad,
                   long long ld, long long C, double* __restrict__ gvals, long long ldg, double* __restrict__ partial,
                   int* __restrict__ flags, double* __restrict__ logp_out, int* __restrict__ status_out,
                   int* __restrict__ any_bad, int want_grad, unsigned* __restrict__ tickets) {
    pdl_wait();  // theta may come from the kernel before (a leapfrog update)
    const long long c = (long long)blockIdx.x * BLK + threadIdx.x;
    if (c < C) prologue_body<true>(Q.G, Q.TP, theta, ld, c, gvals, ldg, flags);   // this thread reads back its own writes only
    plate_body<S, BLK, false, true>(Q.P, theta, grad, ld, C, gvals, ldg, partial, flags, want_grad, blockIdx.y);
    __shared__ unsigned last;
    __threadfence();  // partials (and a raised flag) are visible device-wide before the ticket is taken
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(tickets + blockIdx.x, 1u) == gridDim.y - 1 ? 1u : 0u;
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x == 0) tickets[blockIdx.x] = 0u;  // ready for the next evaluation
    if (c < C)
        finalize_body<true>(Q.G, Q.TP, Q.F, theta, grad, ld, c, gvals, ldg, partial, flags, logp_out, status_out, any_bad, want_grad, 0, nullptr);
}

#ifdef JBUGS_HOST_TU
__global__ void __launch_bounds__(BLOCK)
finalize_kernel(const __grid_constant__ GlobalsParams G, const __grid_constant__ GTapeParams TP,
                const __grid_constant__ FinalParams F, const double* __restrict__ theta, double* __restrict__ grad,
                long long ld, long long C, const double* __restrict__ gvals, long long ldg,
                const double* __restrict__ partial, int* __restrict__ flags, double* __restrict__ logp_out,
                int* __restrict__ status_out, int* __restrict__ any_bad, int want_grad, int mode,
                double* __restrict__ red) {
    pdl_wait();  // partials, gvals and flags come from the kernels before this one
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    finalize_body<false>(G, TP, F, theta, grad, ld, c, gvals, ldg, partial, flags, logp_out, status_out, any_bad, want_grad, mode, red);
}

// First level of the fixed-order tile reduction, used when a plate has many tiles: super-tile r adds the tiles
// [r * n / R, (r + 1) * n / R) in ascending order; finalize_kernel then adds the R super-tiles in ascending order.
// Same layout in and out ([tile][row][ldg]), so finalize_kernel does not care which level it reads.
constexpr int REDUCE_R = 64;
__global__ void __launch_bounds__(BLOCK)
reduce_tiles_kernel(const double* __restrict__ partial, double* __restrict__ out, int n_tiles, int rows, long long C,
                    long long ldg) {
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C) return;
    const int row = blockIdx.y, r = blockIdx.z;
    const int t0 = (int)((long long)r * n_tiles / REDUCE_R), t1 = (int)((long long)(r + 1) * n_tiles / REDUCE_R);
    const double* src = partial + (long long)row * ldg + c;
    const long long stride = (long long)rows * ldg;
    double s = 0.0;
    for (int t = t0; t < t1; ++t) s += src[t * stride];
    out[((long long)r * rows + row) * ldg + c] = s;
}

// DomainError -> NaN gradient for the whole chain (logdensityproblems.jl:226-229). Only does work when a
// chain of this batch was flagged.
__global__ void __launch_bounds__(BLOCK)
poison_kernel(double* __restrict__ grad, long long ld, long long C, long long D, const int* __restrict__ flags,
              const int* __restrict__ any_bad) {
    if (*any_bad == 0) return;
    const long long c = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (c >= C || flags[c] == 0) return;
    for (long long d = blockIdx.y; d < D; d += gridDim.y) grad[d * ld + c] = dev_nan();
}

// ------------------------------------------------------------------------------------------------
// layout conversion (PARAM_FAST <-> CHAIN_FAST), 32x32 tiles through shared memory
__global__ void __launch_bounds__(256)
transpose_kernel(const double* __restrict__ src, long long lds, double* __restrict__ dst, long long ldd,
                 long long rows, long long cols) {
    // src is rows x cols with leading dimension lds (element (r, q) at r * lds + q); dst(q, r) = src(r, q)
    __shared__ double tile[32][33];
    const long long r0 = (long long)blockIdx.y * 32, q0 = (long long)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8) {
        const long long r = r0 + k, q = q0 + tx;
        if (r < rows && q < cols) tile[k][tx] = src[r * lds + q];
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const long long q = q0 + k, r = r0 + tx;
        if (r < rows && q < cols) dst[q * ldd + r] = tile[tx][k];
    }
}
#endif  // JBUGS_HOST_TU

}  // namespace jbugs
