// jbugs_cuda.cu -- host side of libjbugs_cuda: plan handling, device buffers, kernel selection and the
// C ABI declared in include/jbugs_cuda.h.  No torch types, no CPU fallback.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/jbugs_cuda.h"
#include "../../include/jbugs_trunc.h"
#define JBUGS_HOST_TU
#include "jbugs_kernels.cuh"
#include "jbugs_obsconst.cuh"
#include "jbugs_dense.cuh"
#include "jbugs_hmc.cuh"
#include "jbugs_nuts.cuh"
#include "jbugs_gq.cuh"
#define JBUGS_SPEC_TABLE
#include "jbugs_specs.cuh"
static_assert(jbugs::CHAIN_OK == JBUGS_CHAIN_OK && jbugs::CHAIN_DOMAIN == JBUGS_CHAIN_DOMAIN, "status codes of the kernels and of the ABI differ");

using namespace jbugs;

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(expr)                                                                                       \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return fail(JBUGS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------------ NCCL (dlopen'ed)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
    if (g_nccl.h) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) return fail(JBUGS_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", dlerror());
    g_nccl.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(g_nccl.h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(g_nccl.h, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(g_nccl.h, "ncclCommDestroy");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(g_nccl.h, "ncclAllReduce");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce)
        return fail(JBUGS_ERR_NCCL, "libnccl lacks a required symbol");
    return 0;
}

// ------------------------------------------------------------------------------------------------ plan object
struct PlateRt {
    PlateParams pp;        // device-side description (pointers resolved)
    int variant = 0;       // index into the spec table (0 = dynamic interpreter)
    long long n_tiles = 1; // for the current capacity
    long long partial_offset = 0, partial2_offset = 0;
    double obs_const = 0.0;
    bool const_dirty = true;  // obs_const must be recomputed (new data / shard / kernel variant)
    size_t smem_bytes = 0;    // dynamic shared memory of the plate kernel (J region + two stages)
    // run-time specialisation (jbugs_plan_specialize): a kernel compiled for exactly this plate's shape
    struct JitKernels* jit = nullptr;  // (jbugs_jit.inc; owned by the plan)
    int jit_fused = FUSED_NONE;
    bool jit_valid = true;    // false: the data are outside the fused formulas' domain -> interpreter
    XOpDev* d_xtape = nullptr;  // expression plates: the resolved tape in device memory
    std::vector<XOpDev> xtape_host;  // ... and on the host (the source the specialiser writes the tape's code from)
};
static cudaError_t jit_launch(JitKernels& jk, JBUGS_LAUNCH_PARAMS);
static void jit_free(JitKernels* jk);
static bool jit_is_best(const JitKernels* jk);
struct JitJob;
static void jit_job_free(JitJob* j);
static int jit_poll(jbugs_plan* p);

struct jbugs_plan_s {
    int device = 0;
    int num_sms = 148;
    jbugs_plan_header h{};
    std::vector<jbugs_global> globals;
    std::vector<jbugs_gop> gtape;
    std::vector<jbugs_data_desc> data;
    std::vector<jbugs_plate> plates;
    std::vector<jbugs_local> locals;
    std::vector<jbugs_term> terms;
    std::vector<jbugs_dense> dense;
    std::vector<jbugs_xop> xops;
    // dense predictors: residual matrix R [N x ldr], split-K partials of X^T R, per-chain logp partial rows
    double* d_R = nullptr;
    double* d_Gp = nullptr;
    size_t R_bytes = 0, Gp_bytes = 0;
    long long dense_partial_offset = 0;  // rows (of ldg) inside d_partial
    int dense_row_groups = 0, dense_splits = 0;
    double dense_const = 0.0;
    bool dense_const_dirty = true;
    long long dense_i0 = 0, dense_i1 = -1;  // this rank's row range of X (jbugs_plan_set_dense_shard); -1 = all rows
    std::vector<void*> d_slots;
    std::vector<char> uploaded;
    std::vector<PlateRt> rt;
    JitJob* jit_job = nullptr;  // asynchronous run-time specialisation under way (jbugs_jit.inc)
    std::vector<struct JitKernels*> jit_retired;  // quick variants replaced by better ones (unloaded with the plan)
    std::string jit_cache_dir;  // cache directory of the last specialise call (the background upgrade uses the same)
    bool jit_cache_dir_set = false;
    std::vector<long long> shard_i0, shard_i1;
    std::vector<char> shard_set;  // jbugs_plan_set_shard was called for the plate (else it is replicated on every rank)
    GlobalsParams gp{};
    GTapeParams tp{};
    // workspaces sized for cap_C chains
    long long cap_C = 0, ldg = 0;
    double* d_gvals = nullptr;
    double* d_partial = nullptr;
    double* d_partial2 = nullptr;  // super-tile level of the fixed-order reduction
    size_t partial_bytes = 0, partial2_bytes = 0, gvals_bytes = 0, red_bytes = 0, flags_bytes = 0;
    double* d_red = nullptr;   // [1 + NG][ldg] reduced partials (multi-GPU path)
    int* d_flags = nullptr;
    int* d_any_bad = nullptr;
    // one-launch evaluation of small single-plate models (fused_small_kernel): its own DomainError flags and one ticket
    // per chain block, both zero between evaluations (the kernel cleans up after itself)
    int* d_fflags = nullptr;
    unsigned* d_tickets = nullptr;
    size_t fflags_bytes = 0, tickets_bytes = 0;
    int one_launch = 1;  // option "one_launch"
    // pipelined host path: two staging slots, copy-in / copy-out streams, per-slot events
    struct Slot {
        double *theta = nullptr, *grad = nullptr, *theta_t = nullptr, *grad_t = nullptr, *logp = nullptr;
        int* status = nullptr;
        cudaEvent_t in_done = nullptr, comp_done = nullptr, out_done = nullptr;
    } slot[2];
    long long slot_C = 0;
    long long chunk_seq = 0;  // staged chunks issued so far (slots alternate across calls)
    bool slot_t = false;
    cudaStream_t st_in = nullptr, st_out = nullptr;
    // library-owned batch
    double *b_theta = nullptr, *b_grad = nullptr, *b_logp = nullptr;
    int* b_status = nullptr;
    // options / stats
    int force_dynamic = 0;
    long long tile_override = 0;
    long long host_chunk = 0;  // chains per pipelined chunk of the host path (0 = automatic)
    // stage plate data with cp.async.bulk + mbarrier (TMA) instead of cp.async.  Off by default: measured on B200
    // (rats 10^6 x 5 x 4096) the two-stage TMA ring is 27.8 ms against 20.3 ms for cp.async -- the ~1.5 us
    // issue-to-complete latency of a tiny bulk copy is longer than one 32-group chunk of compute, and the data
    // streams are < 1 % of the traffic anyway.  `jbugs_plan_set_option(plan, "tma", 1)` selects the TMA kernels.
    int use_tma = 0;
    int use_pdl = 1;  // programmatic dependent launch of the plate kernels and finalize (option "pdl")
    int parts = 0;    // 1: evaluate the prior part only (parameters' densities + log-Jacobians; observations skipped) --
                      // the second pass of jbugs_eval_batch_tempered
    long long launches = 0;
    // bumped whenever anything a captured launch sequence bakes in may have changed (tiling, workspace pointers,
    // data-only constants): the HMC trajectory graph is rebuilt when it sees a new epoch
    unsigned long long epoch = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_last[4] = {nullptr, nullptr, nullptr, nullptr};  // the events the most recent evaluation recorded
    float last_plate_ms = 0.f, last_total_ms = 0.f;
    bool events_recorded = false;
    bool timing = false;  // record per-phase CUDA events (jbugs_last_timing); off by default: four extra stream ops per call
    // "timing" = 2: every evaluation records into its own four events of a ring (jbugs_timing_history reads the
    // plate-kernel and total time of each call since the option was set), so that a benchmark can report the mean and
    // the maximum of the dominant kernel over exactly the steps of its timed region
    std::vector<cudaEvent_t> ring;  // 4 events per entry
    long long ring_n = 0;           // evaluations recorded since the ring was (re)started
    static constexpr int RING = 1024;
    // comm
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
};

static void free_slots(jbugs_plan* p);

static int set_device(const jbugs_plan* p) {
    CU(cudaSetDevice(p->device));
    return 0;
}

// ------------------------------------------------------------------------------------------------ plate lowering to device params
struct StreamBuilder {
    PlateVals* V;
    long long T;
    int chunk_acc = 0;  // doubles per group so far
    int j_acc = 0;      // doubles in the J region so far
    int add(const double* ptr, int kind) {  // returns soff, or -1 when the table is full
        for (int k = 0; k < V->n_streams; ++k)
            if (V->streams[k].ptr == ptr && V->streams[k].kind == kind) return V->streams[k].soff;
        if (V->n_streams >= MAX_STREAMS) return -1;
        Stream& s = V->streams[V->n_streams++];
        s.ptr = ptr;
        s.kind = kind;
        if (kind == JBUGS_ARG_DATA_J) {
            s.soff = j_acc;
            j_acc += (int)T;
        } else {
            s.soff = chunk_acc * STAGE_G;
            chunk_acc += kind == JBUGS_ARG_DATA_IJ ? (int)T : 1;
        }
        return s.soff;
    }
};

static int resolve_arg(const jbugs_plan* p, const jbugs_arg& a, ArgS& s, ArgV& v, std::vector<int>& gref, int max_local,
                       StreamBuilder& sb) {
    s.kind = (int)a.kind;
    s.id = 0;
    s.soff = 0;
    v.ptr = nullptr;
    v.value = a.value;
    v.soff = 0;
    v.pad = 0;
    switch (a.kind) {
    case JBUGS_ARG_CONST:
    case JBUGS_ARG_ONE: return 0;
    case JBUGS_ARG_GVAL: {
        if (a.id < 0 || a.id >= (int)(p->h.n_globals + p->h.n_gtape)) return fail(JBUGS_ERR_INVALID, "gval id out of range");
        auto it = std::find(gref.begin(), gref.end(), a.id);
        if (it == gref.end()) {
            gref.push_back(a.id);
            s.id = (int)gref.size() - 1;
        } else {
            s.id = (int)(it - gref.begin());
        }
        return 0;
    }
    case JBUGS_ARG_LOCAL:
        if (a.id < 0 || a.id >= max_local) return fail(JBUGS_ERR_INVALID, "local id out of range");
        s.id = a.id;
        return 0;
    case JBUGS_ARG_DATA_I:
    case JBUGS_ARG_DATA_J:
    case JBUGS_ARG_DATA_IJ:
        if (a.id < 0 || a.id >= (int)p->h.n_data) return fail(JBUGS_ERR_INVALID, "data slot out of range");
        if (p->data[a.id].dtype != 0) return fail(JBUGS_ERR_INVALID, "data slot %d is not float64", a.id);
        v.ptr = (const double*)p->d_slots[a.id];
        v.soff = sb.add(v.ptr, (int)a.kind);
        if (v.soff < 0) return fail(JBUGS_ERR_UNSUPPORTED, "a plate uses more than %d distinct data streams", MAX_STREAMS);
        s.soff = v.soff;  // also part of the shape (static specs with a static inner extent bake the layout in)
        return 0;
    default: return fail(JBUGS_ERR_INVALID, "unknown argument kind %u", a.kind);
    }
}

static int check_len(const jbugs_plan* p, const jbugs_arg& a, long long N, long long T, const char* what) {
    long long need = 0;
    if (a.kind == JBUGS_ARG_DATA_I) need = N;
    else if (a.kind == JBUGS_ARG_DATA_J) need = T;
    else if (a.kind == JBUGS_ARG_DATA_IJ) need = N * T;
    else return 0;
    if (p->data[a.id].len < need) return fail(JBUGS_ERR_INVALID, "%s: data slot %d has %lld elements, plate needs %lld", what, a.id, (long long)p->data[a.id].len, need);
    return 0;
}

// the part of a plate's device description shared by linear-predictor plates and expression plates: the locals with
// their priors and theta-slot maps, the gref table, staging sizes, the kernel variant
static int pick_variant(const jbugs_plan* p, const PlateRt& r, long long T);
static int build_locals_and_finish(jbugs_plan* p, int k, std::vector<int>& gref, StreamBuilder& sb) {
    const jbugs_plate& pl = p->plates[k];
    PlateRt& r = p->rt[k];
    PlateParams& pp = r.pp;
    PlateShape& sh = pp.sh;
    PlateVals& V = pp.v;
    const jbugs_local* L = p->locals.data() + pl.first_local;
    int rc;
    for (int l = 0; l < (int)pl.n_locals; ++l) {
        LocalS& ls = sh.loc[l];
        LocalV& lv = V.loc[l];
        if (L[l].family >= JBUGS_FAM_COUNT) return fail(JBUGS_ERR_INVALID, "plate %d local %d: unknown family", k, l);
        ls.level = (int)L[l].level; ls.transform = (int)L[l].transform; ls.family = (int)L[l].family;
        ls.has_idx = L[l].idx_slot >= 0;
        lv.base = L[l].theta_base; lv.si = L[l].stride_i; lv.sj = L[l].stride_j;
        lv.lb = L[l].lb; lv.ub = L[l].ub; lv.idx = nullptr;
        if (ls.has_idx) {
            if (L[l].idx_slot >= (int)p->h.n_data || p->data[L[l].idx_slot].dtype != 1)
                return fail(JBUGS_ERR_INVALID, "plate %d local %d: index slot must be int64", k, l);
            const long long need = ls.level == JBUGS_LEVEL_OBS ? pl.N * pl.T : pl.N;
            if (p->data[L[l].idx_slot].len < need) return fail(JBUGS_ERR_INVALID, "plate %d local %d: index array too short", k, l);
            lv.idx = (const long long*)p->d_slots[L[l].idx_slot];
        } else {
            // every addressed slot must lie inside [0, D)
            const long long jmax = ls.level == JBUGS_LEVEL_OBS ? pl.T - 1 : 0;
            const long long c[4] = {lv.base, lv.base + lv.si * (pl.N - 1), lv.base + lv.sj * jmax,
                                    lv.base + lv.si * (pl.N - 1) + lv.sj * jmax};
            for (long long x : c)
                if (pl.N > 0 && (x < 0 || x >= p->h.D)) return fail(JBUGS_ERR_INVALID, "plate %d local %d: theta slot out of range", k, l);
        }
        for (int q = 0; q < 2; ++q) {
            if ((rc = resolve_arg(p, L[l].args[q], ls.args[q], lv.args[q], gref, (int)pl.n_locals, sb))) return rc;
            if (ls.level == JBUGS_LEVEL_GROUP && (L[l].args[q].kind == JBUGS_ARG_DATA_J || L[l].args[q].kind == JBUGS_ARG_DATA_IJ))
                return fail(JBUGS_ERR_INVALID, "plate %d local %d: group-level prior indexed by j", k, l);
            if ((rc = check_len(p, L[l].args[q], pl.N, pl.T, "prior argument"))) return rc;
        }
        ls.mu_zero = (L[l].args[0].kind == JBUGS_ARG_CONST && L[l].args[0].value == 0.0) ? 1 : 0;
        lv.args[2].value = L[l].args[2].kind == JBUGS_ARG_ONE ? 1.0 : L[l].args[2].value;  // dt's constant degrees of freedom
        if ((L[l].family == JBUGS_FAM_T || L[l].family == JBUGS_FAM_BETABIN) && L[l].args[2].kind != JBUGS_ARG_CONST && L[l].args[2].kind != JBUGS_ARG_ONE)
            return fail(JBUGS_ERR_UNSUPPORTED, "plate %d local %d: the degrees of freedom of dt must be a constant", k, l);
    }
    // expression plates (XSpec, compiled tape kernels) keep MAX_GREF global values per chain, the interpreter of
    // linear-predictor plates MAX_GREF_LINEAR (DynSpec::NG_MAX)
    const int gref_cap = (p->plates[k].has_obs == 2 || p->plates[k].has_obs == 3) ? MAX_GREF : MAX_GREF_LINEAR;
    if ((int)gref.size() > gref_cap) return fail(JBUGS_ERR_UNSUPPORTED, "plate %d references more than %d global values", k, gref_cap);
    sh.n_gref = (int)gref.size();
    for (int q = 0; q < MAX_GREF; ++q) V.gref[q] = q < (int)gref.size() ? gref[q] : 0;
    V.i0 = p->shard_i0[k]; V.i1 = p->shard_i1[k];
    V.per_group = sb.chunk_acc;
    V.j_doubles = (sb.j_acc + 1) / 2 * 2;
    // TMA staging needs 16-byte aligned sources: slot bases are cudaMalloc'ed, so element offsets must be even --
    // the kernel checks the chunk start and length, here: the column stride N of rank-2 streams
    V.bulk_ok = p->use_tma ? 1 : 0;
    for (int q = 0; q < V.n_streams; ++q)
        if (V.streams[q].kind == JBUGS_ARG_DATA_IJ && (pl.N % 2) != 0) V.bulk_ok = 0;
    r.smem_bytes = (size_t)(V.j_doubles + 2 * V.per_group * STAGE_G) * sizeof(double);
    if (r.smem_bytes > 200 * 1024)
        return fail(JBUGS_ERR_UNSUPPORTED, "plate %d: %zu bytes of staged data per chunk exceed shared memory (inner extent too large)", k, r.smem_bytes);
    r.variant = pick_variant(p, r, pl.T);
    r.const_dirty = true;
    r.obs_const = 0.0;
    return 0;
}

// expression plate (has_obs == 2): the predictor is an SSA tape of the composed node functions; leaves are resolved
// like any other argument (data leaves become staged streams, global leaves plate-local gref indices)
static int build_xplate(jbugs_plan* p, int k) {
    const jbugs_plate& pl = p->plates[k];
    PlateRt& r = p->rt[k];
    PlateParams& pp = r.pp;
    PlateShape& sh = pp.sh;
    PlateVals& V = pp.v;
    const jbugs_xop* X = p->xops.data() + pl.first_term;
    const int nx = (int)pl.n_terms;
    if (nx < 1 || nx > JBUGS_MAX_XOPS) return fail(JBUGS_ERR_UNSUPPORTED, "plate %d: expression tape of %d ops (1..%d supported)", k, nx, JBUGS_MAX_XOPS);
    if (pl.n_link != 0) return fail(JBUGS_ERR_INVALID, "plate %d: an expression plate has no link tape", k);
    // has_obs == 3: a discrete latent per observation is summed out; eta_arg holds its number of states, link[0] the
    // tape slot of log p(z = k)
    const int K = pl.has_obs == 3 ? (int)pl.eta_arg : 0;
    if (pl.has_obs == 3 && (K < 1 || K > JBUGS_MAX_STATES || (int)pl.link[0] >= nx))
        return fail(JBUGS_ERR_INVALID, "plate %d: %d latent states (1..%d) / log-prior slot %u outside the tape", k, K, JBUGS_MAX_STATES, pl.link[0]);
    sh.n_states = K; sh.lp_slot = K > 0 ? (int)pl.link[0] : 0;
    sh.zobs_slot = K > 0 ? (int)pl.link[1] - 1 : -1;
    if (sh.zobs_slot >= nx || (sh.zobs_slot >= 0 && (X[sh.zobs_slot].op != JBUGS_OP_LEAF ||
            (X[sh.zobs_slot].leaf.kind != JBUGS_ARG_DATA_I && X[sh.zobs_slot].leaf.kind != JBUGS_ARG_DATA_IJ))))
        return fail(JBUGS_ERR_INVALID, "plate %d: the latent's observed-state slot must be a data leaf of the tape", k);
    sh.kg = 0; sh.nl = (int)pl.n_locals; sh.has_off = 0;
    sh.family = (int)pl.family; sh.eta_arg = 0; sh.has_obs = 1; sh.n_link = 0; sh.n_x = nx;
    V.N = pl.N; V.T = pl.T;
    StreamBuilder sb;
    sb.V = &V;
    sb.T = pl.T;
    std::vector<int> gref;
    const int n_gv = (int)(p->h.n_globals + p->h.n_gtape);
    // parameter vectors indexed by the inner index first: their T elements must sit in consecutive gref positions
    // (the same for a vector picked by a data index, JBUGS_OP_GVEC_AT: leaf.value elements from leaf.id)
    for (int q = 0; q < nx; ++q) {
        const bool by_j = X[q].op == JBUGS_OP_LEAF && X[q].leaf.kind == JBUGS_ARG_GVEC_J, by_idx = X[q].op == JBUGS_OP_GVEC_AT;
        if (by_j || by_idx) {
            const int base = X[q].leaf.id;
            const long long cnt = by_j ? pl.T : (long long)X[q].leaf.value;
            if (base < 0 || cnt < 1 || base + cnt > n_gv) return fail(JBUGS_ERR_INVALID, "plate %d: parameter vector out of range", k);
            if (std::find(gref.begin(), gref.end(), base) == gref.end()) {
                for (long long e = 0; e < cnt; ++e)
                    if (std::find(gref.begin(), gref.end(), base + (int)e) != gref.end())
                        return fail(JBUGS_ERR_UNSUPPORTED, "plate %d: overlapping parameter vectors in one tape", k);
                for (long long e = 0; e < cnt; ++e) gref.push_back(base + (int)e);
            }
        }
    }
    std::vector<XOpDev> tape((size_t)nx);
    int rc;
    for (int q = 0; q < nx; ++q) {
        XOpDev& o = tape[q];
        memset(&o, 0, sizeof o);
        o.op = (int)X[q].op;
        o.a = X[q].a; o.b = X[q].b; o.c = X[q].c;
        const bool unary = X[q].op < 16, binary = X[q].op >= 16 && X[q].op <= JBUGS_OP_POW;
        if (X[q].op == JBUGS_OP_LEAF) {
            const jbugs_arg& a = X[q].leaf;
            if (a.kind == JBUGS_ARG_GVEC_J) {
                o.kind = JBUGS_ARG_GVEC_J;
                o.id = (int)(std::find(gref.begin(), gref.end(), a.id) - gref.begin());
                continue;
            }
            if (a.kind == JBUGS_ARG_XVAL) return fail(JBUGS_ERR_INVALID, "plate %d: a tape leaf cannot be a tape value", k);
            ArgS s2; ArgV v2;
            if ((rc = resolve_arg(p, a, s2, v2, gref, (int)pl.n_locals, sb))) return rc;
            if ((rc = check_len(p, a, pl.N, pl.T, "tape leaf"))) return rc;
            o.kind = s2.kind; o.id = s2.id; o.soff = v2.soff; o.value = a.kind == JBUGS_ARG_ONE ? 1.0 : a.value;
        } else if (X[q].op == JBUGS_OP_SELECT) {
            if (o.a >= q || o.b >= q || o.c >= q) return fail(JBUGS_ERR_INVALID, "plate %d: tape op %d looks forward", k, q);
        } else if (X[q].op == JBUGS_OP_PICK_K) {
            if (K < 1 || (int)o.a + K > q) return fail(JBUGS_ERR_INVALID, "plate %d: tape op %d picks outside the tape or without a latent", k, q);
        } else if (X[q].op == JBUGS_OP_GVEC_AT) {
            const jbugs_xop& ix = X[o.a < q ? o.a : 0];
            if (o.a >= q || ix.op != JBUGS_OP_LEAF ||
                (ix.leaf.kind != JBUGS_ARG_DATA_I && ix.leaf.kind != JBUGS_ARG_DATA_J && ix.leaf.kind != JBUGS_ARG_DATA_IJ))
                return fail(JBUGS_ERR_INVALID, "plate %d: tape op %d needs a data leaf as its index", k, q);
            o.id = (int)(std::find(gref.begin(), gref.end(), X[q].leaf.id) - gref.begin());
            o.value = X[q].leaf.value;
        } else if (unary || binary) {
            if (o.a >= q || (binary && o.b >= q)) return fail(JBUGS_ERR_INVALID, "plate %d: tape op %d looks forward", k, q);
            if (unary && X[q].op > JBUGS_OP_NEG) return fail(JBUGS_ERR_INVALID, "plate %d: unknown tape op %u", k, X[q].op);
        } else {
            return fail(JBUGS_ERR_INVALID, "plate %d: unknown tape op %u", k, X[q].op);
        }
    }
    if ((rc = resolve_arg(p, pl.y, sh.y, V.y, gref, 0, sb))) return rc;
    if ((rc = check_len(p, pl.y, pl.N, pl.T, "y"))) return rc;
    for (int q = 0; q < 2; ++q) {
        if (pl.aux[q].kind == JBUGS_ARG_XVAL) {
            if (pl.aux[q].id < 0 || pl.aux[q].id >= nx) return fail(JBUGS_ERR_INVALID, "plate %d: family argument outside the tape", k);
            sh.aux[q] = ArgS{JBUGS_ARG_XVAL, pl.aux[q].id, 0};
            continue;
        }
        if ((rc = resolve_arg(p, pl.aux[q], sh.aux[q], V.aux[q], gref, (int)pl.n_locals, sb))) return rc;
        if ((rc = check_len(p, pl.aux[q], pl.N, pl.T, "aux"))) return rc;
    }
    V.aux[2].value = pl.aux[2].kind == JBUGS_ARG_ONE ? 1.0 : pl.aux[2].value;
    if (pl.aux[2].kind == JBUGS_ARG_DATA_I || pl.aux[2].kind == JBUGS_ARG_DATA_IJ) {
        if (pl.family == JBUGS_FAM_T || pl.family == JBUGS_FAM_BETABIN) return fail(JBUGS_ERR_UNSUPPORTED, "plate %d: dt / dbetabin observations cannot carry a weight", k);
        if ((rc = resolve_arg(p, pl.aux[2], sh.aux[2], V.aux[2], gref, (int)pl.n_locals, sb))) return rc;
        if ((rc = check_len(p, pl.aux[2], pl.N, pl.T, "weight"))) return rc;
    }
    cudaFree(r.d_xtape);
    r.d_xtape = nullptr;
    CU(cudaMalloc(&r.d_xtape, tape.size() * sizeof(XOpDev)));
    CU(cudaMemcpy(r.d_xtape, tape.data(), tape.size() * sizeof(XOpDev), cudaMemcpyHostToDevice));
    V.xtape = r.d_xtape;
    r.xtape_host = tape;
    return build_locals_and_finish(p, k, gref, sb);
}

static int build_xplate(jbugs_plan* p, int k);
static int build_plate(jbugs_plan* p, int k) {
    const jbugs_plate& pl = p->plates[k];
    PlateRt& r = p->rt[k];
    PlateParams& pp = r.pp;
    memset(&pp, 0, sizeof pp);
    if (pl.N < 0 || pl.T < 1) return fail(JBUGS_ERR_INVALID, "plate %d: bad extents", k);
    if (pl.n_locals > (uint32_t)MAX_LOC) return fail(JBUGS_ERR_UNSUPPORTED, "plate %d: more than %d locals", k, MAX_LOC);
    if (pl.n_link > JBUGS_MAX_LINK) return fail(JBUGS_ERR_INVALID, "plate %d: link tape too long", k);
    if (pl.family >= JBUGS_FAM_COUNT) return fail(JBUGS_ERR_INVALID, "plate %d: unknown family", k);
    if (pl.has_obs > 3) return fail(JBUGS_ERR_INVALID, "plate %d: unknown plate kind", k);
    if (pl.has_obs >= 2) return build_xplate(p, k);
    if (pl.eta_arg > 1) return fail(JBUGS_ERR_UNSUPPORTED, "plate %d: eta_arg > 1", k);
    const jbugs_local* L = p->locals.data() + pl.first_local;
    const jbugs_term* T = p->terms.data() + pl.first_term;
    std::vector<int> gref;
    // canonical term order: GVAL-coef terms, then one term per local (in local order), then an offset
    int t = 0, kg = 0, nl = 0, has_off = 0;
    while (t < (int)pl.n_terms && T[t].coef.kind == JBUGS_ARG_GVAL) { ++t; ++kg; }
    while (t < (int)pl.n_terms && T[t].coef.kind == JBUGS_ARG_LOCAL && T[t].coef.id == nl) { ++t; ++nl; }
    if (t < (int)pl.n_terms && T[t].coef.kind == JBUGS_ARG_CONST && T[t].coef.value == 1.0) { ++t; has_off = 1; }
    if (t != (int)pl.n_terms || (pl.has_obs && nl != (int)pl.n_locals))
        return fail(JBUGS_ERR_UNSUPPORTED, "plate %d: terms are not in canonical order (global terms, one term per local, offset)", k);
    if (kg > MAX_GT) return fail(JBUGS_ERR_UNSUPPORTED, "plate %d: more than %d global terms", k, MAX_GT);
    PlateShape& sh = pp.sh;
    PlateVals& V = pp.v;
    sh.kg = kg; sh.nl = (int)pl.n_locals; sh.has_off = has_off;
    sh.family = (int)pl.family; sh.eta_arg = (int)pl.eta_arg; sh.has_obs = (int)pl.has_obs; sh.n_link = (int)pl.n_link;
    for (int q = 0; q < JBUGS_MAX_LINK; ++q) sh.link[q] = q < (int)pl.n_link ? (int)pl.link[q] : 0;
    V.N = pl.N; V.T = pl.T;
    StreamBuilder sb;
    sb.V = &V;
    sb.T = pl.T;
    int rc;
    for (int q = 0; q < kg; ++q) {
        ArgS s; ArgV v;
        if ((rc = resolve_arg(p, T[q].coef, s, v, gref, 0, sb))) return rc;
        if (s.id != q) return fail(JBUGS_ERR_UNSUPPORTED, "plate %d: a global appears in two terms", k);
    }
    for (int q = 0; q < (int)pl.n_terms; ++q) {
        if ((rc = resolve_arg(p, T[q].cov, sh.cov[q], V.cov[q], gref, 0, sb))) return rc;
        if (sh.cov[q].kind == JBUGS_ARG_GVAL || sh.cov[q].kind == JBUGS_ARG_LOCAL)
            return fail(JBUGS_ERR_UNSUPPORTED, "plate %d: covariates must be data", k);
        if ((rc = check_len(p, T[q].cov, pl.N, pl.T, "covariate"))) return rc;
    }
    if (pl.has_obs) {
        if ((rc = resolve_arg(p, pl.y, sh.y, V.y, gref, 0, sb))) return rc;
        if ((rc = check_len(p, pl.y, pl.N, pl.T, "y"))) return rc;
        for (int q = 0; q < 2; ++q) {
            if (q == (int)pl.eta_arg) continue;
            if ((rc = resolve_arg(p, pl.aux[q], sh.aux[q], V.aux[q], gref, (int)pl.n_locals, sb))) return rc;
            if ((rc = check_len(p, pl.aux[q], pl.N, pl.T, "aux"))) return rc;
        }
        // aux[2]: the constant third argument of a three-argument family (dt), or the optional 0/1 observation weight of
        // a gathered (padded) plate
        V.aux[2].value = pl.aux[2].kind == JBUGS_ARG_ONE ? 1.0 : pl.aux[2].value;
        if ((pl.family == JBUGS_FAM_T || pl.family == JBUGS_FAM_BETABIN) && pl.aux[2].kind != JBUGS_ARG_CONST && pl.aux[2].kind != JBUGS_ARG_ONE)
            return fail(JBUGS_ERR_UNSUPPORTED, "plate %d: the third argument of dt / dbetabin must be a constant", k);
        if (pl.aux[2].kind == JBUGS_ARG_DATA_I || pl.aux[2].kind == JBUGS_ARG_DATA_IJ) {
            if ((rc = resolve_arg(p, pl.aux[2], sh.aux[2], V.aux[2], gref, (int)pl.n_locals, sb))) return rc;
            if ((rc = check_len(p, pl.aux[2], pl.N, pl.T, "weight"))) return rc;
        }
    }
    if ((rc = build_locals_and_finish(p, k, gref, sb))) return rc;
    return 0;
}

// data-only normalising constants of a fused plate (and validation of the data the fused formulas assume)
static int refresh_obs_const(jbugs_plan* p, int k, cudaStream_t st) {
    PlateRt& r = p->rt[k];
    if (!r.const_dirty) return 0;
    r.obs_const = 0.0;
    r.const_dirty = false;
    p->epoch++;
    const int fused = (r.variant == 0 && r.jit && !p->force_dynamic) ? r.jit_fused : g_specs[r.variant].fused;
    r.jit_valid = true;
    if (fused == FUSED_NONE || fused == FUSED_NORMAL_IDENTITY) return 0;  // the other paths drop a data-only normaliser
    const PlateParams& pp = r.pp;
    if (pp.v.i1 <= pp.v.i0) return 0;
    double* d_out = p->d_gvals ? p->d_red : nullptr;  // d_red has room for >= 2 doubles once a workspace exists
    if (!d_out) return fail(JBUGS_ERR_STATE, "internal: workspace missing");
    obs_const_kernel<<<1, 1024, 0, st>>>(fused, pp.v.y.ptr, pp.v.aux[1].ptr, pp.sh.y.kind == JBUGS_ARG_DATA_IJ,
                                         pp.sh.aux[1].kind, pp.v.aux[1].value, pp.v.N, pp.v.T, pp.v.i0, pp.v.i1, d_out);
    p->launches++;
    double h[2] = {0.0, 0.0};
    CU(cudaMemcpyAsync(h, d_out, sizeof h, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (h[1] != 0.0) {
        r.variant = 0;  // data outside the fused formulas' domain: the interpreter reproduces the -Inf semantics
        r.jit_valid = false;
        r.obs_const = 0.0;
    } else {
        r.obs_const = h[0];
    }
    return 0;
}

// kernel variant of a plate: expression plates run the tape interpreter; everything else a static spec whose shape is
// identical, else the generic interpreter (variant 0; also when the caller forces it)
static int pick_variant(const jbugs_plan* p, const PlateRt& r, long long T) {
    if (r.pp.sh.n_x > 0) return r.pp.sh.n_gref > MAX_GREF_LINEAR ? g_xwspec : g_xspec;
    return p->force_dynamic ? 0 : select_spec(r.pp.sh, T);
}

static int rebuild_plates(jbugs_plan* p) {
    for (int k = 0; k < (int)p->h.n_plates; ++k) {
        int rc = build_plate(p, k);
        if (rc) return rc;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ C ABI: basics
extern "C" const char* jbugs_last_error(void) { return g_err.c_str(); }
extern "C" int jbugs_version(void) { return 100; }

extern "C" int jbugs_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int jbugs_plan_create(const void* blob, size_t nbytes, int device, jbugs_plan** out) {
    if (!blob || !out) return fail(JBUGS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (nbytes < sizeof(jbugs_plan_header)) return fail(JBUGS_ERR_INVALID, "plan blob shorter than its header");
    jbugs_plan_header h;
    memcpy(&h, blob, sizeof h);
    if (memcmp(h.magic, JBUGS_PLAN_MAGIC, 8) != 0) return fail(JBUGS_ERR_INVALID, "bad plan magic");
    if (h.version != JBUGS_PLAN_VERSION) return fail(JBUGS_ERR_INVALID, "plan version %u, library speaks %u", h.version, JBUGS_PLAN_VERSION);
    const size_t need = sizeof h + h.n_globals * sizeof(jbugs_global) + h.n_gtape * sizeof(jbugs_gop) +
                        h.n_data * sizeof(jbugs_data_desc) + h.n_plates * sizeof(jbugs_plate) +
                        h.n_locals * sizeof(jbugs_local) + h.n_terms * sizeof(jbugs_term) +
                        h.n_dense * sizeof(jbugs_dense) + h.n_xops * sizeof(jbugs_xop);
    if (need != nbytes) return fail(JBUGS_ERR_INVALID, "plan blob is %zu bytes, records need %zu", nbytes, need);
    if (h.D < 0) return fail(JBUGS_ERR_INVALID, "negative dimension");
    if (h.n_globals + h.n_gtape > (uint32_t)MAX_GVALS) return fail(JBUGS_ERR_UNSUPPORTED, "more than %d global values", MAX_GVALS);
    if (h.n_plates + h.n_dense > 8) return fail(JBUGS_ERR_UNSUPPORTED, "more than 8 plates");
    if (h.n_dense > 1) return fail(JBUGS_ERR_UNSUPPORTED, "more than one dense predictor");
    int ndev = jbugs_device_count();
    if (ndev <= 0) return fail(JBUGS_ERR_CUDA, "no CUDA device is available (libjbugs_cuda has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(JBUGS_ERR_INVALID, "device %d out of range (0..%d)", device, ndev - 1);

    jbugs_plan* p = new jbugs_plan_s();
    p->device = device;
    p->h = h;
    const char* c = (const char*)blob + sizeof h;
    auto take = [&](auto& vec, size_t n) {
        vec.resize(n);
        if (n) memcpy(vec.data(), c, n * sizeof(vec[0]));
        c += n * sizeof(vec[0]);
    };
    take(p->globals, h.n_globals);
    take(p->gtape, h.n_gtape);
    take(p->data, h.n_data);
    take(p->plates, h.n_plates);
    take(p->locals, h.n_locals);
    take(p->terms, h.n_terms);
    take(p->dense, h.n_dense);
    take(p->xops, h.n_xops);
    auto destroy = [&](int rc) {
        jbugs_plan_destroy(p);
        return rc;
    };
    // validation of record ranges
    for (auto& pl : p->plates) {
        const size_t tape_len = pl.has_obs >= 2 ? p->xops.size() : p->terms.size();
        if ((size_t)pl.first_local + pl.n_locals > p->locals.size() || (size_t)pl.first_term + pl.n_terms > tape_len)
            return destroy(fail(JBUGS_ERR_INVALID, "plate record ranges exceed the blob"));
    }
    for (size_t k = 0; k < p->globals.size(); ++k) {
        const auto& g = p->globals[k];
        if (g.theta_idx < 0 || g.theta_idx >= h.D) return destroy(fail(JBUGS_ERR_INVALID, "global %zu: theta slot out of range", k));
        if (g.family >= JBUGS_FAM_COUNT) return destroy(fail(JBUGS_ERR_INVALID, "global %zu: unknown family", k));
        for (int q = 0; q < 2; ++q)
            if (g.args[q].kind == JBUGS_ARG_GVAL && (g.args[q].id < 0 || g.args[q].id >= (int)(h.n_globals + h.n_gtape)))
                return destroy(fail(JBUGS_ERR_INVALID, "global %zu: argument out of range", k));
            else if (g.args[q].kind != JBUGS_ARG_GVAL && g.args[q].kind != JBUGS_ARG_CONST && g.args[q].kind != JBUGS_ARG_ONE)
                return destroy(fail(JBUGS_ERR_INVALID, "global %zu: prior arguments must be constants or global values", k));
    }
    for (size_t k = 0; k < p->gtape.size(); ++k) {
        const auto& o = p->gtape[k];
        const int lim = (int)(h.n_globals + k);  // may only look backwards
        if ((o.a.kind == JBUGS_ARG_GVAL && (o.a.id < 0 || o.a.id >= lim)) ||
            (o.op >= 16 && o.b.kind == JBUGS_ARG_GVAL && (o.b.id < 0 || o.b.id >= lim)))
            return destroy(fail(JBUGS_ERR_INVALID, "gtape op %zu references a later value", k));
    }
    if (cudaSetDevice(device) != cudaSuccess) return destroy(fail(JBUGS_ERR_CUDA, "cudaSetDevice(%d) failed", device));
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return destroy(fail(JBUGS_ERR_CUDA, "cudaGetDeviceProperties failed"));
    p->num_sms = prop.multiProcessorCount;
    p->d_slots.assign(h.n_data, nullptr);
    p->uploaded.assign(h.n_data, 0);
    for (uint32_t k = 0; k < h.n_data; ++k) {
        if (p->data[k].len < 0) return destroy(fail(JBUGS_ERR_INVALID, "data slot %u: negative length", k));
        const size_t bytes = (size_t)std::max<long long>(p->data[k].len, 1) * 8;
        if (cudaMalloc(&p->d_slots[k], bytes) != cudaSuccess) return destroy(fail(JBUGS_ERR_CUDA, "cudaMalloc of data slot %u (%zu bytes) failed", k, bytes));
    }
    p->rt.resize(h.n_plates);
    p->shard_i0.assign(h.n_plates, 0);
    p->shard_set.assign(h.n_plates, 0);
    p->shard_i1.resize(h.n_plates);
    for (uint32_t k = 0; k < h.n_plates; ++k) p->shard_i1[k] = p->plates[k].N;
    p->gp.n_globals = (int)h.n_globals;
    p->gp.n_gtape = (int)h.n_gtape;
    p->gp.transformed = (h.flags & JBUGS_FLAG_TRANSFORMED) != 0;
    for (uint32_t k = 0; k < h.n_globals; ++k) {
        const jbugs_global& g = p->globals[k];
        p->gp.g[k] = g;
        p->gp.fast[k] = 0;
        p->gp.cst[k] = 0.0;
        auto lit = [](const jbugs_arg& a) { return a.kind == JBUGS_ARG_CONST || a.kind == JBUGS_ARG_ONE; };
        auto val = [](const jbugs_arg& a) { return a.kind == JBUGS_ARG_ONE ? 1.0 : a.value; };
        if (lit(g.args[0]) && lit(g.args[1])) {
            const double a0 = val(g.args[0]), a1 = val(g.args[1]);
            p->gp.g[k].args[0].value = a0;
            p->gp.g[k].args[1].value = a1;
            if (g.family == JBUGS_FAM_NORMAL && a1 > 0.0) {
                p->gp.fast[k] = JBUGS_FAM_NORMAL + 1;
                p->gp.cst[k] = 0.5 * (std::log(a1) - 1.8378770664093454835606594728112);
            } else if (g.family == JBUGS_FAM_GAMMA && a0 > 0.0 && a1 > 0.0) {
                p->gp.fast[k] = JBUGS_FAM_GAMMA + 1;
                p->gp.cst[k] = a0 * std::log(a1) - std::lgamma(a0);
            }
        }
        // truncated(dist(args), lb, ub) (BUGS: dist(args) T(lb, ub)): a global whose support bounds are narrower than the
        // natural support of its family (include/jbugs_trunc.h).  The bounds already steer the bijector; the normaliser
        // log(cdf(ub) - cdf(lb)) is a constant for literal arguments and joins the host-folded part of the prior (a prior
        // with parameter-valued arguments and narrowed bounds is refused by the lowering).
        if (lit(g.args[0]) && lit(g.args[1]) && jbugs_is_truncated(g.family, val(g.args[0]), val(g.args[1]), g.lb, g.ub) == 1) {
            if (!p->gp.transformed)
                return destroy(fail(JBUGS_ERR_UNSUPPORTED, "global %u: truncated priors are evaluated in transformed space only "
                                                           "(the support check of the untransformed value is not on the device)", k));
            if (g.family == JBUGS_FAM_T && !lit(g.args[2]))
                return destroy(fail(JBUGS_ERR_UNSUPPORTED, "global %u: a truncated dt prior needs literal degrees of freedom", k));
            const double a2 = lit(g.args[2]) ? val(g.args[2]) : 0.0;   // (dt's degrees of freedom; unused by the two-argument families)
            const double lm = jbugs_trunc_logmass(g.family, val(g.args[0]), val(g.args[1]), a2, g.lb, g.ub);
            if (!std::isfinite(lm)) return destroy(fail(JBUGS_ERR_INVALID, "global %u: the truncation bounds hold no mass", k));
            p->gp.cst[k] -= lm;
        }
    }
    for (uint32_t k = 0; k < h.n_gtape; ++k) p->tp.op[k] = p->gtape[k];
    int rc = rebuild_plates(p);
    if (rc) return destroy(rc);
    for (size_t k = 0; k < p->dense.size(); ++k) {
        const jbugs_dense& dn = p->dense[k];
        auto slot_ok = [&](int s, long long need) { return s >= 0 && s < (int)h.n_data && p->data[s].dtype == 0 && p->data[s].len >= need; };
        if (dn.N < 0 || dn.P < 1) return destroy(fail(JBUGS_ERR_INVALID, "dense %zu: bad extents", k));
        if (!slot_ok(dn.x_slot, dn.N * dn.P) || !slot_ok(dn.y_slot, dn.N) || (dn.n_slot >= 0 && !slot_ok(dn.n_slot, dn.N)))
            return destroy(fail(JBUGS_ERR_INVALID, "dense %zu: data slot missing or too short", k));
        const long long last = dn.theta_base + dn.theta_stride * (dn.P - 1);
        if (dn.theta_base < 0 || dn.theta_base >= h.D || last < 0 || last >= h.D)
            return destroy(fail(JBUGS_ERR_INVALID, "dense %zu: coefficient slots out of range", k));
        if (dn.link != JBUGS_OP_LOGISTIC || (dn.family != JBUGS_FAM_BERNOULLI && dn.family != JBUGS_FAM_BINOMIAL) ||
            (dn.family == JBUGS_FAM_BINOMIAL) != (dn.n_slot >= 0))
            return destroy(fail(JBUGS_ERR_UNSUPPORTED, "dense %zu: only dbern / dbin with a logit link", k));
    }
    for (auto& e : p->ev)
        if (cudaEventCreate(&e) != cudaSuccess) return destroy(fail(JBUGS_ERR_CUDA, "cudaEventCreate failed"));
    if (cudaMalloc(&p->d_any_bad, sizeof(int)) != cudaSuccess) return destroy(fail(JBUGS_ERR_CUDA, "cudaMalloc failed"));
    *out = p;
    return JBUGS_OK;
}

static void free_workspace(jbugs_plan* p) {
    cudaFree(p->d_gvals); p->d_gvals = nullptr; p->gvals_bytes = 0;
    cudaFree(p->d_partial); p->d_partial = nullptr; p->partial_bytes = 0;
    cudaFree(p->d_partial2); p->d_partial2 = nullptr; p->partial2_bytes = 0;
    cudaFree(p->d_red); p->d_red = nullptr; p->red_bytes = 0;
    cudaFree(p->d_flags); p->d_flags = nullptr; p->flags_bytes = 0;
    cudaFree(p->d_R); p->d_R = nullptr; p->R_bytes = 0;
    cudaFree(p->d_Gp); p->d_Gp = nullptr; p->Gp_bytes = 0;
    p->cap_C = 0;
}

extern "C" int jbugs_batch_free(jbugs_plan* p) {
    if (!p) return fail(JBUGS_ERR_INVALID, "null plan");
    cudaSetDevice(p->device);
    cudaFree(p->b_theta); cudaFree(p->b_grad); cudaFree(p->b_logp); cudaFree(p->b_status);
    p->b_theta = p->b_grad = p->b_logp = nullptr;
    p->b_status = nullptr;
    return JBUGS_OK;
}

extern "C" int jbugs_comm_destroy(jbugs_plan* p) {
    if (!p) return fail(JBUGS_ERR_INVALID, "null plan");
    if (p->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(p->comm);
    p->comm = nullptr;
    p->nranks = 1;
    p->rank = 0;
    return JBUGS_OK;
}

extern "C" int jbugs_plan_destroy(jbugs_plan* p) {
    if (!p) return JBUGS_OK;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();  // asynchronous evaluations may still be running on the plan's buffers and kernels
    jbugs_comm_destroy(p);
    for (void* d : p->d_slots) cudaFree(d);
    jit_job_free(p->jit_job);  // (waits for the worker thread)
    for (auto* jk : p->jit_retired) jit_free(jk);
    for (auto& r : p->rt) {
        cudaFree(r.d_xtape);
        jit_free(r.jit);
    }
    free_workspace(p);
    free_slots(p);
    for (auto& s : p->slot) {
        if (s.in_done) cudaEventDestroy(s.in_done);
        if (s.comp_done) cudaEventDestroy(s.comp_done);
        if (s.out_done) cudaEventDestroy(s.out_done);
    }
    if (p->st_in) cudaStreamDestroy(p->st_in);
    if (p->st_out) cudaStreamDestroy(p->st_out);
    jbugs_batch_free(p);
    cudaFree(p->d_any_bad);
    cudaFree(p->d_fflags);
    cudaFree(p->d_tickets);
    for (auto& e : p->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : p->ring)
        if (e) cudaEventDestroy(e);
    delete p;
    return JBUGS_OK;
}

static bool is_device_ptr_strict(const void* ptr) {
    cudaPointerAttributes a;
    if (!ptr || cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice;
}

extern "C" int64_t jbugs_dim(const jbugs_plan* p) { return p ? p->h.D : -1; }
extern "C" int64_t jbugs_launch_count(const jbugs_plan* p) { return p ? p->launches : -1; }

extern "C" int jbugs_plan_upload_data(jbugs_plan* p, int slot, const void* src, size_t nbytes, int dtype) {
    if (!p || !src) return fail(JBUGS_ERR_INVALID, "null argument");
    if (slot < 0 || slot >= (int)p->h.n_data) return fail(JBUGS_ERR_INVALID, "data slot %d out of range", slot);
    if ((int)p->data[slot].dtype != dtype) return fail(JBUGS_ERR_INVALID, "data slot %d: dtype %d expected", slot, (int)p->data[slot].dtype);
    if (nbytes != (size_t)p->data[slot].len * 8) return fail(JBUGS_ERR_INVALID, "data slot %d: %zu bytes given, %zu expected", slot, nbytes, (size_t)p->data[slot].len * 8);
    int rc = set_device(p);
    if (rc) return rc;
    if (dtype == 1) {
        // an int64 slot used as the explicit theta-slot map of a local (gathered plates): the kernels use its values
        // as theta / gradient row numbers, so every one must lie in [0, D)
        bool is_index = false;
        for (const jbugs_local& L : p->locals) is_index |= L.idx_slot == slot;
        if (is_index) {
            const size_t n = nbytes / 8;
            std::vector<long long> tmp;
            const long long* v = (const long long*)src;
            if (is_device_ptr_strict(src)) {
                tmp.resize(n);
                CU(cudaMemcpy(tmp.data(), src, nbytes, cudaMemcpyDeviceToHost));
                v = tmp.data();
            }
            for (size_t q = 0; q < n; ++q)
                if (v[q] < 0 || v[q] >= p->h.D)
                    return fail(JBUGS_ERR_INVALID, "data slot %d: index %lld at element %zu is outside [0, %lld)", slot, v[q], q, (long long)p->h.D);
        }
    }
    CU(cudaMemcpy(p->d_slots[slot], src, nbytes, cudaMemcpyDefault));
    p->epoch++;  // data-only constants baked into captured launch sequences are stale now
    p->uploaded[slot] = 1;
    // new observations: constants of the fused paths (and their validity check) must be redone
    p->dense_const_dirty = true;
    if (!p->force_dynamic)
        for (size_t k = 0; k < p->rt.size(); ++k) {
            p->rt[k].variant = pick_variant(p, p->rt[k], p->rt[k].pp.v.T);
            p->rt[k].const_dirty = true;
        }
    return JBUGS_OK;
}

extern "C" int jbugs_plan_set_shard(jbugs_plan* p, int plate, int64_t i0, int64_t i1) {
    if (!p) return fail(JBUGS_ERR_INVALID, "null plan");
    if (plate < 0 || plate >= (int)p->h.n_plates) return fail(JBUGS_ERR_INVALID, "plate %d out of range", plate);
    if (i0 < 0 || i1 > p->plates[plate].N || i0 > i1) return fail(JBUGS_ERR_INVALID, "shard [%lld, %lld) outside [0, %lld)", (long long)i0, (long long)i1, (long long)p->plates[plate].N);
    p->shard_i0[plate] = i0;
    p->shard_set[plate] = 1;
    p->shard_i1[plate] = i1;
    p->rt[plate].pp.v.i0 = i0;
    p->rt[plate].pp.v.i1 = i1;
    p->rt[plate].const_dirty = true;
    p->cap_C = 0;  // force re-tiling
    return JBUGS_OK;
}

extern "C" int jbugs_plan_set_dense_shard(jbugs_plan* p, int64_t i0, int64_t i1) {
    if (!p) return fail(JBUGS_ERR_INVALID, "null plan");
    if (p->dense.empty()) return fail(JBUGS_ERR_INVALID, "the plan has no dense predictor");
    const long long N = p->dense[0].N;
    if (i0 < 0 || i1 > N || i0 > i1) return fail(JBUGS_ERR_INVALID, "shard [%lld, %lld) outside [0, %lld)", (long long)i0, (long long)i1, N);
    p->dense_i0 = i0;
    p->dense_i1 = i1;
    p->dense_const_dirty = true;
    p->cap_C = 0;  // row groups / split-K factor depend on the local row count
    return JBUGS_OK;
}

extern "C" int jbugs_plan_set_option(jbugs_plan* p, const char* key, int64_t value) {
    if (!p || !key) return fail(JBUGS_ERR_INVALID, "null argument");
    if (!strcmp(key, "dynamic")) {
        p->force_dynamic = value != 0;
        return rebuild_plates(p);
    }
    if (!strcmp(key, "tile")) {
        if (value < 0) return fail(JBUGS_ERR_INVALID, "tile must be >= 0");
        p->tile_override = value;
        p->cap_C = 0;
        return JBUGS_OK;
    }
    if (!strcmp(key, "tma")) {
        p->use_tma = value != 0;
        return rebuild_plates(p);
    }
    if (!strcmp(key, "host_chunk")) {
        if (value < 0) return fail(JBUGS_ERR_INVALID, "host_chunk must be >= 0");
        p->host_chunk = value;
        return JBUGS_OK;
    }
    if (!strcmp(key, "one_launch")) {  // 0: always the three-kernel sequence (A/B against fused_small_kernel)
        p->one_launch = value != 0;
        p->epoch++;
        return JBUGS_OK;
    }
    if (!strcmp(key, "pdl")) {
        p->use_pdl = value != 0;
        p->epoch++;
        return JBUGS_OK;
    }
    if (!strcmp(key, "timing")) {
        p->timing = value != 0;
        p->ring_n = 0;
        if (value == 2 && p->ring.empty()) {  // one set of four events per evaluation, RING evaluations deep
            int rc = set_device(p);
            if (rc) return rc;
            p->ring.assign(4 * jbugs_plan_s::RING, nullptr);
            for (auto& e : p->ring) CU(cudaEventCreate(&e));
        }
        if (value != 2) {
            for (auto& e : p->ring)
                if (e) cudaEventDestroy(e);
            p->ring.clear();
        }
        return JBUGS_OK;
    }
    return fail(JBUGS_ERR_INVALID, "unknown option '%s'", key);
}

extern "C" int jbugs_plan_describe(const jbugs_plan* p, char* buf, size_t nbuf) {
    if (!p || !buf || !nbuf) return fail(JBUGS_ERR_INVALID, "null argument");
    std::string s;
    char line[256];
    snprintf(line, sizeof line, "D=%lld globals=%u gtape=%u plates=%u sms=%d\n", (long long)p->h.D, p->h.n_globals, p->h.n_gtape, p->h.n_plates, p->num_sms);
    s += line;
    for (size_t k = 0; k < p->rt.size(); ++k) {
        const auto& r = p->rt[k];
        snprintf(line, sizeof line, "plate %zu: N=%lld T=%lld kernel=%s tiles=%lld tile=%lld gref=%d\n", k, r.pp.v.N, r.pp.v.T,
                 ((r.variant == 0 || is_tape_variant(r.variant)) && r.jit && r.jit_valid && !p->force_dynamic)
                     ? (r.variant == 0 ? (jit_is_best(r.jit) ? "specialised_at_run_time" : "specialised_at_run_time_quick_variant")
                                       : "expression_tape_specialised_at_run_time") : spec_name(r.variant),
                 r.n_tiles, r.pp.v.tile, r.pp.sh.n_gref);
        s += line;
    }
    snprintf(buf, nbuf, "%s", s.c_str());
    return JBUGS_OK;
}

extern "C" int jbugs_last_timing(const jbugs_plan* p, float* plate_ms, float* total_ms) {
    if (!p) return fail(JBUGS_ERR_INVALID, "null plan");
    if (p->timing && p->ev_last[3] && p->events_recorded) {
        // events of the most recent evaluation (also an asynchronous one): wait for its last event, then read them
        jbugs_plan* q = const_cast<jbugs_plan*>(p);
        CU(cudaSetDevice(p->device));
        CU(cudaEventSynchronize(p->ev_last[3]));
        CU(cudaEventElapsedTime(&q->last_plate_ms, p->ev_last[1], p->ev_last[2]));
        CU(cudaEventElapsedTime(&q->last_total_ms, p->ev_last[0], p->ev_last[3]));
    }
    if (plate_ms) *plate_ms = p->last_plate_ms;
    if (total_ms) *total_ms = p->last_total_ms;
    return JBUGS_OK;
}

extern "C" int jbugs_timing_history(const jbugs_plan* p, float* plate_ms, float* total_ms, int cap, int* n_out) {
    if (!p || !n_out) return fail(JBUGS_ERR_INVALID, "null argument");
    *n_out = 0;
    if (p->ring.empty() || p->ring_n == 0) return JBUGS_OK;
    CU(cudaSetDevice(p->device));
    const long long n = std::min<long long>(p->ring_n, jbugs_plan_s::RING), first = p->ring_n - n;
    int k = 0;
    for (long long q = first; q < p->ring_n && k < cap; ++q, ++k) {
        const cudaEvent_t* e = p->ring.data() + 4 * (q % jbugs_plan_s::RING);
        CU(cudaEventSynchronize(e[3]));
        float a = 0.f, b = 0.f;
        CU(cudaEventElapsedTime(&a, e[1], e[2]));
        CU(cudaEventElapsedTime(&b, e[0], e[3]));
        if (plate_ms) plate_ms[k] = a;
        if (total_ms) total_ms[k] = b;
    }
    *n_out = k;
    return JBUGS_OK;
}

// FP64 pipe peak measured on this device: independent DFMA chains, 8 per thread, enough warps to cover the pipe latency.
// The logit / log-link plate kernels are bound by this pipe, not by HBM; bench.py reports them
// against the number this returns (fused multiply-adds per second, all SMs).
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678) out[0] = s;  // never true: keeps the chains alive
}

extern "C" int jbugs_fp64_peak(int device, double* fma_per_s) {
    if (!fma_per_s) return fail(JBUGS_ERR_INVALID, "null argument");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    double* d_out = nullptr;
    CU(cudaMalloc(&d_out, 8));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const int ctas = prop.multiProcessorCount * 8, iters = 1 << 14;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(e0, 0));
        dfma_peak_kernel<<<ctas, 256>>>(d_out, iters, 0.999999, 1e-9);
        CU(cudaEventRecord(e1, 0));
        CU(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double rate = (double)ctas * 256.0 * 8.0 * iters / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;  // first launch: warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    *fma_per_s = best;
    return JBUGS_OK;
}

extern "C" int jbugs_host_alloc(void** ptr, size_t nbytes) {
    if (!ptr) return fail(JBUGS_ERR_INVALID, "null argument");
    CU(cudaHostAlloc(ptr, nbytes, cudaHostAllocDefault));
    return JBUGS_OK;
}
extern "C" int jbugs_host_alloc_ex(void** ptr, size_t nbytes, int flags) {
    if (!ptr) return fail(JBUGS_ERR_INVALID, "null argument");
    CU(cudaHostAlloc(ptr, nbytes, (flags & JBUGS_HOST_WRITE_COMBINED) ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
    return JBUGS_OK;
}

// Bind the calling process (threads created later inherit it) to the CPUs and the memory of the NUMA node the GPU hangs
// off, so that pinned staging buffers allocated afterwards are node-local: eight ranks that all stage through node 0
// share one memory controller and one PCIe root complex (round 1: 0.19 scaling of the host-buffer path at 8 GPUs).
// Topology comes from sysfs (/sys/bus/pci/devices/<bdf>/numa_node, /sys/devices/system/node/node<k>/cpulist).
// Returns the node (>= 0) in *node_out, -1 when the platform reports none; never fails the caller for that.
extern "C" int jbugs_bind_host_to_device(int device, int* node_out, int* n_cpus_out) {
    if (node_out) *node_out = -1;
    if (n_cpus_out) *n_cpus_out = 0;
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, device) != cudaSuccess) {
        cudaGetLastError();
        return fail(JBUGS_ERR_CUDA, "cudaDeviceGetPCIBusId(%d) failed", device);
    }
    for (char* c = bdf; *c; ++c) *c = (char)tolower(*c);
    char path[256];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bdf);
    FILE* f = fopen(path, "r");
    int node = -1;
    if (f) {
        if (fscanf(f, "%d", &node) != 1) node = -1;
        fclose(f);
    }
    if (node < 0) return JBUGS_OK;  // single-node box or a VM that hides the topology
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return JBUGS_OK;
    char list[4096] = {0};
    if (!fgets(list, sizeof list, f)) list[0] = 0;
    fclose(f);
    cpu_set_t cur, want;
    CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof cur, &cur) != 0) return JBUGS_OK;
    int n = 0;
    for (char* tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int got = sscanf(tok, "%d-%d", &a, &b);
        if (got < 1) continue;
        if (got == 1) b = a;
        for (int c = a; c <= b && c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &cur)) { CPU_SET(c, &want); ++n; }
    }
    if (n > 0) sched_setaffinity(0, sizeof want, &want);  // (none of the node's CPUs allowed: keep the current set)
    // memory: prefer the node (MPOL_PREFERRED = 1): falls back silently when the cpuset forbids it
    unsigned long mask[16] = {0};
    if (node < (int)(sizeof mask * 8)) {
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, sizeof mask * 8);
    }
    if (node_out) *node_out = node;
    if (n_cpus_out) *n_cpus_out = n;
    return JBUGS_OK;
}
extern "C" int jbugs_host_free(void* ptr) {
    CU(cudaFreeHost(ptr));
    return JBUGS_OK;
}

// ------------------------------------------------------------------------------------------------ workspaces
static long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

static int grow(void** ptr, size_t* cap, size_t need) {
    if (need <= *cap && *ptr) return 0;
    cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    CU(cudaMalloc(ptr, need));
    *cap = need;
    return 0;
}

// (re)tile the plates for a batch of C chains and make sure the workspaces are large enough (grow only)
static int ensure_workspace(jbugs_plan* p, long long C) {
    if (C == p->cap_C) return 0;
    p->epoch++;
    const long long ldg = round_up(C, 16);
    const int NG = (int)(p->h.n_globals + p->h.n_gtape);
    const int blk = plate_block(C);
    const long long chain_tiles = (C + blk - 1) / blk;
    // tiling: enough CTAs for ~4 full waves of 2048 resident threads per SM; tiles are whole staged chunks
    const long long target = (long long)p->num_sms * (2048 / blk) * 4;
    long long rows = 0, rows2 = 0;
    for (size_t k = 0; k < p->rt.size(); ++k) {
        PlateRt& r = p->rt[k];
        const long long n = std::max<long long>(r.pp.v.i1 - r.pp.v.i0, 0);
        long long nt = std::max<long long>(1, std::min<long long>((target + chain_tiles - 1) / chain_tiles, (n + 3) / 4));
        // small plates (at most ~two waves of CTAs even at 4 groups per tile) are latency-bound: one full wave of
        // ~4 resident CTAs per SM beats a ragged second wave (Epil, 8192 chains: 58 -> 55 us per evaluation)
        const long long slots = 4LL * p->num_sms;
        if (chain_tiles * nt <= 2 * slots) nt = std::max<long long>(1, std::min<long long>(nt, slots / chain_tiles));
        long long tile = std::max<long long>(1, (n + nt - 1) / nt);
        if (tile > STAGE_G) tile = round_up(tile, STAGE_G);
        if (p->tile_override > 0) tile = p->tile_override;
        nt = std::max<long long>(1, (n + tile - 1) / tile);
        if (nt > 65535) { tile = round_up((n + 65534) / 65535, STAGE_G); nt = (n + tile - 1) / tile; }
        r.pp.v.tile = tile;
        r.n_tiles = nt;
        r.partial_offset = rows;
        r.partial2_offset = rows2;
        rows += nt * (1 + r.pp.sh.n_gref);
        rows2 += (long long)REDUCE_R * (1 + r.pp.sh.n_gref);
    }
    int rc;
    if (!p->dense.empty()) {
        // dense predictor: one logp partial row per forward-GEMM row group; R and the split-K partials of X^T R
        const jbugs_dense& dn = p->dense[0];
        const long long Nloc = (p->dense_i1 >= 0 ? p->dense_i1 : dn.N) - p->dense_i0;  // rows of X on this rank
        const long long col_blocks = (C + DG_BN - 1) / DG_BN;
        const long long m_tiles = std::max<long long>(1, (Nloc + DG_FWD_BM - 1) / DG_FWD_BM);
        p->dense_row_groups = (int)std::max<long long>(1, std::min<long long>(m_tiles, (2LL * p->num_sms + col_blocks - 1) / col_blocks));
        const long long out_tiles = col_blocks * ((dn.P + DG_BM - 1) / DG_BM);
        // split-K factor of the reverse GEMM: fill whole waves of one CTA per SM (wave quantisation costs up to 2x)
        {
            int best = 1;
            double best_eff = 0.0;
            for (int s = 1; s <= 96; ++s) {
                if (s > 1 && Nloc / s < 8 * DgBwd::BK) break;
                const long long ctas = out_tiles * s, waves = (ctas + p->num_sms - 1) / p->num_sms;
                const double eff = (double)ctas / (double)(waves * p->num_sms);
                if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
            }
            p->dense_splits = best;
        }
        p->dense_partial_offset = rows;
        rows += p->dense_row_groups;
        if ((rc = grow((void**)&p->d_R, &p->R_bytes, (size_t)std::max<long long>(Nloc, 1) * ldg * 8))) return rc;
        if ((rc = grow((void**)&p->d_Gp, &p->Gp_bytes, (size_t)p->dense_splits * dn.P * ldg * 8))) return rc;
    }
    if ((rc = grow((void**)&p->d_gvals, &p->gvals_bytes, (size_t)std::max(NG, 1) * ldg * 8))) return rc;
    if ((rc = grow((void**)&p->d_partial, &p->partial_bytes, (size_t)std::max<long long>(rows, 1) * ldg * 8))) return rc;
    if ((rc = grow((void**)&p->d_partial2, &p->partial2_bytes, (size_t)std::max<long long>(rows2, 1) * ldg * 8))) return rc;
    // all-reduce buffer of the observation-sharded path: [logp | NG global adjoints | flag | P rows of X^T R]
    const long long red_rows = 2 + std::max(NG, 1) + (p->dense.empty() ? 0 : p->dense[0].P);
    if ((rc = grow((void**)&p->d_red, &p->red_bytes, (size_t)red_rows * ldg * 8))) return rc;
    if ((rc = grow((void**)&p->d_flags, &p->flags_bytes, (size_t)ldg * sizeof(int)))) return rc;
    if ((size_t)ldg * sizeof(int) > p->fflags_bytes) {
        if ((rc = grow((void**)&p->d_fflags, &p->fflags_bytes, (size_t)ldg * sizeof(int)))) return rc;
        if ((rc = grow((void**)&p->d_tickets, &p->tickets_bytes, (size_t)(ldg / 32 + 1) * sizeof(unsigned)))) return rc;
        CU(cudaMemset(p->d_fflags, 0, p->fflags_bytes));
        CU(cudaMemset(p->d_tickets, 0, p->tickets_bytes));
    }
    p->cap_C = C;
    p->ldg = ldg;
    return 0;
}

// a thread-per-chain helper kernel, optionally as a programmatic dependent launch (the kernel starts with
// griddepcontrol.wait, so it is safe after any predecessor)
template <typename K, typename... Args>
static void launch_helper(K kernel, int pdl, unsigned grid, unsigned block, cudaStream_t st, Args&&... args) {
    if (!pdl) {
        kernel<<<grid, block, 0, st>>>(args...);
        return;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, args...);
}

// ------------------------------------------------------------------------------------------------ evaluation on device-resident CHAIN_FAST buffers
static int eval_device(jbugs_plan* p, const double* theta, long long ld, long long C, double* logp, double* grad,
                       int* status, int want_grad, cudaStream_t st) {
    if (C == 0) return 0;
    for (uint32_t k = 0; k < p->h.n_data; ++k)
        if (!p->uploaded[k]) return fail(JBUGS_ERR_STATE, "data slot %u has not been uploaded", k);
    int rc = ensure_workspace(p, C);
    if (rc) return rc;
    const long long ldg = p->ldg;
    const unsigned chain_tiles = (unsigned)((C + BLOCK - 1) / BLOCK);  // thread-per-chain helper kernels
    // prologue / finalize: one warp per CTA while that still leaves SMs idle (small ensembles are latency-bound)
    const unsigned hblk = C <= 64LL * p->num_sms ? 32 : BLOCK;
    const unsigned helper_tiles = (unsigned)((C + hblk - 1) / hblk);
    for (size_t k = 0; k < p->rt.size(); ++k)
        if ((rc = refresh_obs_const(p, (int)k, st))) return rc;
    // small single-plate model: prologue + plate + finalize in ONE kernel (latency-bound: BASELINE config 5)
    if (p->one_launch && p->rt.size() == 1 && p->dense.empty() && !(p->comm && p->nranks > 1) && !p->timing && p->parts == 0 &&
        p->h.D <= 4096 && p->rt[0].n_tiles <= 64 && p->rt[0].pp.v.i1 > p->rt[0].pp.v.i0) {
        PlateRt& r = p->rt[0];
        const bool jit = (r.variant == 0 || is_tape_variant(r.variant)) && r.jit && r.jit_valid && !p->force_dynamic;
        const fused_launch_fn fn = jit ? nullptr : g_specs[r.variant].one_launch;
        if (fn) {
            FusedParams fq;
            fq.G = p->gp;
            fq.TP = p->tp;
            memset(&fq.F, 0, sizeof fq.F);
            fq.F.n_plates = 1;
            fq.F.poison_rows = (want_grad && grad) ? p->h.D : 0;
            fq.F.p[0].n_tiles = (int)r.n_tiles;
            fq.F.p[0].n_gref = r.pp.sh.n_gref;
            fq.F.p[0].offset = 0;  // (the kernel gets the plate's own slice of the partial buffer)
            fq.F.p[0].obs_const = r.variant == 0 ? 0.0 : r.obs_const;  // the interpreter evaluates the full log-density itself
            for (int q = 0; q < MAX_GREF; ++q) fq.F.p[0].gref[q] = r.pp.v.gref[q];
            fq.P = r.pp;
            for (int l = 0; l < fq.P.sh.nl; ++l) {
                fq.P.v.loc[l].row0 = fq.P.v.loc[l].base * ld;
                fq.P.v.loc[l].step_i = fq.P.v.loc[l].si * ld;
                fq.P.v.loc[l].step_j = fq.P.v.loc[l].sj * ld;
            }
            cudaError_t le = fn(r.n_tiles, r.smem_bytes, st, fq, theta, grad, ld, C, p->d_gvals, ldg, p->d_partial + r.partial_offset * ldg,
                                p->d_fflags, logp, status, p->d_any_bad, want_grad && grad, p->d_tickets, p->use_pdl);
            if (le != cudaSuccess) return fail(JBUGS_ERR_CUDA, "one-launch kernel failed: %s", cudaGetErrorString(le));
            p->launches++;
            return 0;
        }
    }
    cudaEvent_t* tev = p->ev;  // the four events of this call
    if (p->timing && !p->ring.empty()) tev = p->ring.data() + 4 * (p->ring_n % jbugs_plan_s::RING);
    if (p->timing) CU(cudaEventRecord(tev[0], st));
    // (launched the ordinary way: as a dependent launch behind the sampler's leapfrog kernel it slowed the captured
    //  trajectory graph down, 1.4 s -> 2.2 s on the Rats probe, for 1.5 us per stand-alone evaluation)
    launch_helper(prologue_kernel, 0, helper_tiles, hblk, st, p->gp, p->tp, theta, ld, C, p->d_gvals, ldg, p->d_flags, p->d_any_bad);
    p->launches++;
    if (p->timing) CU(cudaEventRecord(tev[1], st));
    FinalParams fp{};
    std::vector<int> reduce_plates;
    fp.n_plates = (int)p->rt.size();
    // small models: the NaN gradient of a flagged chain is written by finalize itself (one launch fewer)
    fp.poison_rows = (want_grad && grad && p->h.D <= 4096) ? p->h.D : 0;
    for (size_t k = 0; k < p->rt.size(); ++k) {
        PlateRt& r = p->rt[k];
        fp.p[k].n_tiles = (int)r.n_tiles;
        fp.p[k].n_gref = r.pp.sh.n_gref;
        fp.p[k].offset = r.partial_offset;
        fp.p[k].obs_const = r.obs_const;
        fp.p[k].replicated = (p->comm && p->nranks > 1 && !p->shard_set[k]) ? 1 : 0;
        for (int q = 0; q < MAX_GREF; ++q) fp.p[k].gref[q] = r.pp.v.gref[q];
        if (r.pp.v.i1 <= r.pp.v.i0) { fp.p[k].n_tiles = 0; continue; }
        PlateParams pp = r.pp;  // per-launch copy: slot maps in element offsets of this batch's leading dimension
        for (int l = 0; l < pp.sh.nl; ++l) {
            pp.v.loc[l].row0 = pp.v.loc[l].base * ld;
            pp.v.loc[l].step_i = pp.v.loc[l].si * ld;
            pp.v.loc[l].step_j = pp.v.loc[l].sj * ld;
        }
        bool jit = (r.variant == 0 || is_tape_variant(r.variant)) && r.jit && r.jit_valid && !p->force_dynamic;
        int variant = r.variant;
        if (p->parts == 1) {
            // prior part only: the interpreter with the observation switched off (has_obs = 0 is a run-time shape field)
            jit = false;
            variant = pp.sh.n_x > 0 ? (pp.sh.n_gref > MAX_GREF_LINEAR ? g_xwspec : g_xspec) : 0;
            pp.sh.has_obs = 0;
            fp.p[k].obs_const = 0.0;
        }
        if (!jit && variant == 0) fp.p[k].obs_const = 0.0;  // the interpreter evaluates the full log-density itself
        cudaError_t le = jit ? jit_launch(*r.jit, r.n_tiles, r.smem_bytes, st, pp, theta, grad, ld, C, p->d_gvals, ldg,
                                          p->d_partial + r.partial_offset * ldg, p->d_flags, want_grad && grad, p->use_pdl)
                             : g_specs[variant].launch(r.n_tiles, r.smem_bytes + (size_t)g_specs[variant].smpf_doubles * 8, st, pp, theta, grad, ld, C, p->d_gvals, ldg,
                                                       p->d_partial + r.partial_offset * ldg, p->d_flags, want_grad && grad, p->use_pdl);
        if (le != cudaSuccess) return fail(JBUGS_ERR_CUDA, "plate kernel launch failed: %s", cudaGetErrorString(le));
        p->launches++;
        reduce_plates.push_back((int)k);
    }
    if (!p->dense.empty() && p->parts != 1) {  // (the dense predictor is all likelihood: no part of the prior pass)
        const jbugs_dense& dn = p->dense[0];
        const bool sharded = p->comm && p->nranks > 1 && p->dense_i1 >= 0;  // rows of X split over the ranks
        const long long r0 = p->dense_i0, Nloc = (p->dense_i1 >= 0 ? p->dense_i1 : dn.N) - r0;
        const unsigned col_blocks = (unsigned)((C + DG_BN - 1) / DG_BN);
        static bool opted_dev[64] = {false};  // per device: a process may hold plans on several GPUs
        bool& opted = opted_dev[p->device & 63];
        if (!opted) {
            CU(cudaFuncSetAttribute(dense_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DgFwd::SMEM_BYTES));
            CU(cudaFuncSetAttribute(dense_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DgBwd::SMEM_BYTES));
            opted = true;
        }
        const double* X = (const double*)p->d_slots[dn.x_slot] + r0;  // column-major N x P: a row range is an offset
        DenseFwdParams fw{};
        fw.op.A = X; fw.op.lda = dn.N;
        fw.op.B = theta + dn.theta_base * ld; fw.op.ldb = dn.theta_stride * ld;
        fw.op.M = Nloc; fw.op.N = C; fw.op.K = dn.P;
        const bool x_vec = ((uintptr_t)X % 16 == 0) && (dn.N % 2 == 0);
        fw.op.vec_ok = x_vec && ((uintptr_t)theta % 16 == 0) && (ld % 2 == 0);
        fw.y = (const double*)p->d_slots[dn.y_slot] + r0;
        fw.n = dn.n_slot >= 0 ? (const double*)p->d_slots[dn.n_slot] + r0 : nullptr;
        fw.R = p->d_R; fw.ldr = ldg;
        fw.partial = p->d_partial + p->dense_partial_offset * ldg; fw.ldg = ldg;
        fw.flags = p->d_flags;
        dense_fwd_kernel<<<dim3(col_blocks, (unsigned)p->dense_row_groups), DgFwd::THREADS, DgFwd::SMEM_BYTES, st>>>(fw);
        p->launches++;
        const int k_fp = (int)p->rt.size();
        fp.p[k_fp].n_tiles = p->dense_row_groups;
        fp.p[k_fp].n_gref = 0;
        fp.p[k_fp].offset = p->dense_partial_offset;
        fp.p[k_fp].replicated = (p->comm && p->nranks > 1 && !sharded) ? 1 : 0;
        if (p->dense_const_dirty) {
            // data-only normaliser of the binomial (0 for Bernoulli) + validation of the counts, once per upload
            double h2[2] = {0.0, 0.0};
            obs_const_kernel<<<1, 1024, 0, st>>>(dn.n_slot >= 0 ? FUSED_BINOMIAL_LOGIT : FUSED_BERNOULLI_LOGIT, fw.y, fw.n, 0,
                                                 JBUGS_ARG_DATA_I, 1.0, Nloc, 1, 0, Nloc, p->d_red);
            p->launches++;
            CU(cudaMemcpyAsync(h2, p->d_red, sizeof h2, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            if (h2[1] != 0.0) return fail(JBUGS_ERR_UNSUPPORTED, "dense predictor: observations outside the support of the family");
            p->dense_const = h2[0];
            p->dense_const_dirty = false;
            p->epoch++;
        }
        fp.p[k_fp].obs_const = p->dense_const;
        fp.n_plates = k_fp + 1;
        if (want_grad && grad) {
            DenseBwdParams bw{};
            bw.op.A = X; bw.op.lda = dn.N;
            bw.op.B = p->d_R; bw.op.ldb = ldg;
            bw.op.M = dn.P; bw.op.N = C; bw.op.K = Nloc;
            bw.op.vec_ok = x_vec;
            bw.Gp = p->d_Gp; bw.ldg = ldg;
            bw.k_per_split = ((Nloc + p->dense_splits - 1) / p->dense_splits + DgBwd::BK - 1) / DgBwd::BK * DgBwd::BK;
            dense_bwd_kernel<<<dim3(col_blocks, (unsigned)((dn.P + DG_BM - 1) / DG_BM), (unsigned)p->dense_splits), DgBwd::THREADS, DgBwd::SMEM_BYTES, st>>>(bw);
            p->launches++;
            if (sharded) {
                // this rank's rows of X^T R go into the all-reduce buffer (rows 2 + NG ...); added to grad afterwards
                const long long NGr = std::max<long long>(p->h.n_globals + p->h.n_gtape, 1);
                dense_grad_reduce_kernel<<<dim3(chain_tiles, (unsigned)dn.P), BLOCK, 0, st>>>(p->d_Gp, p->dense_splits, dn.P, ldg, p->d_red, ldg,
                                                                                           2 + NGr, 1, C, 0);
            } else {
                dense_grad_reduce_kernel<<<dim3(chain_tiles, (unsigned)dn.P), BLOCK, 0, st>>>(p->d_Gp, p->dense_splits, dn.P, ldg, grad, ld,
                                                                                           dn.theta_base, dn.theta_stride, C, 1);
            }
            p->launches++;
        }
    }
    if (p->timing) CU(cudaEventRecord(tev[2], st));
    const double* partial_in = p->d_partial;
    {
        // many tiles: first level of the fixed-order tree (super-tiles), so that finalize's serial loop stays short
        bool any = false;
        for (int k : reduce_plates) any |= p->rt[k].n_tiles > 2 * REDUCE_R;
        if (any) {
            for (int k : reduce_plates) {
                PlateRt& r = p->rt[k];
                const int rows = 1 + r.pp.sh.n_gref;
                dim3 g2(chain_tiles, (unsigned)rows, REDUCE_R);
                reduce_tiles_kernel<<<g2, BLOCK, 0, st>>>(p->d_partial + r.partial_offset * ldg, p->d_partial2 + r.partial2_offset * ldg,
                                                          (int)r.n_tiles, rows, C, ldg);
                p->launches++;
                fp.p[k].n_tiles = REDUCE_R;
                fp.p[k].offset = r.partial2_offset;
            }
            partial_in = p->d_partial2;
        }
    }
    if (p->comm && p->nranks > 1) {
        // observation-sharded: reduce this rank's tiles, all-reduce [logp part, global adjoints, flag] over NVLink,
        // then add the priors of the globals once and finish
        const int NG = (int)(p->h.n_globals + p->h.n_gtape);
        launch_helper(finalize_kernel, p->use_pdl, helper_tiles, hblk, st, p->gp, p->tp, fp, theta, grad, ld, C, p->d_gvals, ldg, partial_in,
                                                       p->d_flags, logp, status, p->d_any_bad, want_grad && grad, 1, p->d_red);
        p->launches++;
        const bool dense_grad = !p->dense.empty() && p->dense_i1 >= 0 && want_grad && grad;
        const long long red_rows = dense_grad ? 2 + std::max(NG, 1) + p->dense[0].P : 2 + NG;
        const int e = g_nccl.AllReduce(p->d_red, p->d_red, (size_t)red_rows * ldg, /*ncclDouble*/ 8, /*ncclSum*/ 0, p->comm, st);
        if (e) return fail(JBUGS_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?");
        if (dense_grad) {
            const jbugs_dense& dn = p->dense[0];
            dense_grad_reduce_kernel<<<dim3(chain_tiles, (unsigned)dn.P), BLOCK, 0, st>>>(p->d_red + (2 + std::max(NG, 1)) * ldg, 1, dn.P, ldg, grad, ld,
                                                                                       dn.theta_base, dn.theta_stride, C, 1);
            p->launches++;
        }
        launch_helper(finalize_kernel, p->use_pdl, helper_tiles, hblk, st, p->gp, p->tp, fp, theta, grad, ld, C, p->d_gvals, ldg, partial_in,
                                                       p->d_flags, logp, status, p->d_any_bad, want_grad && grad, 2, p->d_red);
        p->launches++;
    } else {
        launch_helper(finalize_kernel, p->use_pdl, helper_tiles, hblk, st, p->gp, p->tp, fp, theta, grad, ld, C, p->d_gvals, ldg, partial_in,
                                                       p->d_flags, logp, status, p->d_any_bad, want_grad && grad, 0, p->d_red);
        p->launches++;
    }
    if (want_grad && grad && fp.poison_rows == 0) {
        // almost always a no-op (no chain was flagged): keep the grid near one wave, rows are strided over it
        const long long rows = std::max<long long>(2LL * p->num_sms / chain_tiles, 1);
        dim3 grid(chain_tiles, (unsigned)std::min<long long>(std::max<long long>(p->h.D, 1), rows));
        poison_kernel<<<grid, BLOCK, 0, st>>>(grad, ld, C, p->h.D, p->d_flags, p->d_any_bad);
        p->launches++;
    }
    if (p->timing) {
        CU(cudaEventRecord(tev[3], st));
        p->events_recorded = true;
        if (!p->ring.empty()) {
            for (int q = 0; q < 4; ++q) p->ev_last[q] = tev[q];
            p->ring_n++;
        } else {
            for (int q = 0; q < 4; ++q) p->ev_last[q] = p->ev[q];
        }
    }
    CU(cudaGetLastError());
    return 0;
}

static bool is_device_ptr(const void* ptr) {
    if (!ptr) return true;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

#include "jbugs_staged.inc"

extern "C" int jbugs_eval_batch(jbugs_plan* p, const double* theta, int64_t ld, int64_t C, int layout, double* logp,
                                double* grad, int32_t* status, int want_grad, int flags, void* stream) {
    if (!p || !theta || !logp) return fail(JBUGS_ERR_INVALID, "null argument");
    if (C < 0) return fail(JBUGS_ERR_INVALID, "negative chain count");
    if (want_grad && !grad) return fail(JBUGS_ERR_INVALID, "want_grad set but grad is null");
    if (layout != JBUGS_LAYOUT_CHAIN_FAST && layout != JBUGS_LAYOUT_PARAM_FAST) return fail(JBUGS_ERR_INVALID, "unknown layout %d", layout);
    const long long D = p->h.D;
    if (layout == JBUGS_LAYOUT_CHAIN_FAST ? ld < C : ld < D) return fail(JBUGS_ERR_INVALID, "leading dimension %lld too small", (long long)ld);
    int rc = set_device(p);
    if (rc) return rc;
    if ((rc = jit_poll(p))) return rc;  // an asynchronous specialisation that has finished takes over from here on
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 0) return JBUGS_OK;
    const bool dev_io = is_device_ptr(theta) && is_device_ptr(logp) && is_device_ptr(grad) && is_device_ptr(status);
    if (dev_io && layout == JBUGS_LAYOUT_CHAIN_FAST) {
        rc = eval_device(p, theta, ld, C, logp, grad, status, want_grad, st);
        if (rc) return rc;
        if (!(flags & JBUGS_EVAL_ASYNC)) {
            CU(cudaStreamSynchronize(st));
            if (p->timing) {
                cudaEventElapsedTime(&p->last_plate_ms, p->ev_last[1], p->ev_last[2]);
                cudaEventElapsedTime(&p->last_total_ms, p->ev_last[0], p->ev_last[3]);
            }
        }
        return JBUGS_OK;
    }
    // host buffers and/or PARAM_FAST: pipelined through library staging in chunks of chains
    return eval_staged(p, theta, ld, C, layout, logp, grad, status, want_grad, st, (flags & JBUGS_EVAL_ASYNC) != 0);
}

extern "C" int jbugs_plan_synchronize(jbugs_plan* p, void* stream) {
    if (!p) return fail(JBUGS_ERR_INVALID, "null plan");
    int rc = set_device(p);
    if (rc) return rc;
    if (p->st_out) {
        CU(cudaStreamSynchronize(p->st_in));
        CU(cudaStreamSynchronize(p->st_out));
    }
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return JBUGS_OK;
}

// ---- logprior / loglikelihood / tempered log joint (evaluation.jl:116, 269-274) -------------------------------------
__global__ void __launch_bounds__(256)
temper_kernel(double* __restrict__ a /* total -> tempered */, const double* __restrict__ b /* prior part */, long long rows,
              long long cols, long long ld, double t) {
    const long long q = (long long)blockIdx.x * 256 + threadIdx.x, r = blockIdx.y;
    if (q < cols && r < rows) {
        const double tot = a[r * ld + q], pr = b[r * ld + q];
        a[r * ld + q] = fma(t, tot - pr, pr);
    }
}
__global__ void __launch_bounds__(256)
temper_logp_kernel(double* __restrict__ total, double* __restrict__ prior, double* __restrict__ lik, const int* __restrict__ status,
                   long long C, double t) {
    const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
    if (c >= C) return;
    const double tot = total[c], pr = prior[c];
    const bool bad = status && status[c] != 0;
    const double ll = bad ? dev_nan() : tot - pr;
    if (lik) lik[c] = ll;
    total[c] = bad ? -dev_inf() : fma(t, ll, pr);
    if (bad) prior[c] = dev_nan();
}

extern "C" int jbugs_eval_batch_tempered(jbugs_plan* p, const double* theta, int64_t ld, int64_t C, int layout, double temperature,
                                         double* logprior, double* loglik, double* logp, double* grad, int32_t* status,
                                         int want_grad, void* stream) {
    if (!p || !theta || !logp || !logprior) return fail(JBUGS_ERR_INVALID, "null argument");
    if (want_grad && !grad) return fail(JBUGS_ERR_INVALID, "want_grad set but grad is null");
    if (p->comm && p->nranks > 1) return fail(JBUGS_ERR_UNSUPPORTED, "the prior / likelihood split is not available on an observation-sharded plan");
    const bool dev = is_device_ptr_strict(theta);
    if (dev != is_device_ptr_strict(logp) || dev != is_device_ptr_strict(logprior) || (grad && dev != is_device_ptr_strict(grad)) ||
        (loglik && dev != is_device_ptr_strict(loglik)) || (status && dev != is_device_ptr_strict(status)))
        return fail(JBUGS_ERR_INVALID, "buffers must all be host pointers or all be device pointers");
    // pass 1: everything -> (logp, grad, status); pass 2: the prior part -> (logprior, scratch gradient)
    int rc = jbugs_eval_batch(p, theta, ld, C, layout, logp, grad, status, want_grad, 0, stream);
    if (rc) return rc;
    if (C == 0) return JBUGS_OK;
    const long long D = p->h.D;
    const long long rows = layout == JBUGS_LAYOUT_CHAIN_FAST ? D : C, cols = layout == JBUGS_LAYOUT_CHAIN_FAST ? C : D;
    const size_t gbytes = (size_t)std::max<long long>(rows, 1) * (size_t)ld * 8;
    double* g2 = nullptr;
    if (want_grad) {
        if (dev) CU(cudaMalloc(&g2, gbytes));
        else if (!(g2 = (double*)malloc(gbytes))) return fail(JBUGS_ERR_STATE, "out of host memory");
    }
    p->parts = 1;
    rc = jbugs_eval_batch(p, theta, ld, C, layout, logprior, g2, nullptr, want_grad, 0, stream);
    p->parts = 0;
    p->epoch++;
    if (!rc) {
        if (dev) {
            cudaStream_t st = (cudaStream_t)stream;
            if (want_grad) temper_kernel<<<dim3((unsigned)((cols + 255) / 256), (unsigned)rows), 256, 0, st>>>(grad, g2, rows, cols, ld, temperature);
            temper_logp_kernel<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(logp, logprior, loglik, status, C, temperature);
            p->launches += want_grad ? 2 : 1;
            if (cudaStreamSynchronize(st) != cudaSuccess) rc = fail(JBUGS_ERR_CUDA, "tempering kernels failed");
        } else {
            if (want_grad)
                for (long long r = 0; r < rows; ++r)
                    for (long long q = 0; q < cols; ++q) {
                        const double tot = grad[r * ld + q], pr = g2[r * ld + q];
                        grad[r * ld + q] = std::fma(temperature, tot - pr, pr);
                    }
            for (long long c = 0; c < C; ++c) {
                const bool bad = status && status[c] != 0;
                const double ll = bad ? NAN : logp[c] - logprior[c];
                if (loglik) loglik[c] = ll;
                logp[c] = bad ? -INFINITY : std::fma(temperature, ll, logprior[c]);
                if (bad) logprior[c] = NAN;
            }
        }
    }
    if (dev) cudaFree(g2);
    else free(g2);
    return rc;
}

#include "jbugs_hmc.inc"
#include "jbugs_nuts.inc"
#include "jbugs_jit.inc"
#include "jbugs_gq.inc"

extern "C" int jbugs_batch_alloc(jbugs_plan* p, int64_t C, double** theta_dev, double** grad_dev, double** logp_dev,
                                 int32_t** status_dev, int64_t* ld) {
    if (!p || !theta_dev || !grad_dev || !logp_dev || !ld) return fail(JBUGS_ERR_INVALID, "null argument");
    if (C <= 0) return fail(JBUGS_ERR_INVALID, "C must be positive");
    int rc = set_device(p);
    if (rc) return rc;
    jbugs_batch_free(p);
    const long long ldc = round_up(C, 16);
    const size_t bytes = (size_t)std::max<long long>(p->h.D, 1) * ldc * 8;
    CU(cudaMalloc(&p->b_theta, bytes));
    CU(cudaMalloc(&p->b_grad, bytes));
    CU(cudaMalloc(&p->b_logp, (size_t)ldc * 8));
    CU(cudaMalloc(&p->b_status, (size_t)ldc * 4));
    *theta_dev = p->b_theta;
    *grad_dev = p->b_grad;
    *logp_dev = p->b_logp;
    if (status_dev) *status_dev = p->b_status;
    *ld = ldc;
    return JBUGS_OK;
}

// ------------------------------------------------------------------------------------------------ comm
extern "C" int jbugs_comm_unique_id(void* out) {
    if (!out) return fail(JBUGS_ERR_INVALID, "null argument");
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    int e = g_nccl.GetUniqueId(&id);
    if (e) return fail(JBUGS_ERR_NCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?");
    memcpy(out, &id, sizeof id);
    return JBUGS_OK;
}

extern "C" int jbugs_comm_init(jbugs_plan* p, int nranks, int rank, const void* unique_id) {
    if (!p || !unique_id) return fail(JBUGS_ERR_INVALID, "null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(JBUGS_ERR_INVALID, "bad rank %d of %d", rank, nranks);
    int rc = load_nccl();
    if (rc) return rc;
    rc = set_device(p);
    if (rc) return rc;
    jbugs_comm_destroy(p);
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof id);
    int e = g_nccl.CommInitRank(&p->comm, nranks, id, rank);
    if (e) return fail(JBUGS_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?");
    p->nranks = nranks;
    p->rank = rank;
    return JBUGS_OK;
}
