/*
 * jbugs_cuda.h -- C ABI of libjbugs_cuda, the B200 (sm_100a) batched evaluator for
 * `LogDensityProblems.logdensity` / `logdensity_and_gradient` on a compiled BUGS model.
 *
 * Each entry point names the reference interface it stands in for (paths relative to
 * /root/reference/JuliaBUGS).  The reference evaluates ONE theta per call on the CPU; this
 * library evaluates a batch of C parameter vectors (HMC chains) per launch.
 *
 * Conventions
 *   - every function returns 0 on success or a negative jbugs_status; nothing throws across the ABI;
 *     jbugs_last_error() returns a thread-local message for the last failure.
 *   - plain pointers and sizes only.  `theta`, `grad`, `logp`, `status` may be HOST or DEVICE
 *     pointers (detected with cudaPointerGetAttributes).  Host buffers are copied through
 *     library-owned staging inside the call; device buffers are used in place.
 *   - the library owns the device copies of the model data, and (jbugs_batch_alloc) the theta /
 *     gradient batch.  Caller-owned memory is only read/written during a call.
 *   - a plan handle is not thread-safe; distinct handles are.
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point fails
 *     with JBUGS_ERR_CUDA.
 */
#ifndef JBUGS_CUDA_H
#define JBUGS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#include "jbugs_plan.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jbugs_plan_s jbugs_plan;

enum jbugs_status {
    JBUGS_OK = 0,
    JBUGS_ERR_INVALID = -1,     /* malformed plan / bad argument            */
    JBUGS_ERR_UNSUPPORTED = -2, /* valid plan the kernels cannot run (yet)  */
    JBUGS_ERR_CUDA = -3,        /* CUDA runtime failure (incl. "no device") */
    JBUGS_ERR_NCCL = -4,        /* NCCL failure / libnccl not loadable      */
    JBUGS_ERR_STATE = -5        /* e.g. data slot not uploaded              */
};

/* theta/grad batch layouts.  D = jbugs_dim(plan), C = number of chains.
 *   CHAIN_FAST : element (d, c) at  d * ld + c   (ld >= C)  -- the library's native layout:
 *                every parameter is a contiguous row over chains (coalesced for any theta order).
 *   PARAM_FAST : element (d, c) at  c * ld + d   (ld >= D)  -- a Julia `Matrix{Float64}(D, C)`,
 *                one column per chain, what `logdensity_and_gradient(m, x)` sees for C = 1. */
enum jbugs_layout { JBUGS_LAYOUT_CHAIN_FAST = 0, JBUGS_LAYOUT_PARAM_FAST = 1 };

/* per-chain status written by jbugs_eval_batch */
enum jbugs_chain_status {
    JBUGS_CHAIN_OK = 0,
    JBUGS_CHAIN_DOMAIN = 1 /* a distribution argument left its domain: logp = -Inf, grad = NaN
                              (logdensityproblems.jl:22-26, 226-229) */
};

const char* jbugs_last_error(void);
int jbugs_version(void);

/* Number of CUDA devices visible (0 on a CPU-only box; never fails). */
int jbugs_device_count(void);

/* Replaces: set_evaluation_mode(model, mode) building the evaluator state
 * (src/model/bugsmodel.jl:736-801) from the lowered program (src/source_gen.jl:459-552).
 * `plan_blob` is the serialized jbugs_plan_header + records (jbugs_plan.h); it is copied. */
int jbugs_plan_create(const void* plan_blob, size_t nbytes, int device, jbugs_plan** out);

/* Replaces: dropping `log_density_computation_function` on condition/fix/decondition
 * (src/model/abstractppl.jl:796-799).  Frees every device buffer the plan owns. */
int jbugs_plan_destroy(jbugs_plan* plan);

/* Replaces: LogDensityProblems.dimension(model) (src/model/logdensityproblems.jl:29-51). */
int64_t jbugs_dim(const jbugs_plan* plan);

/* Replaces: the `evaluation_env` argument of the generated function (src/source_gen.jl:673-674)
 * and `set_observed_values!` (src/model/abstractppl.jl:872-888): (re)uploads one data slot.
 * `src` is a host or device pointer to `nbytes` bytes of float64 (dtype 0) or int64 (dtype 1,
 * index arrays). */
int jbugs_plan_upload_data(jbugs_plan* plan, int slot, const void* src, size_t nbytes, int dtype);

/* Observation sharding: this rank owns groups [i0, i1) of plate `plate`; group-local
 * theta slots outside the range are neither read nor written.  Globals are replicated.  With a
 * communicator attached (jbugs_comm_init) the per-chain logp and global gradients are summed
 * over ranks in a fixed order.  Default: the whole plate. */
int jbugs_plan_set_shard(jbugs_plan* plan, int plate, int64_t i0, int64_t i1);

/* Same for the dense predictor `inprod(X[i, 1:P], beta[1:P])`: this rank
 * owns rows [i0, i1) of X, y (and n); i0 even keeps the 16-byte loads.  The coefficient vector and
 * its prior plate stay replicated; the P x C block X^T R travels in the same all-reduce as the
 * per-chain logp and the global adjoints.  A plate (or the dense predictor) for which no shard
 * was set is treated as replicated: evaluated on every rank, counted once. */
int jbugs_plan_set_dense_shard(jbugs_plan* plan, int64_t i0, int64_t i1);

/* Replaces: LogDensityProblems.logdensity (src/model/logdensityproblems.jl:19-27) and
 * logdensity_and_gradient (:219-232) -> evaluate_with_values!! (src/model/evaluation.jl:217-275)
 * plus the AD backend, for C chains at once.
 *   theta  : D x C batch in `layout` with leading dimension `ld` (read only)
 *   logp   : C doubles out
 *   grad   : same layout/ld as theta, out (ignored when want_grad == 0; may be NULL then)
 *   status : C int32 out (jbugs_chain_status), may be NULL
 *   stream : cudaStream_t (NULL = default stream).  The call returns after the stream has been
 *            synchronized unless flags & JBUGS_EVAL_ASYNC.  With device buffers an ASYNC call only enqueues work
 *            on `stream`.  With host buffers an ASYNC call enqueues the whole staged pipeline (copy-in, kernels,
 *            copy-out on the library's streams) and returns: consecutive ASYNC calls run as ONE pipeline, the
 *            copy-in of a call overlapping the kernels and copy-out of the one before; the host buffers must
 *            be ready when the call is made, stay untouched until jbugs_plan_synchronize, and hold the results
 *            after it.
 *            The launch sequence contains no host synchronisation once the plan is warm, so an
 *            ASYNC call on a non-default stream can be captured into a CUDA graph by the caller
 *            (the on-device sampler does exactly that, jbugs_hmc_run). */
#define JBUGS_EVAL_ASYNC 1
int jbugs_eval_batch(jbugs_plan* plan, const double* theta, int64_t ld, int64_t C, int layout,
                     double* logp, double* grad, int32_t* status, int want_grad, int flags,
                     void* stream);

/* Waits for everything earlier (ASYNC) evaluations of this plan put on `stream` and on the library's copy streams. */
int jbugs_plan_synchronize(jbugs_plan* plan, void* stream);

/* Replaces: the `(logprior, loglikelihood, tempered_logjoint)` triple of `evaluate_with_values!!`
 * (src/model/evaluation.jl:116, 269-274; `AbstractPPL.evaluate!!`, src/model/abstractppl.jl:989-1005), for C chains:
 *   logprior : densities of the unobserved stochastic nodes + log-Jacobians of their bijectors
 *   loglik   : densities of the observed nodes (may be NULL)
 *   logp     : logprior + temperature * loglik;  grad : its gradient (same layout as theta)
 * Two evaluations: everything, then the prior part alone (observations switched off); the likelihood part is their
 * difference.  Buffers are all host pointers or all device pointers.  DomainError chains: logp = -Inf, the parts NaN. */
int jbugs_eval_batch_tempered(jbugs_plan* plan, const double* theta, int64_t ld, int64_t C, int layout,
                              double temperature, double* logprior, double* loglik, double* logp, double* grad,
                              int32_t* status, int want_grad, void* stream);

/* Library-owned theta / gradient batch (north_star: "owns the device buffers for data, the theta
 * batch and the gradients").  Returns device pointers in CHAIN_FAST layout with *ld >= C padded so
 * that every row is 128-byte aligned. */
int jbugs_batch_alloc(jbugs_plan* plan, int64_t C, double** theta_dev, double** grad_dev,
                      double** logp_dev, int32_t** status_dev, int64_t* ld);
int jbugs_batch_free(jbugs_plan* plan);

/* Pinned host memory for callers that stage from the host (bench e2e path).  _ex: flags = JBUGS_HOST_WRITE_COMBINED
 * for buffers the host only writes and the device only reads (theta): no cache snooping on the PCIe read path. */
#define JBUGS_HOST_WRITE_COMBINED 1
int jbugs_host_alloc(void** ptr, size_t nbytes);
int jbugs_host_alloc_ex(void** ptr, size_t nbytes, int flags);
int jbugs_host_free(void* ptr);

/* Bind the calling process to the CPUs and (preferred) memory of the NUMA node of `device` (sysfs topology), so that
 * pinned staging buffers allocated afterwards sit next to the GPU's PCIe root.  *node_out = -1 when the platform
 * reports no node.  One process per GPU: call once, before allocating host buffers. */
int jbugs_bind_host_to_device(int device, int* node_out, int* n_cpus_out);

/* Multi-GPU (one process per GPU).  Replaces: AbstractMCMC MCMCThreads/MCMCDistributed chain
 * parallelism (docs/src/inference/parallel.md) -- which needs nothing here, chains shard with no
 * collective -- and adds observation sharding, which needs one all-reduce of C x (1 + G) doubles.
 * `unique_id` is the 128-byte ncclUniqueId produced by rank 0 with jbugs_comm_unique_id. */
int jbugs_comm_unique_id(void* unique_id_128);
int jbugs_comm_init(jbugs_plan* plan, int nranks, int rank, const void* unique_id_128);
int jbugs_comm_destroy(jbugs_plan* plan);

/* Introspection used by tests/bench: number of kernels launched so far by this plan, and a
 * human-readable description of the kernel variants selected for each plate. */
int64_t jbugs_launch_count(const jbugs_plan* plan);
int jbugs_plan_describe(const jbugs_plan* plan, char* buf, size_t nbuf);

/* Run-time specialisation (optional; the analogue of the reference generating and compiling a Julia function
 * per model, J/src/model/bugsmodel.jl:736-801).  Plates whose shape has no in-tree specialised kernel run the
 * interpreter instantiations; this call writes the constexpr description of each such plate -- and, for an
 * expression plate, its tape as straight-line code -- as C++ source and compiles the SAME kernel body for it
 * in-process with NVRTC (libnvrtc is dlopen'ed: $JBUGS_NVRTC, libnvrtc.so.12, /usr/local/cuda/lib64; the
 * kernel headers are embedded in this library, so neither nvcc nor a source tree is needed).  Cubins are cached
 * in `cache_dir` (NULL: $JBUGS_JIT_CACHE, $XDG_CACHE_HOME/jbugs_jit or ~/.cache/jbugs_jit; "none": no cache),
 * created 0700 and used only if owned by the caller and not writable by others.  Returns JBUGS_ERR_UNSUPPORTED
 * when libnvrtc cannot be loaded; nothing changes then.  Tiered: the call returns with a quick-to-compile variant of
 * each kernel (~3 s cold) and a worker thread compiles the best variant, which takes over at a later evaluation
 * (jbugs_plan_describe shows `..._quick_variant` until then).  jbugs_jit_last_compile_seconds: time NVRTC took in
 * the last call (0 when every kernel came from the cache). */
int jbugs_plan_specialize(jbugs_plan* plan, const char* cache_dir);
double jbugs_jit_last_compile_seconds(void);
/* The same without blocking: NVRTC runs on a worker thread, evaluations keep using the interpreter instantiations
 * and switch to the specialised kernels at the first jbugs_eval_batch after they are ready (the two agree to rounding,
 * not bit for bit).  Cached kernels are adopted at once; jbugs_plan_specialize() afterwards waits for the worker. */
int jbugs_plan_specialize_async(jbugs_plan* plan, const char* cache_dir);

/* Tuning knobs (tests force the generic interpreter path with "dynamic=1"). Unknown keys fail. */
int jbugs_plan_set_option(jbugs_plan* plan, const char* key, int64_t value);

/* Device-side timing of the last synchronous jbugs_eval_batch on this plan: milliseconds spent in the
 * plate kernels and in total between the first and last enqueue (CUDA events on the call's stream).
 * Recorded only after jbugs_plan_set_option(plan, "timing", 1); off by default because the four
 * event records are a measurable share of a small model's evaluation. */
int jbugs_last_timing(const jbugs_plan* plan, float* plate_ms, float* total_ms);
/* After jbugs_plan_set_option(plan, "timing", 2) every evaluation records into its own events (a ring of 1024):
 * plate_ms[k] / total_ms[k] of the (up to `cap`) most recent evaluations since the option was set, oldest first.
 * Waits for the last of them. */
int jbugs_timing_history(const jbugs_plan* plan, float* plate_ms, float* total_ms, int cap, int* n_out);

/* FP64-pipe peak of `device` in fused multiply-adds per second, measured with independent DFMA chains (the bound of the
 * logit / log-link plate kernels; bench.py reports them against it). */
int jbugs_fp64_peak(int device, double* fma_per_s);

/* ---- on-device batched HMC (the caller of the hot path) ---------------------------------------
 * Replaces: the AdvancedHMC loop behind `AbstractMCMC.sample(model, HMC/NUTS, n)`
 * (ext/JuliaBUGSAdvancedHMCExt.jl:40-66), which calls logdensity_and_gradient once per leapfrog step with theta on the
 * host.  C chains advance in lock step; theta, momentum and gradient stay in HBM (5 x D x C doubles).
 *   jbugs_hmc_init : every chain starts at theta0 + jitter * N(0,1) (unconstrained space)
 *   jbugs_hmc_run  : n_iter transitions of n_leapfrog steps.  adapt != 0: warm-up (dual averaging of a per-chain step
 *                    size on the device, windowed diagonal metric from moments pooled over chains).  The monitored theta rows of
 *                    every iteration are written to out_host[n_iter][n_monitor][C]. */
typedef struct jbugs_hmc_s jbugs_hmc;
int jbugs_hmc_create(jbugs_plan* plan, int64_t C, uint64_t seed, jbugs_hmc** out);
int jbugs_hmc_destroy(jbugs_hmc* hmc);
int jbugs_hmc_init(jbugs_hmc* hmc, const double* theta0_host, double jitter, double step_size);
int jbugs_hmc_run(jbugs_hmc* hmc, int n_iter, int n_leapfrog, int adapt, const int64_t* monitor_rows, int n_monitor,
                  double* out_host, double* mean_accept);
int jbugs_hmc_get_state(jbugs_hmc* hmc, double* theta_host /* [D][C] */, double* logp_host /* [C] */, double* step_size,
                        double* var_host /* [D] diagonal of the inverse metric */);
/*   jbugs_hmc_draw_momentum : one momentum refresh p ~ N(0, M) exactly as a transition draws it (Philox counter =
 *                    iteration `it`), copied to p_host[D][C]. */
int jbugs_hmc_draw_momentum(jbugs_hmc* hmc, uint32_t it, double* p_host);
/*   jbugs_hmc_run_nuts : same, with the No-U-Turn transition the reference uses (`NUTS(0.8)`): multinomial sampling
 *                    along the trajectory, generalised U-turn criterion, at most max_depth (<= 12) doublings.  All chains
 *                    build their trees in lock step; a chain that has finished waits for the deepest tree of the batch.
 *                    Tree state: (12 + 2 x 12) x D x C doubles, so this is for small and medium models.
 *   jbugs_hmc_nuts_stats : divergent transitions so far, mean tree depth of the last transition. */
int jbugs_hmc_run_nuts(jbugs_hmc* hmc, int n_iter, int max_depth, int adapt, const int64_t* monitor_rows, int n_monitor,
                       double* out_host, double* mean_accept);
int jbugs_hmc_nuts_stats(jbugs_hmc* hmc, int64_t* n_divergent, double* mean_depth);

#ifdef __cplusplus
}
#endif
#endif /* JBUGS_CUDA_H */
