// jbugs_plate.cuh -- the plate kernel: fused forward + reverse over one tile of groups for BLOCK chains.
// (see the mapping notes at the top of jbugs_kernels.cuh)
//
// Structure of one CTA:
//   tile of groups [ib, ie)  ->  chunks of STAGE_G groups
//     chunk k+1's observation / covariate streams are copied global -> shared with cp.async while chunk k is
//     being evaluated (two stages, one __syncthreads per chunk);
//     inside a chunk the group loop runs UNROLL groups at a time: all theta loads of the UNROLL groups are issued
//     before any is used, then each group is evaluated from registers + shared-memory broadcasts and its
//     gradient rows are stored.
#pragma once

namespace jbugs {

// Per-thread accumulators of the fused (sufficient-statistic) paths.  Unused members are dead code
// in instantiations that do not take the corresponding path.
struct FusedAcc {
    double ss;            // sum r^2 over the tile's observations            (normal / identity)
    double sg[MAX_GT];    // sum r * cov_t for the global terms               (normal / identity)
    double pss[MAX_LOC];  // sum (x_l - mu_l)^2 over the tile                 (normal prior, invariant args)
    double ps[MAX_LOC];   // sum (x_l - mu_l)
    double pd[MAX_LOC];   // x_l - mu_l of the instance being evaluated (its prior adjoint is applied at the store)
};

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// ---- TMA (bulk async copy) + mbarrier: the staging path for 16-byte aligned chunks -------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
// global -> shared bulk copy (the 1-D TMA path, SASS UBLKCP); completion is signalled on the mbarrier as transaction bytes
__device__ __forceinline__ void bulk_g2s(double* smem_dst, const double* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

// copy the streams of groups [i, i + cnt) into one stage.  Returns true when the chunk went through the TMA path
// (then `bar` completes when the bytes have landed); otherwise every thread issued its share with cp.async.
template <int BLK, bool TMA>
__device__ __forceinline__ bool stage_chunk(const PlateVals& V, double* dst, long long i, int cnt, int T, int tid,
                                            unsigned long long* bar) {
    // warp 0 drives the copy engine when source, destination and size are 16-byte multiples
    if (TMA && V.bulk_ok && V.per_group > 0 && ((i & 1) == 0) && ((cnt & 1) == 0)) {
        if (tid < 32) {
            // warp 0 issues the copies lane-parallel (copy q = one row of `cnt` doubles: a DATA_I stream has one row,
            // a DATA_IJ stream T rows), so that no single thread is held up by the whole list
            const unsigned row_bytes = (unsigned)cnt * 8u;
            if (tid == 0) mbar_expect_tx(bar, row_bytes * (unsigned)V.per_group);
            __syncwarp();
            for (int q = tid; q < V.per_group; q += 32) {
                int s = 0, base = 0;
                for (; s < V.n_streams; ++s) {
                    const int rows = V.streams[s].kind == JBUGS_ARG_DATA_IJ ? T : (V.streams[s].kind == JBUGS_ARG_DATA_I ? 1 : 0);
                    if (q < base + rows) break;
                    base += rows;
                }
                const Stream sm = V.streams[s];
                const int j = q - base;
                bulk_g2s(dst + sm.soff + j * STAGE_G, sm.ptr + i + V.N * (long long)j, row_bytes, bar);
            }
        }
        return true;
    }
    for (int s = 0; s < V.n_streams; ++s) {
        const Stream sm = V.streams[s];
        if (sm.kind == JBUGS_ARG_DATA_I) {
            for (int q = tid; q < cnt; q += BLK) cp_async8(dst + sm.soff + q, sm.ptr + i + q);
        } else if (sm.kind == JBUGS_ARG_DATA_IJ) {
            for (int q = tid; q < T * STAGE_G; q += BLK) {
                const int j = q / STAGE_G, g = q % STAGE_G;
                if (g < cnt) cp_async8(dst + sm.soff + q, sm.ptr + i + g + V.N * (long long)j);
            }
        }
    }
    return false;
}

// wait until the staged chunk is visible: TMA chunks complete on the mbarrier (bounded spin: a protocol error must
// surface as a flagged chain, never as a hung GPU), cp.async chunks on the group counter; then the CTA barrier.
template <bool TMA>
__device__ __forceinline__ bool stage_wait(bool was_bulk, unsigned long long* bar, unsigned parity) {
    bool ok = true;
    if (TMA && was_bulk) {
        ok = false;
#pragma unroll 1  // keep the kernel inside the 32 KB instruction cache: the unrolled spin added 600 instructions
        for (int it = 0; it < (1 << 22); ++it)
            if (mbar_try_wait(bar, parity)) { ok = true; break; }
    }
    cp_async_wait_all();
    __syncthreads();
    return ok;
}

template <class S>
__device__ __forceinline__ long long local_slot(const S& sp, const PlateVals& V, int l, long long i, long long j) {
    const LocalV& L = V.loc[l];
    if (sp.loc_has_idx(l)) return sp.loc_level(l) == JBUGS_LEVEL_OBS ? __ldg(L.idx + i + V.N * j) : __ldg(L.idx + i);
    return L.base + L.si * i + L.sj * j;
}

// constrained value + logjac of local l from its raw theta slot value
template <class S>
__device__ __forceinline__ void local_enter(const S& sp, const PlateVals& V, int l, double raw, ChainState& st,
                                            double (&ldx)[MAX_LOC], double (&ldlj)[MAX_LOC], double& lp) {
    // -0.0, the additive identity for EVERY double (x + 0.0 turns -0.0 into +0.0, so the compiler must keep that add;
    // x + -0.0 is x, and it folds): the first contribution to an accumulator then costs no FP64 instruction
    st.la[l] = -0.0;
    if (sp.loc_transform(l) == JBUGS_TR_IDENTITY) {
        // (written out: `lp += 0.0`, `la * 1.0 + 0.0` are not foldable under IEEE rules and would each cost an FP64 slot)
        st.lx[l] = raw;
        ldx[l] = 1.0;
        ldlj[l] = 0.0;
        return;
    }
    const TrOut t = transform_eval(sp.loc_transform(l), V.loc[l].lb, V.loc[l].ub, raw);
    st.lx[l] = t.x;
    ldx[l] = t.dxdy;
    ldlj[l] = t.dljdy;
    lp += t.lj;
}

// gradient of the slot of local l: (d logp / d x) * dx/dy + d logjac / dy
template <class S>
__device__ __forceinline__ double local_grad(const S& sp, int l, const ChainState& st, const double (&ldx)[MAX_LOC],
                                             const double (&ldlj)[MAX_LOC]) {
    return sp.loc_transform(l) == JBUGS_TR_IDENTITY ? st.la[l] : fma(st.la[l], ldx[l], ldlj[l]);
}

template <class S>
__device__ __forceinline__ bool local_prior(const S& sp, const PlateVals& V, int l, const StageView& sv, int g, int j,
                                            ChainState& st, FusedAcc& fa, double& lp) {
    const LocalV& L = V.loc[l];
    const double a0 = arg_value<S::SOFF, S::NG_MAX>(sp.loc_arg(l, 0), L.args[0], st, sv, g, j);
    const double a1 = arg_value<S::SOFF, S::NG_MAX>(sp.loc_arg(l, 1), L.args[1], st, sv, g, j);
    if (sp.prior_fused(l)) {
        // dnorm(mu, tau) with chain-invariant arguments: keep sufficient statistics, expand at tile end.  The adjoint
        // -tau (x - mu) is applied where the gradient is stored (eval_groups: one fma with the likelihood's part).
        if (sp.prior_mu_zero(l)) {  // mu is the literal 0 (random effects): sum x^2 is all that is needed
            fa.pss[l] = fma(st.lx[l], st.lx[l], fa.pss[l]);
            fa.pd[l] = st.lx[l];
            return false;
        }
        const double d = st.lx[l] - a0;
        fa.pss[l] = fma(d, d, fa.pss[l]);
        if (sp.loc_arg(l, 0).kind == JBUGS_ARG_GVAL) fa.ps[l] += d;  // only the adjoint of a parameter mean needs it
        fa.pd[l] = d;
        return false;
    }
    if (sp.loc_family(l) == JBUGS_FAM_GAMMA && arg_invariant(sp.loc_arg(l, 0)) && arg_invariant(sp.loc_arg(l, 1))) {
        // dgamma(a, b), chain-invariant arguments: the argument-only terms were hoisted (plate_kernel prologue)
        const double x = st.lx[l], lx = log(x);
        lp += x < 0.0 ? -dev_inf() : (fma(a0, st.pre_lb[l], -st.pre_lg[l]) + (a0 == 1.0 ? 0.0 : (a0 - 1.0) * lx) - a1 * x);
        st.la[l] += (a0 - 1.0) / x - a1;
        arg_adjoint<S::NG_MAX>(sp.loc_arg(l, 0), st, st.pre_lb[l] - st.pre_dg[l] + lx);
        arg_adjoint<S::NG_MAX>(sp.loc_arg(l, 1), st, a0 / a1 - x);
        return !(a0 > 0.0) || !(a1 > 0.0);
    }
    FamOut o;
    const bool bad = fam_eval(sp.loc_family(l), st.lx[l], a0, a1, o, L.args[2].value);
    lp += o.lp;
    st.la[l] += o.dx;
    arg_adjoint<S::NG_MAX>(sp.loc_arg(l, 0), st, o.da0);
    arg_adjoint<S::NG_MAX>(sp.loc_arg(l, 1), st, o.da1);
    return bad;
}

// linear predictor of observation (g, j): eta = sum_t coef_t * cov_t (+ offset); cv receives the covariate values
template <class S>
__device__ __forceinline__ double predictor(const S& sp, const PlateVals& V, const StageView& sv, int g, int j,
                                            const ChainState& st, double (&cv)[MAX_TERMS]) {
    double eta = -0.0;  // (additive identity that folds, see local_enter)
#pragma unroll
    for (int t = 0; t < S::KG_MAX; ++t)
        if (t < sp.kg()) {
            cv[t] = arg_value<S::SOFF, S::NG_MAX>(sp.cov(t), V.cov[t], st, sv, g, j);
            eta = fma(st.gv[t], cv[t], eta);
        }
#pragma unroll
    for (int l = 0; l < S::NL_MAX; ++l)
        if (l < sp.nl()) {
            cv[MAX_GT + l] = arg_value<S::SOFF, S::NG_MAX>(sp.cov(sp.kg() + l), V.cov[sp.kg() + l], st, sv, g, j);
            eta = fma(st.lx[l], cv[MAX_GT + l], eta);
        }
    if (sp.has_off()) eta += arg_value<S::SOFF, S::NG_MAX>(sp.cov(sp.kg() + sp.nl()), V.cov[sp.kg() + sp.nl()], st, sv, g, j);
    return eta;
}

template <int NG>
__device__ __forceinline__ double tape_leaf(const XOpDev& o, const StageView& sv, int g, int j, const ChainState& st) {
    switch (o.kind) {
    case JBUGS_ARG_CONST: return o.value;
    case JBUGS_ARG_ONE: return 1.0;
    case JBUGS_ARG_GVAL: return pick_n<NG>(st.gv, o.id);
    case JBUGS_ARG_GVEC_J: return pick_n<NG>(st.gv, o.id + j);
    case JBUGS_ARG_LOCAL: return pick(st.lx, o.id);
    case JBUGS_ARG_DATA_I: return sv.chunk[o.soff + g];
    case JBUGS_ARG_DATA_J: return sv.jreg[o.soff + j];
    case JBUGS_ARG_DATA_IJ: return sv.chunk[o.soff + j * STAGE_G + g];
    default: return dev_nan();
    }
}

// forward sweep of an expression tape for one observation; `state` is the state of the discrete latent being summed out
// (PICK_K ops), 0 when there is none.  d1 / d2 receive the partial derivatives of each op for the reverse sweep.
template <class S>
__device__ __noinline__ bool tape_forward(const S& sp, const PlateVals& V, const StageView& sv, int g, int j,
                                          const ChainState& st, int state, double* v, double* d1, double* d2) {
    const int nx = sp.n_x();
    bool bad = false;
    for (int k = 0; k < nx; ++k) {
        const XOpDev o = V.xtape[k];
        d1[k] = 0.0; d2[k] = 0.0;
        if (o.op == JBUGS_OP_LEAF) {
            v[k] = tape_leaf<S::NG_MAX>(o, sv, g, j, st);
        } else if (o.op == JBUGS_OP_SELECT) {
            v[k] = v[o.a] != 0.0 ? v[o.b] : v[o.c];
        } else if (o.op == JBUGS_OP_PICK_K) {
            v[k] = v[o.a + state];
        } else if (o.op == JBUGS_OP_GVEC_AT) {
            v[k] = pick_n<S::NG_MAX>(st.gv, o.id + (int)v[o.a]);  // (an index outside the vector selects nothing: 0)
        } else if (o.op < 16) {
            bad |= unary_op(o.op, v[o.a], v[k], d1[k]);
        } else {
            const double a = v[o.a], b = v[o.b];
            switch (o.op) {
            case JBUGS_OP_ADD: v[k] = a + b; d1[k] = 1.0; d2[k] = 1.0; break;
            case JBUGS_OP_SUB: v[k] = a - b; d1[k] = 1.0; d2[k] = -1.0; break;
            case JBUGS_OP_MUL: v[k] = a * b; d1[k] = b; d2[k] = a; break;
            case JBUGS_OP_DIV: v[k] = a / b; d1[k] = 1.0 / b; d2[k] = -a / (b * b); break;
            default:  // POW
                v[k] = pow(a, b); d1[k] = b * pow(a, b - 1.0); d2[k] = a > 0.0 ? v[k] * log(a) : 0.0;
                bad |= a < 0.0 && b != floor(b);
                break;
            }
        }
    }
    return bad;
}

// the family of an expression plate at the tape's values; seeds the adjoints w[] of the tape slots its arguments sit in
// (scaled by r: the posterior weight of the latent's state, 1 without a latent) when `seed` is set
template <class S>
__device__ __noinline__ bool tape_family(const S& sp, const PlateVals& V, const StageView& sv, int g, int j, ChainState& st,
                                         const double* v, double* w, double yv, double r, bool seed, double& lp_out) {
    double a[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const ArgS s = sp.aux(q);
        a[q] = s.kind == JBUGS_ARG_XVAL ? v[s.id] : arg_value<false, S::NG_MAX>(s, V.aux[q], st, sv, g, j);
    }
    FamOut o;
    const bool bad = fam_eval(sp.family(), yv, a[0], a[1], o, V.aux[2].value);
    lp_out = o.lp;
    if (seed) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const ArgS s = sp.aux(q);
            const double da = r * (q == 0 ? o.da0 : o.da1);
            if (s.kind == JBUGS_ARG_XVAL) w[s.id] += da;
            else arg_adjoint<S::NG_MAX>(s, st, da);
        }
    }
    return bad;
}

// reverse sweep: adjoints w[] back to the leaves (global / local adjoint accumulators of the chain)
template <class S>
__device__ __noinline__ void tape_reverse(const S& sp, const PlateVals& V, int j, ChainState& st, int state, const double* v,
                                          double* w, const double* d1, const double* d2) {
    for (int k = sp.n_x() - 1; k >= 0; --k) {
        const double wk = w[k];
        if (wk == 0.0) continue;
        const XOpDev o = V.xtape[k];
        if (o.op == JBUGS_OP_LEAF) {
            if (o.kind == JBUGS_ARG_GVAL) bump_n<S::NG_MAX>(st.ga, o.id, wk);
            else if (o.kind == JBUGS_ARG_GVEC_J) bump_n<S::NG_MAX>(st.ga, o.id + j, wk);
            else if (o.kind == JBUGS_ARG_LOCAL) bump(st.la, o.id, wk);
        } else if (o.op == JBUGS_OP_SELECT) {
            w[v[o.a] != 0.0 ? o.b : o.c] += wk;
        } else if (o.op == JBUGS_OP_PICK_K) {
            w[o.a + state] += wk;
        } else if (o.op == JBUGS_OP_GVEC_AT) {
            bump_n<S::NG_MAX>(st.ga, o.id + (int)v[o.a], wk);
        } else {
            w[o.a] += wk * d1[k];
            if (o.op >= 16) w[o.b] += wk * d2[k];
        }
    }
}

// one observation of an expression plate: forward sweep over the tape, the family, reverse sweep back to the leaves
// (the composed node functions of the observation's logical parents, compiler_pass.jl:751-814, in SSA form).
// With a discrete latent that is summed out (n_states > 0; evaluation.jl:739-961 for one latent per plate cell): the
// cell contributes logsumexp_k [log p(z = k) + log p(y | z = k)] -- a first round of forward sweeps for the K branch
// values, a second one whose reverse sweeps are weighted by the posterior probability of each state.
template <class S>
__device__ __noinline__ bool observe_tape(const S& sp, const PlateVals& V, const StageView& sv, int g, int j,
                                          ChainState& st, double& lp) {
    double v[JBUGS_MAX_XOPS], w[JBUGS_MAX_XOPS], d1[JBUGS_MAX_XOPS], d2[JBUGS_MAX_XOPS];
    const int nx = sp.n_x();
    const int K = sp.n_states();
    const double yv = arg_value<false, S::NG_MAX>(sp.y(), V.y, st, sv, g, j);
    bool bad = false;
    double l1;
    if (K <= 0) {
        bad |= tape_forward(sp, V, sv, g, j, st, 0, v, d1, d2);
        for (int k = 0; k < nx; ++k) w[k] = 0.0;
        bad |= tape_family(sp, V, sv, g, j, st, v, w, yv, 1.0, true, l1);
        lp += l1;
        tape_reverse(sp, V, j, st, 0, v, w, d1, d2);
        return bad;
    }
    double branch[JBUGS_MAX_STATES];
    const int lps = sp.lp_slot();
    // a partially observed latent: an observed cell (state > 0 in the data leaf) takes the branch of its value only
    const int zo = sp.zobs_slot() >= 0 ? (int)tape_leaf<S::NG_MAX>(V.xtape[sp.zobs_slot()], sv, g, j, st) : 0;
    double mx = -dev_inf();
    for (int k = 0; k < K; ++k) {
        if (zo > 0 && k != zo - 1) { branch[k] = -dev_inf(); continue; }
        bad |= tape_forward(sp, V, sv, g, j, st, k, v, d1, d2);
        bad |= tape_family(sp, V, sv, g, j, st, v, w, yv, 0.0, false, l1);
        branch[k] = l1 + v[lps];
        mx = branch[k] > mx ? branch[k] : mx;
    }
    if (!(mx > -dev_inf())) {  // no state has support (or NaN): nothing to differentiate
        lp += mx;
        return bad;
    }
    double sum = 0.0;
    for (int k = 0; k < K; ++k) sum += exp(branch[k] - mx);
    const double lse = mx + log(sum);
    lp += lse;
    for (int k = 0; k < K; ++k) {
        const double r = exp(branch[k] - lse);
        if (r == 0.0) continue;
        tape_forward(sp, V, sv, g, j, st, k, v, d1, d2);
        for (int q = 0; q < nx; ++q) w[q] = 0.0;
        w[lps] = r;
        tape_family(sp, V, sv, g, j, st, v, w, yv, r, true, l1);
        tape_reverse(sp, V, j, st, k, v, w, d1, d2);
    }
    return bad;
}

// one observation: linear predictor, inverse link, family; accumulates adjoints
template <class S>
__device__ __forceinline__ bool observe(const S& sp, const PlateVals& V, const StageView& sv, int g, int j,
                                        ChainState& st, FusedAcc& fa, double (&sl)[MAX_LOC], double& lp) {
    double cv[MAX_TERMS];
    bool bad = false;
    if (sp.has_w() && arg_value<S::SOFF, S::NG_MAX>(sp.aux(2), V.aux[2], st, sv, g, j) == 0.0) return false;  // padded cell: no observation
    if constexpr (S::XTAPE) return observe_tape(sp, V, sv, g, j, st, lp);
    if constexpr (S::XGEN) return S::D::observe_gen(sp, V, sv, g, j, st, lp);
    const double eta = predictor(sp, V, sv, g, j, st, cv);
    const double yv = arg_value<S::SOFF, S::NG_MAX>(sp.y(), V.y, st, sv, g, j);

    if constexpr (S::FUSED == FUSED_NORMAL_IDENTITY) {
        // y ~ dnorm(eta, tau), tau chain-invariant: d/d eta = tau * r; tau is factored out of every sum
        const double r = yv - eta;
        fa.ss = fma(r, r, fa.ss);
#pragma unroll
        for (int t = 0; t < S::KG_MAX; ++t)
            if (t < sp.kg()) fa.sg[t] = fma(r, cv[t], fa.sg[t]);
#pragma unroll
        for (int l = 0; l < S::NL_MAX; ++l)
            if (l < sp.nl()) sl[l] = fma(r, cv[MAX_GT + l], sl[l]);
        return false;
    }

    double deta;
    if constexpr (S::FUSED == FUSED_BINOMIAL_LOGIT || S::FUSED == FUSED_BERNOULLI_LOGIT) {
        // k ~ dbin(logistic(eta), n): branch-free, selects and range tests on the integer pipe (binomial_logit);
        // the saturated tails are fixed up once per unrolled block of groups (eval_groups), not per observation
        const double n = S::FUSED == FUSED_BERNOULLI_LOGIT ? 1.0 : arg_value<S::SOFF, S::NG_MAX>(sp.aux(1), V.aux[1], st, sv, g, j);
        binomial_logit(eta, yv, n, lp, deta, st.sat);
    } else if constexpr (S::FUSED == FUSED_POISSON_LOG) {
        // k ~ dpois(exp(eta)): k * eta - exp(eta); -lgamma(k+1) goes to obs_const
        const double lam = exp(eta);
        // xlogy(k, lambda) saturates to -Inf once exp underflows to 0 (reference semantics)
        lp += (yv == 0.0 ? 0.0 : (lam == 0.0 ? -dev_inf() : yv * eta)) - lam;
        deta = yv - lam;
    } else if constexpr (S::FUSED == FUSED_BINOMIAL_LINK || S::FUSED == FUSED_BERNOULLI_LINK || S::FUSED == FUSED_POISSON_LINK) {
        // any inverse-link tape; only the theta-dependent part of the log-density (the rest is in obs_const)
        double v = eta, dv = 1.0;
#pragma unroll
        for (int q = 0; q < JBUGS_MAX_LINK; ++q)
            if (q < sp.n_link()) {
                double nv, d;
                bad |= unary_op(sp.link(q), v, nv, d);
                v = nv;
                dv *= d;
            }
        if constexpr (S::FUSED == FUSED_POISSON_LINK) {
            bad |= !(v >= 0.0);
            lp += xlogy(yv, v) - v;
            deta = ((yv == 0.0 ? 0.0 : yv / v) - 1.0) * dv;
        } else {
            const double n = S::FUSED == FUSED_BERNOULLI_LINK ? 1.0 : arg_value<S::SOFF, S::NG_MAX>(sp.aux(1), V.aux[1], st, sv, g, j);
            const double m = n - yv;
            bad |= !(v >= 0.0 && v <= 1.0);
            lp += xlogy(yv, v) + xlog1py(m, -v);
            deta = ((yv == 0.0 ? 0.0 : yv / v) - (m == 0.0 ? 0.0 : m / (1.0 - v))) * dv;
        }
    } else {
        double v = eta, dv = 1.0;
#pragma unroll
        for (int q = 0; q < JBUGS_MAX_LINK; ++q)
            if (q < sp.n_link()) {
                double nv, d;
                bad |= unary_op(sp.link(q), v, nv, d);
                v = nv;
                dv *= d;
            }
        const int ea = sp.eta_arg();
        const double a0 = ea == 0 ? v : arg_value<S::SOFF, S::NG_MAX>(sp.aux(0), V.aux[0], st, sv, g, j);
        const double a1 = ea == 1 ? v : arg_value<S::SOFF, S::NG_MAX>(sp.aux(1), V.aux[1], st, sv, g, j);
        FamOut o;
        bad |= fam_eval(sp.family(), yv, a0, a1, o, V.aux[2].value);
        lp += o.lp;
        deta = (ea == 0 ? o.da0 : o.da1) * dv;
        if (ea != 0) arg_adjoint<S::NG_MAX>(sp.aux(0), st, o.da0);
        if (ea != 1) arg_adjoint<S::NG_MAX>(sp.aux(1), st, o.da1);
    }
#pragma unroll
    for (int t = 0; t < S::KG_MAX; ++t)
        if (t < sp.kg()) st.ga[t] = fma(deta, cv[t], st.ga[t]);
#pragma unroll
    for (int l = 0; l < S::NL_MAX; ++l)
        if (l < sp.nl()) st.la[l] = fma(deta, cv[MAX_GT + l], st.la[l]);
    return bad;
}

// precision of a fused dnorm prior (chain-invariant: a referenced global value or a literal)
template <class S>
__device__ __forceinline__ double prior_tau(const S& sp, const PlateVals& V, int l, const ChainState& st) {
    const ArgS a1 = sp.loc_arg(l, 1);
    return a1.kind == JBUGS_ARG_GVAL ? pick_n<S::NG_MAX>(st.gv, a1.id) : (a1.kind == JBUGS_ARG_ONE ? 1.0 : V.loc[l].args[1].value);
}

// raw theta values (and row offsets) of the group-level locals of UU consecutive groups; when the inner extent is
// a small compile-time constant (TS > 0) the observation-level locals of those groups as well, so that no theta
// load sits right in front of its first use
template <int UU, int TS = 0>
struct GroupLoads {
    double raw[UU][MAX_LOC];
    long long off[UU][MAX_LOC];  // element offset of the slot's row: slot * ld
    double oraw[UU][MAX_LOC][TS > 0 ? TS : 1];
    long long obase[UU][MAX_LOC];  // offset of (i, j = 0); j advances by step_j
};

// observation-level locals are preloaded with their group when the inner extent is static and small
template <class S>
struct PreObs {
    static constexpr int TS = (S::OBS_LOCALS && S::T_STATIC > 0 && S::T_STATIC <= 8) ? S::T_STATIC : 0;
};

// issue the theta loads of groups i .. i+UU-1.  `pos[l]` walks the affine slot map incrementally
// (row0 + step_i * i, in elements), so calls must come in ascending group order.
template <class S, int UU>
__device__ __forceinline__ void load_groups(const S& sp, const PlateVals& V, long long i, const double* __restrict__ th,
                                            long long ld, long long (&pos)[MAX_LOC],
                                            GroupLoads<UU, PreObs<S>::TS>& gl) {
    constexpr int TS = PreObs<S>::TS;
#pragma unroll
    for (int u = 0; u < UU; ++u)
#pragma unroll
        for (int l = 0; l < S::NL_MAX; ++l) {
            if (l < sp.nl() && sp.loc_level(l) == JBUGS_LEVEL_GROUP) {
                if (sp.loc_has_idx(l)) {
                    gl.off[u][l] = local_slot(sp, V, l, i + u, 0) * ld;
                } else {
                    gl.off[u][l] = pos[l];
                    pos[l] += V.loc[l].step_i;
                }
                gl.raw[u][l] = th[gl.off[u][l]];
                // FP64-bound plates: pull the row S::L2PF groups ahead into L2 (no register cost; the later load finds it
                // there).  The theta load right in front of its use was the top stall of the logit kernels (long
                // scoreboard, profiles/r2_seeds_plate_v5_summary.txt); distance 8 measured best (4..32 tried).
                if (S::L2PF > 0 && !sp.loc_has_idx(l))
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(th + gl.off[u][l] + (long long)S::L2PF * V.loc[l].step_i));
            }
            if (TS > 0 && l < sp.nl() && sp.loc_level(l) == JBUGS_LEVEL_OBS && !sp.loc_has_idx(l)) {
                gl.obase[u][l] = pos[l];
                pos[l] += V.loc[l].step_i;
#pragma unroll
                for (int j = 0; j < TS; ++j) gl.oraw[u][l][j] = th[gl.obase[u][l] + j * V.loc[l].step_j];
            }
        }
}

// SMPF: the same for a block whose theta rows were copied to shared memory while the previous block was evaluated
// (`src`, this thread's column of a [UU][MAX_LOC] table of BLK-wide rows; null: load from global memory), and the
// cp.async copies of the block after it (`dst`; null: none).  pos[] ends up at the start of the next block either way.
template <class S, int UU, int BLK>
__device__ __forceinline__ void load_groups_sm(const S& sp, const PlateVals& V, const double* __restrict__ th,
                                               long long (&pos)[MAX_LOC], GroupLoads<UU, PreObs<S>::TS>& gl,
                                               const double* src, double* dst, bool newer_group) {
    // this thread's own copies (it is the only reader of its column).  `newer_group`: the staging copies of the next
    // chunk were committed after them and may stay in flight
    if (src) {
        if (newer_group) asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        else cp_async_wait_all();
    }
#pragma unroll
    for (int u = 0; u < UU; ++u)
#pragma unroll
        for (int l = 0; l < S::NL_MAX; ++l)
            if (l < sp.nl()) {
                gl.off[u][l] = pos[l];
                pos[l] += V.loc[l].step_i;
                gl.raw[u][l] = src ? src[(u * MAX_LOC + l) * BLK] : th[gl.off[u][l]];
            }
    if (dst) {
#pragma unroll
        for (int u = 0; u < UU; ++u)
#pragma unroll
            for (int l = 0; l < S::NL_MAX; ++l)
                if (l < sp.nl()) cp_async8(dst + (u * MAX_LOC + l) * BLK, th + pos[l] + u * V.loc[l].step_i);
        cp_async_commit();
    }
}

// UU consecutive groups starting at chunk-local index g (global group index i), theta already loaded
template <class S, int UU>
__device__ __forceinline__ bool eval_groups(const S& sp, const PlateVals& V, const StageView& sv, int g, long long i,
                                            int T, const double* __restrict__ th, double* __restrict__ gr,
                                            long long ld, bool store, double obs_tau, long long (&pos)[MAX_LOC],
                                            const GroupLoads<UU, PreObs<S>::TS>& gl, ChainState& st, FusedAcc& fa,
                                            double& lp) {
    constexpr int TS = PreObs<S>::TS;
    bool bad = false;
    const double (&raw)[UU][MAX_LOC] = gl.raw;
    const long long (&off)[UU][MAX_LOC] = gl.off;
#pragma unroll
    for (int u = 0; u < UU; ++u) {
        double ldx[MAX_LOC], ldlj[MAX_LOC], sl[MAX_LOC];
#pragma unroll
        for (int l = 0; l < S::NL_MAX; ++l) {
            sl[l] = -0.0;
            if (l < sp.nl() && sp.loc_level(l) == JBUGS_LEVEL_GROUP) local_enter(sp, V, l, raw[u][l], st, ldx, ldlj, lp);
        }
#pragma unroll
        for (int l = 0; l < S::NL_MAX; ++l)
            if (l < sp.nl() && sp.loc_level(l) == JBUGS_LEVEL_GROUP) bad |= local_prior(sp, V, l, sv, g + u, 0, st, fa, lp);

        long long opos[MAX_LOC];
#pragma unroll
        for (int l = 0; l < S::NL_MAX; ++l) opos[l] = (TS > 0 && !sp.loc_has_idx(l)) ? gl.obase[u][l] : pos[l];
#pragma unroll(S::T_STATIC > 0 ? S::T_STATIC : 1)
        for (int j = 0; j < T; ++j) {
            long long ooff[MAX_LOC];
#pragma unroll
            for (int l = 0; l < S::NL_MAX; ++l)
                if (l < sp.nl() && sp.loc_level(l) == JBUGS_LEVEL_OBS) {
                    if (sp.loc_has_idx(l)) {
                        ooff[l] = local_slot(sp, V, l, i + u, j) * ld;
                    } else {
                        ooff[l] = opos[l];
                        opos[l] += V.loc[l].step_j;
                    }
                    const double rawv = (TS > 0 && !sp.loc_has_idx(l)) ? gl.oraw[u][l][j < TS ? j : 0] : th[ooff[l]];
                    local_enter(sp, V, l, rawv, st, ldx, ldlj, lp);
                }
#pragma unroll
            for (int l = 0; l < S::NL_MAX; ++l)
                if (l < sp.nl() && sp.loc_level(l) == JBUGS_LEVEL_OBS) bad |= local_prior(sp, V, l, sv, g + u, j, st, fa, lp);
            if (sp.has_obs()) bad |= observe(sp, V, sv, g + u, j, st, fa, sl, lp);
#pragma unroll
            for (int l = 0; l < S::NL_MAX; ++l)
                if (l < sp.nl() && sp.loc_level(l) == JBUGS_LEVEL_OBS) {
                    if constexpr (S::FUSED == FUSED_NORMAL_IDENTITY) {
                        st.la[l] = fma(obs_tau, sl[l], st.la[l]);
                        sl[l] = -0.0;
                    }
                    if (sp.prior_fused(l)) st.la[l] = fma(-prior_tau(sp, V, l, st), fa.pd[l], st.la[l]);
                    if (store) gr[ooff[l]] = local_grad(sp, l, st, ldx, ldlj);
                }
        }
#pragma unroll
        for (int l = 0; l < S::NL_MAX; ++l) {
            if (l < sp.nl() && sp.loc_level(l) == JBUGS_LEVEL_OBS && !(TS > 0 && !sp.loc_has_idx(l)))
                pos[l] += V.loc[l].step_i;
            if (l < sp.nl() && sp.loc_level(l) == JBUGS_LEVEL_GROUP) {
                if (S::FUSED == FUSED_NORMAL_IDENTITY) st.la[l] = fma(obs_tau, sl[l], st.la[l]);
                if (sp.prior_fused(l)) st.la[l] = fma(-prior_tau(sp, V, l, st), fa.pd[l], st.la[l]);
                if (store) gr[off[u][l]] = local_grad(sp, l, st, ldx, ldlj);
            }
        }
    }
    if constexpr (S::FUSED == FUSED_BINOMIAL_LOGIT || S::FUSED == FUSED_BERNOULLI_LOGIT) if (__builtin_expect(st.sat != 0, 0)) {
        // rare: some |eta| of this block lies where the reference's logistic is exactly 0 or 1.  Walk the block again,
        // recompute each linear predictor (same arithmetic, same bits) and give the observation the reference's -Inf
        // where it applies; the finite value the hot path added is swallowed by it.
        st.sat = 0;
#pragma unroll
        for (int u = 0; u < UU; ++u) {
            double d0[MAX_LOC], d1[MAX_LOC], dummy = 0.0;
#pragma unroll
            for (int l = 0; l < S::NL_MAX; ++l)
                if (l < sp.nl() && sp.loc_level(l) == JBUGS_LEVEL_GROUP) local_enter(sp, V, l, raw[u][l], st, d0, d1, dummy);
            long long opos[MAX_LOC];
#pragma unroll
            for (int l = 0; l < S::NL_MAX; ++l)  // pos[] has moved past the block: group u sits (UU - u) steps back
                opos[l] = (TS > 0 && !sp.loc_has_idx(l)) ? gl.obase[u][l] : pos[l] - (long long)(UU - u) * V.loc[l].step_i;
#pragma unroll 1
            for (int j = 0; j < T; ++j) {
#pragma unroll
                for (int l = 0; l < S::NL_MAX; ++l)
                    if (l < sp.nl() && sp.loc_level(l) == JBUGS_LEVEL_OBS) {
                        const long long o = sp.loc_has_idx(l) ? local_slot(sp, V, l, i + u, j) * ld : opos[l] + j * V.loc[l].step_j;
                        local_enter(sp, V, l, th[o], st, d0, d1, dummy);
                    }
                if (sp.has_w() && arg_value<S::SOFF, S::NG_MAX>(sp.aux(2), V.aux[2], st, sv, g + u, j) == 0.0) continue;
                double cv[MAX_TERMS];
                const double eta = predictor(sp, V, sv, g + u, j, st, cv);
                const double yv = arg_value<S::SOFF, S::NG_MAX>(sp.y(), V.y, st, sv, g + u, j);
                const double n = S::FUSED == FUSED_BERNOULLI_LOGIT ? 1.0 : arg_value<S::SOFF, S::NG_MAX>(sp.aux(1), V.aux[1], st, sv, g + u, j);
                if (logit_tail_is_inf(eta, yv, n)) lp = -dev_inf();
            }
        }
    }
    return bad;
}

// BLK = chains per CTA (128, or 64 / 32 for small batches so that no lane idles); TMA = stage the data streams with
// cp.async.bulk + mbarrier where aligned (a separate instantiation: the cp.async-only kernel carries none of it)
// GRAD = false: a forward-only instantiation for `logdensity` without the gradient (the adjoint arithmetic is dead
// code then and the kernel reads theta only); built for the large-plate specs, selected when want_grad == 0
__device__ __forceinline__ void copy_chain_state(ChainState& d, const ChainState& s) {
#pragma unroll
    for (int k = 0; k < MAX_GREF; ++k) { d.gv[k] = s.gv[k]; d.ga[k] = s.ga[k]; }
#pragma unroll
    for (int l = 0; l < MAX_LOC; ++l) {
        d.lx[l] = s.lx[l]; d.la[l] = s.la[l]; d.pre_lg[l] = s.pre_lg[l]; d.pre_dg[l] = s.pre_dg[l]; d.pre_lb[l] = s.pre_lb[l];
    }
    d.sat = s.sat;
}
__device__ __forceinline__ void copy_fused_acc(FusedAcc& d, const FusedAcc& s) {
    d.ss = s.ss;
#pragma unroll
    for (int t = 0; t < MAX_GT; ++t) d.sg[t] = s.sg[t];
#pragma unroll
    for (int l = 0; l < MAX_LOC; ++l) { d.pss[l] = s.pss[l]; d.ps[l] = s.ps[l]; d.pd[l] = s.pd[l]; }
}

// one group of a short last block (see plate_body): everything the group loop carries, by value
struct TailState {
    ChainState st;
    FusedAcc fa;
    double lp;
    long long pos[MAX_LOC];
    bool bad;
};
template <class S>
__device__ __noinline__ TailState tail_group(const PlateParams& P, StageView sv, int g, long long i, int T, const double* __restrict__ th,
                                             double* __restrict__ gr, long long ld, bool store, double obs_tau, TailState ts) {
    const S sp(P.sh);
    GroupLoads<1, PreObs<S>::TS> one;
    load_groups<S, 1>(sp, P.v, i, th, ld, ts.pos, one);
    ts.bad |= eval_groups<S, 1>(sp, P.v, sv, g, i, T, th, gr, ld, store, obs_tau, ts.pos, one, ts.st, ts.fa, ts.lp);
    return ts;
}

// the body of plate_kernel: one CTA = BLK chains x tile `tile` of the plate's groups (also the middle part of
// fused_small_kernel, jbugs_kernels.cuh)
template <class S, int BLK, bool TMA, bool GRAD>
__device__ __forceinline__ void plate_body(const PlateParams& P, const double* __restrict__ theta, double* __restrict__ grad,
                                           long long ld, long long C, const double* __restrict__ gvals, long long ldg,
                                           double* __restrict__ partial, int* __restrict__ flags, int want_grad,
                                           const long long tile) {
    extern __shared__ double smem[];
    pdl_launch_dependents();  // finalize (or the next plate) may be scheduled; it waits for this grid before reading
    const S sp(P.sh);
    const PlateVals& V = P.v;
    const int tid = threadIdx.x;
    const long long c_raw = (long long)blockIdx.x * BLK + tid;
    const bool active = c_raw < C;           // inactive lanes still take part in staging and barriers
    const long long c = active ? c_raw : C - 1;
    const long long ib = V.i0 + tile * V.tile;
    const long long ie = (ib + V.tile < V.i1) ? ib + V.tile : V.i1;
    constexpr int U = S::UNROLL;
    const int T = S::T_STATIC > 0 ? S::T_STATIC : (int)V.T;

    // shared memory: J region, then two stages of [per_group][STAGE_G]
    double* jreg = smem;
    double* stage0 = smem + V.j_doubles;
    const int stage_doubles = V.per_group * STAGE_G;
    for (int s = 0; s < V.n_streams; ++s) {
        const Stream sm = V.streams[s];
        if (sm.kind == JBUGS_ARG_DATA_J)
            for (int q = tid; q < T; q += BLK) jreg[sm.soff + q] = __ldg(sm.ptr + q);
    }
    // one mbarrier per stage for the TMA path (transaction-count completion)
    __shared__ unsigned long long mbar[2];
    if (TMA) {
        if (tid == 0) {
            mbar_init(&mbar[0], 1);
            mbar_init(&mbar[1], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    bool next_bulk = false;     // the chunk in flight went through the TMA path
    unsigned parity_bits = 0u;  // bit s: phase parity of mbar[s]
    bool stage_ok = true;
    {
        const long long rem = ie - ib;
        next_bulk = stage_chunk<BLK, TMA>(V, stage0, ib, rem < STAGE_G ? (int)rem : STAGE_G, T, tid, &mbar[0]);
        cp_async_commit();
    }

    pdl_wait();  // everything above (parameters, first data chunk in flight) may overlap the prologue kernel
    ChainState st;
#pragma unroll
    for (int k = 0; k < S::NG_MAX; ++k) {
        st.gv[k] = (k < sp.n_gref()) ? gvals[(long long)V.gref[k] * ldg + c] : 0.0;
        st.ga[k] = 0.0;
    }
#pragma unroll
    for (int l = 0; l < S::NL_MAX; ++l)
        if (l < sp.nl() && sp.loc_family(l) == JBUGS_FAM_GAMMA && arg_invariant(sp.loc_arg(l, 0)) &&
            arg_invariant(sp.loc_arg(l, 1))) {
            const ArgS s0 = sp.loc_arg(l, 0), s1 = sp.loc_arg(l, 1);
            const double a0 = s0.kind == JBUGS_ARG_GVAL ? pick_n<S::NG_MAX>(st.gv, s0.id) : (s0.kind == JBUGS_ARG_ONE ? 1.0 : V.loc[l].args[0].value);
            const double a1 = s1.kind == JBUGS_ARG_GVAL ? pick_n<S::NG_MAX>(st.gv, s1.id) : (s1.kind == JBUGS_ARG_ONE ? 1.0 : V.loc[l].args[1].value);
            st.pre_lg[l] = lgamma(a0);
            st.pre_dg[l] = digamma(a0);
            st.pre_lb[l] = log(a1);
        }
    FusedAcc fa;
    fa.ss = 0.0;
#pragma unroll
    for (int t = 0; t < S::KG_MAX; ++t) fa.sg[t] = 0.0;
#pragma unroll
    for (int l = 0; l < S::NL_MAX; ++l) fa.pss[l] = fa.ps[l] = fa.pd[l] = 0.0;
    st.sat = 0;

    double lp = 0.0;
    bool bad = false;
    // precision of a fused normal observation (chain-invariant by construction of the spec)
    const double obs_tau = (S::FUSED == FUSED_NORMAL_IDENTITY)
                               ? (sp.aux(1).kind == JBUGS_ARG_GVAL ? pick_n<S::NG_MAX>(st.gv, sp.aux(1).id) : V.aux[1].value)
                               : 0.0;
    const double* __restrict__ th = theta + c;
    double* __restrict__ gr = grad + c;
    const bool store = GRAD && want_grad && active;

    long long pos[MAX_LOC];
#pragma unroll
    for (int l = 0; l < S::NL_MAX; ++l) pos[l] = V.loc[l].row0 + V.loc[l].step_i * ib;

    stage_ok &= stage_wait<TMA>(next_bulk, &mbar[0], parity_bits & 1u);
    if (next_bulk) parity_bits ^= 1u;
    next_bulk = false;

    int cur = 0;
    GroupLoads<U, PreObs<S>::TS> gl_cur;
    bool have_cur = false;
    int tail_g = 0, tail_cnt = 0, tail_stage = 0;  // (run-time compiled kernels: remainder groups of the last chunk)
    long long tail_i0 = 0;
    // SMPF: two [U][MAX_LOC][BLK] tables behind the staging buffers
    double* const pf_base = stage0 + 2 * stage_doubles;
    int pf_cur = 0;
    for (long long i0 = ib; i0 < ie; i0 += STAGE_G) {
        const long long rem = ie - i0;
        const int cnt = rem < STAGE_G ? (int)rem : STAGE_G;
        if (i0 + STAGE_G < ie) {  // prefetch the next chunk into the other stage
            const long long rem2 = ie - (i0 + STAGE_G);
            next_bulk = stage_chunk<BLK, TMA>(V, stage0 + (cur ^ 1) * stage_doubles, i0 + STAGE_G,
                                                   rem2 < STAGE_G ? (int)rem2 : STAGE_G, T, tid, &mbar[cur ^ 1]);
        }
        cp_async_commit();
        StageView sv;
        sv.chunk = stage0 + cur * stage_doubles;
        sv.jreg = jreg;
        int g = 0;
        for (; g + U <= cnt; g += U) {
            const long long i = i0 + g;
            if constexpr (S::PREFETCH || !GRAD) {  // forward-only: no store traffic to fill the memory pipe, so keep twice the loads in flight
                // software pipeline over U-blocks (also across staged chunks: the theta rows do not depend on the
                // staged data): the loads of block i+U are in flight while block i is evaluated
                if (!have_cur) load_groups<S, U>(sp, V, i, th, ld, pos, gl_cur);
                GroupLoads<U, PreObs<S>::TS> gl_nxt;
                const bool more = i + 2 * U <= ie;
                if (more) load_groups<S, U>(sp, V, i + U, th, ld, pos, gl_nxt);
                bad |= eval_groups<S, U>(sp, V, sv, g, i, T, th, gr, ld, store, obs_tau, pos, gl_cur, st, fa, lp);
                if (more) gl_cur = gl_nxt;
                have_cur = more;
            } else if constexpr (S::SMPF) {
                // theta of block i came through shared memory if the block before it asked for it; ask for block i + U
                const bool more = i + 2 * U <= ie;
                double* const mine = pf_base + tid;
                load_groups_sm<S, U, BLK>(sp, V, th, pos, gl_cur, have_cur ? mine + pf_cur * (U * MAX_LOC * BLK) : nullptr,
                                          more ? mine + (pf_cur ^ 1) * (U * MAX_LOC * BLK) : nullptr, g == 0 && i0 + STAGE_G < ie);
                pf_cur ^= 1;
                have_cur = more;
                bad |= eval_groups<S, U>(sp, V, sv, g, i, T, th, gr, ld, store, obs_tau, pos, gl_cur, st, fa, lp);
            } else {
                load_groups<S, U>(sp, V, i, th, ld, pos, gl_cur);
                bad |= eval_groups<S, U>(sp, V, sv, g, i, T, th, gr, ld, store, obs_tau, pos, gl_cur, st, fa, lp);
            }
        }
#if !defined(__CUDACC_RTC__) || defined(JB_INLINE_TAIL)
        if (U > 1)
            for (; g < cnt; ++g) {
                GroupLoads<1, PreObs<S>::TS> one;
                load_groups<S, 1>(sp, V, i0 + g, th, ld, pos, one);
                bad |= eval_groups<S, 1>(sp, V, sv, g, i0 + g, T, th, gr, ld, store, obs_tau, pos, one, st, fa, lp);
            }
#else
        // run-time compiled kernels (NVRTC): the groups of a short last block (only the LAST chunk of a tile can have one)
        // are evaluated after the chunk loop, see below
        if (U > 1 && g < cnt) { tail_g = g; tail_cnt = cnt; tail_i0 = i0; tail_stage = cur; }
#endif
        // the next chunk (issued at the top of this iteration) must have landed before it is read; the barrier also
        // protects the stage just consumed from the refill issued at the top of the next iteration
        stage_ok &= stage_wait<TMA>(next_bulk, &mbar[cur ^ 1], (parity_bits >> (cur ^ 1)) & 1u);
        if (next_bulk) parity_bits ^= 1u << (cur ^ 1);
        next_bulk = false;
        cur ^= 1;
    }
    bad |= !stage_ok;
#if defined(__CUDACC_RTC__) && !defined(JB_INLINE_TAIL)
    if (U > 1 && tail_cnt > 0) {
        // They go one at a time through a NON-inlined copy of the body, state by value, OUTSIDE the loop nest so that the
        // hot loop's registers stay registers.  A second inlined copy of the body cost NVVM 9 of the 12 s a kernel took to
        // compile; the in-tree kernels keep it (their ptxas budgets are pinned by a test, nobody waits for their
        // compilation).  The last chunk's staging buffer is still intact: nothing was staged after it.
        StageView sv;
        sv.chunk = stage0 + tail_stage * stage_doubles;
        sv.jreg = jreg;
        TailState ts;
        copy_chain_state(ts.st, st);
        copy_fused_acc(ts.fa, fa);
        ts.lp = lp; ts.bad = bad;
#pragma unroll
        for (int l = 0; l < MAX_LOC; ++l) ts.pos[l] = pos[l];
#pragma unroll 1
        for (int g = tail_g; g < tail_cnt; ++g) ts = tail_group<S>(P, sv, g, tail_i0 + g, T, th, gr, ld, store, obs_tau, ts);
        copy_chain_state(st, ts.st);
        copy_fused_acc(fa, ts.fa);
        lp = ts.lp; bad = ts.bad;
    }
#endif

    // expand the sufficient statistics of the fused paths
    const double n_obs = (double)((ie - ib) * V.T);
    if constexpr (S::FUSED == FUSED_NORMAL_IDENTITY) {
        const bool tg = sp.aux(1).kind == JBUGS_ARG_GVAL;
        const double tau = obs_tau;
        bad |= tau < 0.0;
        lp += n_obs * (0.5 * log(tau) - 0.5 * kLog2Pi) - 0.5 * tau * fa.ss;
        if (tg) bump_n<S::NG_MAX>(st.ga, sp.aux(1).id, n_obs * 0.5 / tau - 0.5 * fa.ss);
#pragma unroll
        for (int t = 0; t < S::KG_MAX; ++t)
            if (t < sp.kg()) st.ga[t] = fma(tau, fa.sg[t], st.ga[t]);
    }
#pragma unroll
    for (int l = 0; l < S::NL_MAX; ++l)
        if (l < sp.nl() && sp.prior_fused(l)) {
            const ArgS a0 = sp.loc_arg(l, 0), a1 = sp.loc_arg(l, 1);
            const double tau = prior_tau(sp, V, l, st);
            // instances of local l in this tile (a padded cell of a gathered plate still holds a parameter)
            const double cnt = (double)((ie - ib) * (sp.loc_level(l) == JBUGS_LEVEL_OBS ? V.T : 1));
            bad |= tau < 0.0;
            lp += cnt * (0.5 * log(tau) - 0.5 * kLog2Pi) - 0.5 * tau * fa.pss[l];
            if (a0.kind == JBUGS_ARG_GVAL) bump_n<S::NG_MAX>(st.ga, a0.id, tau * fa.ps[l]);
            if (a1.kind == JBUGS_ARG_GVAL) bump_n<S::NG_MAX>(st.ga, a1.id, cnt * 0.5 / tau - 0.5 * fa.pss[l]);
        }

    if (!active) return;
    // per (tile, chain) partials: [logp, adjoints of the referenced global values]
    double* out = partial + tile * (long long)(1 + sp.n_gref()) * ldg + c;
    out[0] = lp;
    if (GRAD)
#pragma unroll
        for (int k = 0; k < S::NG_MAX; ++k)
            if (k < sp.n_gref()) out[(long long)(1 + k) * ldg] = st.ga[k];
    if (bad) atomicOr(flags + c, 1);
}

template <class S, int BLK, bool TMA, bool GRAD = true>
__global__ void __launch_bounds__(BLK, (S::MIN_BLOCKS * (128 / BLK) > 24 ? 24 : S::MIN_BLOCKS * (128 / BLK)))
plate_kernel(const __grid_constant__ PlateParams P, const double* __restrict__ theta, double* __restrict__ grad,
             long long ld, long long C, const double* __restrict__ gvals, long long ldg,
             double* __restrict__ partial, int* __restrict__ flags, int want_grad) {
    plate_body<S, BLK, TMA, GRAD>(P, theta, grad, ld, C, gvals, ldg, partial, flags, want_grad, blockIdx.y);
}

}  // namespace jbugs
