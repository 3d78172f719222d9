# lower_to_plan.jl -- the plate-level lowering pass, in Julia, on top of the reference's own source-generation passes.
#
# What a JuliaBUGS maintainer adds next to `src/source_gen.jl`.  SHIPPED UNEXECUTED (no Julia in the build image); it is
# the second implementation of the pass whose first implementation is `jbugs/lowering.py`, and it is checked byte for
# byte against the plan blobs committed under `tests/golden/plans/` by `julia/check_golden_plans.jl` on any machine that
# has Julia and JuliaBUGS.
#
# Division of labour with the reference (paths relative to JuliaBUGS/):
#   * statement ids, coarse statement graph, loop fission, topological order of statements, regrouping of consecutive
#     statements with identical loop nests -> `_generate_lowered_model_def` (src/source_gen.jl:459-552).  Its second
#     return value, the *reconstructed* model definition, is the input here: a sequentially valid program of `for`
#     loops, `lhs ~ dist(...)` and `lhs = expr` statements, transformed data already removed.
#   * observed / parameter / generated-quantity classes (src/source_gen.jl:615-666, src/graphs.jl:27-71) ->
#     `model.graph_evaluation_data` (src/model/bugsmodel.jl:118-246), read per variable.
#   * data and transformed data -> `model.evaluation_env` (src/JuliaBUGS.jl:175-187).
# What this file adds is everything after that: theta slots of the reconstructed program, scalar globals + gtape, and per
# observed statement one plate record -- linear predictor term tape + inverse-link tape, or an expression tape when the
# predictor is not linear in the parameters -- plus the data slots.  Section letters (F, G ...) and function names follow
# `jbugs/lowering.py` so that the two can be read side by side.
module JBugsLowering

export lower_to_plan, LoweringError, Plan, to_bytes

struct LoweringError <: Exception
    msg::String
end
Base.showerror(io::IO, e::LoweringError) = print(io, "LoweringError: ", e.msg)
fail(msg) = throw(LoweringError(msg))

# ------------------------------------------------------------------------------------------------ plan records
# (include/jbugs_plan.h; numeric codes must match the header)
const ARG_CONST, ARG_GVAL, ARG_LOCAL, ARG_DATA_I, ARG_DATA_J, ARG_DATA_IJ, ARG_ONE, ARG_XVAL, ARG_GVEC_J = UInt32.(0:8)
const FAMILY_OF = Dict(:dnorm => 0, :dgamma => 1, :dexp => 2, :dunif => 3, :dbeta => 4, :dlnorm => 5, :dbern => 6,
                       :dbin => 7, :dpois => 8, :dflat => 9, :dlogis => 10, :ddexp => 11, :dweib => 12, :dnegbin => 13,
                       :dchisqr => 14, :dt => 15, :dgeom => 16, :dcat => 17, :dbetabin => 18)
const FAMILY_NARGS = Dict(0 => 2, 1 => 2, 2 => 1, 3 => 2, 4 => 2, 5 => 2, 6 => 1, 7 => 2, 8 => 1, 9 => 0, 10 => 2, 11 => 2,
                          12 => 2, 13 => 2, 14 => 1, 15 => 3, 16 => 1, 17 => 1, 18 => 3)
const FAM_GAMMA, FAM_EXPONENTIAL, FAM_UNIFORM, FAM_BETA, FAM_LOGNORMAL = 1, 2, 3, 4, 5
const FAM_BERNOULLI, FAM_BINOMIAL, FAM_FLAT, FAM_WEIBULL, FAM_CHISQ, FAM_T, FAM_CATEGORICAL = 6, 7, 9, 12, 14, 15, 17
const FAM_NORMAL, FAM_LOGISTIC, FAM_LAPLACE = 0, 10, 11
const FAM_BETABIN = 18   # dbetabin(a, b, n): the third argument (n) is a constant, like dt's degrees of freedom
const DISCRETE = Set([6, 7, 8, 13, 16, 17, 18])
const TR_IDENTITY, TR_LOG, TR_LOGIT, TR_UPPER = 0, 1, 2, 3
const OP_IDENTITY, OP_EXP, OP_LOGISTIC, OP_CEXPEXP, OP_PHI, OP_LOG, OP_SQRT, OP_NEG = 0:7
const OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_LEAF, OP_SELECT, OP_PICK_K = 16, 17, 18, 19, 20, 32, 33, 34
const OP_GVEC_AT = 35   # gval[leaf.id + Int(v[a])] (jbugs_plan.h): LeukFr's b[pair[i]] inside an expression plate
const LINK_OF = Dict(:exp => OP_EXP, :logistic => OP_LOGISTIC, :ilogit => OP_LOGISTIC, :cexpexp => OP_CEXPEXP,
                     :icloglog => OP_CEXPEXP, :phi => OP_PHI)
const UNARY_OF = merge(LINK_OF, Dict(:log => OP_LOG, :sqrt => OP_SQRT))
const LEVEL_GROUP, LEVEL_OBS = 0, 1
const MAX_LINK, MAX_XOPS, MAX_STATES = 4, 64, 16

struct Arg
    kind::UInt32
    id::Int32
    value::Float64
end
Arg(kind, id=0) = Arg(UInt32(kind), Int32(id), 0.0)
const_arg(v) = Arg(ARG_CONST, Int32(0), Float64(v))
one_arg() = Arg(ARG_ONE, Int32(0), 1.0)
const ZERO_ARG = const_arg(0.0)

mutable struct Global
    name::String; theta_idx::Int; transform::Int; family::Int; args::Vector{Arg}; lb::Float64; ub::Float64
end
struct GOp
    op::Int; a::Arg; b::Arg
end
mutable struct DataSlot
    name::String; values::Union{Vector{Float64},Vector{Int64}}; dim0::Int; rank::Int
end
struct Local
    name::String; level::Int; transform::Int; theta_base::Int; stride_i::Int; stride_j::Int; family::Int
    args::Vector{Arg}; lb::Float64; ub::Float64; idx_slot::Int
end
struct Term
    coef::Arg; cov::Arg
end
struct XOp
    op::Int; a::Int; b::Int; c::Int; leaf::Arg
end
XOp(op, a=0, b=0, c=0) = XOp(op, a, b, c, ZERO_ARG)
mutable struct Plate
    N::Int; T::Int; locals::Vector{Local}; terms::Vector{Term}; link::Vector{Int}; family::Int; eta_arg::Int
    has_obs::Bool; y::Arg; aux::Vector{Arg}; name::String; n_real::Int; weight::Union{Nothing,Arg}; xops::Vector{XOp}
    n_states::Int        # > 0: one finite discrete latent per observation is summed out (has_obs == 3)
    logprior_slot::Int   # tape slot of log p(z = k) (+ the log densities of the latent's further observed children)
    zobs_slot::Int       # tape slot of a data leaf with the latent's observed state + 1 per cell (0 = missing); -1: none
end
Plate(N, T; family=0, eta_arg=0, has_obs=true, name="") =
    Plate(N, T, Local[], Term[], Int[], family, eta_arg, has_obs, ZERO_ARG, Arg[], name, -1, nothing, XOp[], 0, -1, -1)
struct Dense
    N::Int; P::Int; theta_base::Int; theta_stride::Int; x_slot::Int; y_slot::Int; n_slot::Int; family::Int; link::Int
end
mutable struct Plan
    D::Int; transformed::Bool; globals::Vector{Global}; gtape::Vector{GOp}; data::Vector{DataSlot}
    plates::Vector{Plate}; dense::Vector{Dense}
end

# ---- serialisation: little-endian packed records in header order (jbugs_plan.h:17-26) ------------------------------
wr(io, a::Arg) = (write(io, htol(a.kind)); write(io, htol(a.id)); write(io, htol(a.value)))
function wr3(io, args::Vector{Arg})
    for q in 1:3
        wr(io, q <= length(args) ? args[q] : ZERO_ARG)
    end
end
function to_bytes(p::Plan)::Vector{UInt8}
    io = IOBuffer()
    nl = sum(length(pl.locals) for pl in p.plates; init=0)
    nt = sum(length(pl.terms) for pl in p.plates; init=0)
    nx = sum(length(pl.xops) for pl in p.plates; init=0)
    write(io, b"JBUGSPL1"); write(io, htol(UInt32(1))); write(io, htol(UInt32(p.transformed ? 1 : 0))); write(io, htol(Int64(p.D)))
    for n in (length(p.globals), length(p.gtape), length(p.data), length(p.plates), nl, nt, length(p.dense), nx)
        write(io, htol(UInt32(n)))
    end
    write(io, htol(UInt64(0)))
    for g in p.globals
        write(io, htol(Int64(g.theta_idx))); write(io, htol(UInt32(g.transform))); write(io, htol(UInt32(g.family)))
        write(io, htol(g.lb)); write(io, htol(g.ub)); wr3(io, g.args)
    end
    for o in p.gtape
        write(io, htol(UInt32(o.op))); write(io, htol(UInt32(0))); wr(io, o.a); wr(io, o.b)
    end
    for d in p.data
        n = length(d.values)
        write(io, htol(Int64(n))); write(io, htol(Int64(d.dim0 == 0 ? n : d.dim0))); write(io, htol(UInt32(d.rank)))
        write(io, htol(UInt32(d.values isa Vector{Int64} ? 1 : 0)))
    end
    il = it = ix = 0
    for pl in p.plates
        length(pl.link) <= MAX_LINK || fail("link tape longer than JBUGS_MAX_LINK")
        is_x = !isempty(pl.xops)
        write(io, htol(Int64(pl.N))); write(io, htol(Int64(pl.T)))
        write(io, htol(UInt32(il))); write(io, htol(UInt32(length(pl.locals))))
        write(io, htol(UInt32(is_x ? ix : it))); write(io, htol(UInt32(is_x ? length(pl.xops) : length(pl.terms))))
        write(io, htol(UInt32(length(pl.link))))
        for q in 1:MAX_LINK
            v = q <= length(pl.link) ? pl.link[q] : OP_IDENTITY
            if pl.n_states > 0       # has_obs == 3 (plan.py to_bytes): link[0] = log-prior slot, link[1] = observed-state slot + 1
                q == 1 && (v = pl.logprior_slot)
                q == 2 && (v = pl.zobs_slot + 1)
            end
            write(io, htol(UInt32(v)))
        end
        write(io, htol(UInt32(pl.family))); write(io, htol(UInt32(pl.n_states > 0 ? pl.n_states : pl.eta_arg)))
        write(io, htol(UInt32(pl.has_obs ? (is_x ? (pl.n_states > 0 ? 3 : 2) : 1) : 0)))
        wr(io, pl.y)
        aux = pl.weight === nothing ? pl.aux : vcat((vcat(pl.aux, [ZERO_ARG, ZERO_ARG]))[1:2], [pl.weight])
        wr3(io, aux)
        il += length(pl.locals); it += length(pl.terms); ix += length(pl.xops)
    end
    for pl in p.plates, l in pl.locals
        write(io, htol(UInt32(l.level))); write(io, htol(UInt32(l.transform))); write(io, htol(Int64(l.theta_base)))
        write(io, htol(Int64(l.stride_i))); write(io, htol(Int64(l.stride_j))); write(io, htol(Int32(l.idx_slot)))
        write(io, htol(UInt32(l.family))); write(io, htol(l.lb)); write(io, htol(l.ub)); wr3(io, l.args)
    end
    for pl in p.plates, t in pl.terms
        wr(io, t.coef); wr(io, t.cov)
    end
    for d in p.dense
        for v in (d.N, d.P, d.theta_base, d.theta_stride)
            write(io, htol(Int64(v)))
        end
        for v in (d.x_slot, d.y_slot, d.n_slot)
            write(io, htol(Int32(v)))
        end
        write(io, htol(UInt32(d.family))); write(io, htol(UInt32(d.link))); write(io, htol(UInt32(0)))
    end
    for pl in p.plates, o in pl.xops
        write(io, htol(UInt32(o.op))); write(io, htol(UInt16(o.a))); write(io, htol(UInt16(o.b))); write(io, htol(UInt16(o.c)))
        write(io, htol(UInt16(0))); write(io, htol(UInt32(0))); wr(io, o.leaf)
    end
    return take!(io)
end

# ------------------------------------------------------------------------------------------------ statements
mutable struct Stmt
    sid::Int                                  # 1-based id in the order of the reconstructed program
    stoch::Bool                               # lhs ~ dist(...)  /  lhs = rhs
    lhs::Any                                  # Symbol | Expr(:ref, name, idx...)
    rhs::Any                                  # Expr(:call, dist, args...) | expression
    loops::Vector{Tuple{Symbol,Any,Any}}      # (var, lo, hi) outermost first
    extents::Vector{Tuple{Int,Int}}
    kind::Symbol                              # :obs | :param | :logical
    is_gq::Bool
    ragged::Bool
    obs_mask::Union{Nothing,Array{Bool}}
    zobs::Union{Nothing,Array{Float64}}       # a discrete latent being summed out, partially observed: state number + 1 per cell, 0 = missing
    trunc::Any                                # (lower, upper) of `truncated(dist, lower, upper)` (BUGS T(,)) | nothing
end
name_of(s::Stmt) = s.lhs isa Symbol ? s.lhs : s.lhs.args[1]
lhs_idx(s::Stmt) = s.lhs isa Symbol ? Any[] : s.lhs.args[2:end]
loop_vars(s::Stmt) = Symbol[l[1] for l in s.loops]
shape_of(s::Stmt) = Int[hi - lo + 1 for (lo, hi) in s.extents]
size_of(s::Stmt) = prod(shape_of(s); init=1)

# Julia parses a + b + c as one n-ary call and unary minus as a one-argument call: bring every arithmetic call to the
# binary / unary forms the rest of the pass assumes (the Python AST has BinOp / Neg)
function binarize(e)
    e isa Expr || return e
    if e.head == :call && e.args[1] in (:+, :*) && length(e.args) > 3
        acc = binarize(e.args[2])
        for a in e.args[3:end]
            acc = Expr(:call, e.args[1], acc, binarize(a))
        end
        return acc
    end
    return Expr(e.head, map(binarize, e.args)...)
end
is_call(e, f) = e isa Expr && e.head == :call && e.args[1] == f
is_neg(e) = e isa Expr && e.head == :call && e.args[1] == :- && length(e.args) == 2
is_binop(e) = e isa Expr && e.head == :call && e.args[1] in (:+, :-, :*, :/, :^) && length(e.args) == 3
is_ref(e) = e isa Expr && e.head == :ref
is_range(e) = e isa Expr && e.head == :call && e.args[1] == :(:) && length(e.args) == 3
is_colon(e) = e === :(:) || e == Colon()

function flatten!(out::Vector{Stmt}, block, loops)
    for st in block.args
        st isa LineNumberNode && continue
        if st isa Expr && st.head == :for
            hdr = st.args[1]
            v, rng = hdr.args[1], hdr.args[2]
            is_range(rng) || fail("loop over something that is not a range: $(rng)")
            flatten!(out, st.args[2], vcat(loops, [(v, binarize(rng.args[2]), binarize(rng.args[3]))]))
        elseif st isa Expr && st.head == :block
            flatten!(out, st, loops)
        elseif st isa Expr && st.head == :call && st.args[1] == :~
            # (censored(dist, lower, upper): jbugs/lowering.py checks the observed values against the bounds and keeps
            #  the inner distribution; not ported to this file yet -- refuse rather than mis-evaluate)
            (st.args[3] isa Expr && st.args[3].head == :call && st.args[3].args[1] == :censored) &&
                fail("censored statements are lowered by the Python front end only")
            if st.args[3] isa Expr && st.args[3].head == :call && st.args[3].args[1] == :truncated
                # truncated(dist(...), lower, upper): the statement keeps the inner distribution, the bounds become the
                # support of a scalar parameter's prior (truncation, below)
                length(st.args[3].args) == 4 || fail("malformed truncated(...)")
                inner, lo, hi = st.args[3].args[2:4]
                push!(out, Stmt(length(out) + 1, true, st.args[2], binarize(inner), loops, Tuple{Int,Int}[], :none, false, false,
                                nothing, nothing, (binarize(lo), binarize(hi))))
                continue
            end
            push!(out, Stmt(length(out) + 1, true, st.args[2], binarize(st.args[3]), loops, Tuple{Int,Int}[], :none, false, false, nothing, nothing, nothing))
        elseif st isa Expr && st.head == :(=)
            push!(out, Stmt(length(out) + 1, false, st.args[1], binarize(st.args[2]), loops, Tuple{Int,Int}[], :none, false, false, nothing, nothing, nothing))
        else
            fail("unsupported statement $(st)")
        end
    end
    return out
end

function free_vars!(acc::Vector{Symbol}, e)
    if e isa Symbol
        is_colon(e) || e in acc || push!(acc, e)
    elseif e isa Expr
        start = e.head == :call ? 2 : 1        # the callee of a call is a function name, not a variable
        for a in e.args[start:end]
            free_vars!(acc, a)
        end
    end
    return acc
end
free_vars(e) = free_vars!(Symbol[], e)

function subst(e, sub::Dict{Symbol,Any})
    e isa Symbol && return get(sub, e, e)
    e isa Expr || return e
    start = e.head == :call ? 2 : 1
    return Expr(e.head, e.args[1:(start - 1)]..., map(a -> subst(a, sub), e.args[start:end])...)
end

# ------------------------------------------------------------------------------------------------ data evaluation
# (lowering.py DataEval: a data-only expression at one point of a loop domain)
mutable struct DataEval
    data::Dict{Symbol,Any}
    lenient::Bool          # clamp out-of-range / missing subscripts (case-split leaves, lowering.py:139-146)
    unknown::Set{Symbol}   # arrays that are partly data, partly unknown (a partially observed discrete latent): not data
end
DataEval(data, lenient) = DataEval(data, lenient, Set{Symbol}())
known(de::DataEval, e, lvs=Symbol[]) = all(v -> (haskey(de.data, v) && !(v in de.unknown)) || v in lvs, free_vars(e))
logistic(x) = 1 / (1 + exp(-x))
function ev(de::DataEval, e, lv::Dict{Symbol,Int})
    e isa Number && return Float64(e)
    if e isa Symbol
        haskey(lv, e) && return Float64(lv[e])
        return de.data[e]
    end
    if is_ref(e)
        arr = de.data[e.args[1]]
        idx = Any[]
        for (d, ie) in enumerate(e.args[2:end])
            if is_colon(ie)
                push!(idx, Colon())
            elseif is_range(ie)
                push!(idx, Int(ev(de, ie.args[2], lv)):Int(ev(de, ie.args[3], lv)))
            else
                v = ev(de, ie, lv)
                if de.lenient
                    v = clamp(isnan(v) ? 1.0 : floor(v), 1, size(arr, d))
                end
                v == floor(v) || fail("non-integer array index")
                push!(idx, Int(v))
            end
        end
        r = arr[idx...]
        return r isa Number ? (ismissing(r) ? NaN : Float64(r)) : r
    end
    if is_neg(e)
        return -ev(de, e.args[2], lv)
    end
    if e isa Expr && e.head == :call
        f = e.args[1]
        a = [ev(de, x, lv) for x in e.args[2:end]]
        f == :+ && return a[1] + a[2]
        f == :- && return a[1] - a[2]
        f == :* && return a[1] * a[2]
        f == :/ && return a[1] / a[2]
        (f == :^ || f == :pow) && return a[1]^a[2]
        f == :exp && return exp(a[1]); f == :log && return log(a[1]); f == :sqrt && return sqrt(a[1]); f == :abs && return abs(a[1])
        (f == :logistic || f == :ilogit) && return logistic(a[1])
        f == :logit && return log(a[1] / (1 - a[1]))
        (f == :cexpexp || f == :icloglog) && return -expm1(-exp(a[1]))
        f == :cloglog && return log(-log1p(-a[1]))
        (f == :step || f == :_step) && return a[1] >= 0 ? 1.0 : 0.0
        f == :mean && return sum(a[1]) / length(a[1])
        f == :sum && return Float64(sum(a[1]))
        f == :sd && return sqrt(sum((a[1] .- sum(a[1]) / length(a[1])) .^ 2) / (length(a[1]) - 1))
        f == :inprod && return Float64(sum(vec(a[1]) .* vec(a[2])))
        f == :max && return max(a[1], a[2]); f == :min && return min(a[1], a[2])
        (f == :equals || f == :__eq__) && return a[1] == a[2] ? 1.0 : 0.0
        f == :__le__ && return a[1] <= a[2] ? 1.0 : 0.0
        f == :__and__ && return (a[1] != 0 && a[2] != 0) ? 1.0 : 0.0
        f == :phi && return 0.5 * erfc_(-a[1] / sqrt(2.0))
        f == :loggam && return lgamma_(a[1])
        f == :logfact && return lgamma_(a[1] + 1.0)
        fail("function $(f) is not supported in data expressions")
    end
    fail("cannot evaluate $(e)")
end
erfc_(x) = ccall(:erfc, Float64, (Float64,), x)     # libm; avoids a SpecialFunctions dependency in this file
lgamma_(x) = ccall(:lgamma, Float64, (Float64,), x)

"values of `e` over the loop domain `extents` (column-major Julia array, one axis per loop variable)"
function grid(de::DataEval, e, lvs::Vector{Symbol}, extents)
    shp = Tuple(hi - lo + 1 for (lo, hi) in extents)
    out = Array{Float64}(undef, shp...)
    for I in CartesianIndices(shp)
        lv = Dict{Symbol,Int}(lvs[d] => extents[d][1] + I[d] - 1 for d in 1:length(lvs))
        v = ev(de, e, lv)
        out[I] = ismissing(v) ? NaN : Float64(v)
    end
    return out
end

# ------------------------------------------------------------------------------------------------ the pass
mutable struct Lowering
    stmts::Vector{Stmt}
    de::DataEval
    transformed::Bool
    node_info::Dict{String,Tuple{Bool,Bool,Bool}}     # "alpha[3]" -> (is_stochastic, is_observed, is_generated_quantity)
    theta_map::Dict{Int,Tuple{Int,Vector{Int}}}        # sid -> (0-based base slot, stride per loop)
    D::Int
    plan::Plan
    slot_ids::Dict{String,Int}
    gval_ids::Dict{String,Int}
    cur_extents::Vector{Tuple{Int,Int}}
    marginalize::Bool                                  # UseAutoMarginalization: finite discrete latents are summed out, not in theta
    latent::Any                                        # (statement of the latent summed out in the plate being built, K) | nothing
    latent_sibs::Dict{Int,Vector{Stmt}}                # sid of a plate's observation -> further observed children of its latent
end

label(name, idx) = isempty(idx) ? string(name) : string(name, "[", join(idx, ","), "]")

# ---- A. loop bounds (lowering.py _bounds) ---------------------------------------------------------------------------
function bounds!(lw::Lowering)
    for s in lw.stmts
        outer = Symbol[]
        for (v, lo, hi) in s.loops
            if !(known(lw.de, lo) && known(lw.de, hi))
                if known(lw.de, lo, outer) && known(lw.de, hi, outer) && !s.stoch
                    s.ragged = true
                    break
                end
                fail("loop bounds of '$(v)' must depend on data only (rectangular plates)")
            end
            push!(s.extents, (Int(ev(lw.de, lo, Dict{Symbol,Int}())), Int(ev(lw.de, hi, Dict{Symbol,Int}()))))
            push!(outer, v)
        end
        s.ragged && empty!(s.extents)
    end
end

# ---- D. classification from the reference's graph data (replaces lowering.py _classify) ------------------------------
function classify!(lw::Lowering)
    for s in lw.stmts
        if !s.stoch
            s.kind = :logical
        else
            s.ragged && fail("stochastic statement '$(name_of(s))' in a loop with ragged bounds")
            obs = grid_flags(lw, s)
            fn = s.rhs.args[1]
            if lw.marginalize && !all(obs) && (fn == :dcat || fn == :dbern)
                # an unobserved finite discrete node: summed out (lowering.py _classify).  Partially observed
                # (auto_marginalization.jl:391-462): the observed cells take the branch of their value, recorded as
                # state number + 1 (a dcat value k is state k - 1, a dbern value v is state v)
                s.kind = :latent
                if any(obs)
                    vals = grid(lw.de, s.lhs, loop_vars(s), s.extents)
                    s.zobs = [obs[k] ? vals[k] + (fn == :dbern ? 1.0 : 0.0) : 0.0 for k in eachindex(vals)]
                    s.zobs = reshape(s.zobs, size(vals))
                    push!(lw.de.unknown, name_of(s))
                end
            elseif lw.marginalize && !all(obs) && fn == :dbin
                fail("'$(name_of(s))': only dcat and dbern latents are summed out")
            elseif all(obs)
                s.kind = :obs
            elseif !any(obs)
                s.kind = :param
            else
                s.kind = :obs          # partially observed: the missing cells must be generated quantities (checked below)
                s.obs_mask = obs
            end
        end
    end
    referenced = Set{Symbol}()
    for t in lw.stmts
        union!(referenced, free_vars(t.rhs))
    end
    for s in lw.stmts
        s.obs_mask !== nothing && name_of(s) in referenced &&
            fail("'$(name_of(s))' is partially observed and has children (mixed observation/parameter statements are outside the plate class)")
    end
    # generated quantities: a statement all of whose instances the reference classified as GeneratedQuantity
    for s in lw.stmts
        (s.kind == :param || s.kind == :logical || s.kind == :latent) || continue
        s.ragged && continue
        s.is_gq = all(instances(lw, s)) do lab
            info = get(lw.node_info, lab, nothing)
            info !== nothing && info[3]
        end
    end
end
function instances(lw::Lowering, s::Stmt)
    lvs = loop_vars(s)
    out = String[]
    for I in CartesianIndices(Tuple(shape_of(s)))
        lv = Dict{Symbol,Int}(lvs[d] => s.extents[d][1] + I[d] - 1 for d in 1:length(lvs))
        push!(out, label(name_of(s), [Int(ev(lw.de, ie, lv)) for ie in lhs_idx(s)]))
    end
    return out
end
function grid_flags(lw::Lowering, s::Stmt)
    flags = Array{Bool}(undef, Tuple(shape_of(s))...)
    for (k, lab) in enumerate(instances(lw, s))
        info = get(lw.node_info, lab, nothing)
        info === nothing && fail("variable $(lab) is not a node of the model graph")
        flags[k] = info[2]
    end
    return flags
end

# ---- E. theta slots: program order of the reconstructed program (bugsmodel.jl:770-799; lowering.py:441-460) -------------
function order_parameters!(lw::Lowering)
    groups = Vector{Vector{Stmt}}()
    for s in lw.stmts
        if !isempty(groups) && groups[end][1].loops == s.loops
            push!(groups[end], s)
        else
            push!(groups, [s])
        end
    end
    base = 0
    for body in groups
        params = [s for s in body if s.kind == :param && !s.is_gq]
        m = length(params)
        m == 0 && continue
        shp = shape_of(body[1])
        strides = Int[]
        acc = m
        for d in length(shp):-1:1
            pushfirst!(strides, acc)
            acc *= shp[d]
        end
        for (q, s) in enumerate(params)
            lw.theta_map[s.sid] = (base + q - 1, strides)
        end
        base += m * size_of(body[1])
    end
    lw.D = base
end
"0-based theta slot of every instance of parameter statement `s` (array over its loop domain)"
function theta_index(lw::Lowering, s::Stmt)
    base, strides = lw.theta_map[s.sid]
    shp = Tuple(shape_of(s))
    out = Array{Int64}(undef, shp...)
    for I in CartesianIndices(shp)
        out[I] = base + sum(strides[d] * (I[d] - 1) for d in 1:length(shp); init=0)
    end
    return out
end
function param_names(lw::Lowering)
    names = Vector{String}(undef, lw.D)
    for s in lw.stmts
        haskey(lw.theta_map, s.sid) || continue
        for (lab, t) in zip(instances(lw, s), vec(theta_index(lw, s)))
            names[t + 1] = lab
        end
    end
    return names
end

# ---- F/G. plan construction ---------------------------------------------------------------------------------------------
function slot!(lw::Lowering, key::String, values; rank=1, dim0=0)
    haskey(lw.slot_ids, key) && return lw.slot_ids[key]
    vals = values isa AbstractArray{<:Integer} ? collect(Int64, vec(values)) : collect(Float64, vec(values))
    push!(lw.plan.data, DataSlot(key, vals, dim0, rank))
    lw.slot_ids[key] = length(lw.plan.data) - 1
    return lw.slot_ids[key]
end
defs(lw::Lowering, name) = [s for s in lw.stmts if name_of(s) == name]
is_data(lw::Lowering, e, lvs) = known(lw.de, e, lvs)

"'name[3]' for a reference with constant (data-only, loop-free) subscripts, else nothing"
function elem_label(lw::Lowering, e)
    is_ref(e) || return nothing
    idx = e.args[2:end]
    any(i -> is_range(i) || is_colon(i), idx) && return nothing
    all(i -> is_data(lw, i, Symbol[]), idx) || return nothing
    return label(e.args[1], [Int(ev(lw.de, i, Dict{Symbol,Int}())) for i in idx])
end

"substitute references to logical variables by their right-hand sides (compiler_pass.jl:751-814 composed along the parent chain)"
function inline(lw::Lowering, e, depth=0)
    depth > 32 && fail("logical nodes nest too deeply")
    (e isa Number || is_colon(e)) && return e
    if e isa Symbol
        haskey(lw.gval_ids, string(e)) && return e
        ds = [s for s in defs(lw, e) if s.kind == :logical]
        if !isempty(ds)
            (length(ds) == 1 && isempty(ds[1].loops)) || fail("cannot inline '$(e)'")
            return inline(lw, ds[1].rhs, depth + 1)
        end
        return e
    end
    if is_ref(e)
        idx = Any[inline(lw, i, depth + 1) for i in e.args[2:end]]
        me = Expr(:ref, e.args[1], idx...)
        lab = elem_label(lw, me)
        lab !== nothing && haskey(lw.gval_ids, lab) && return me
        ds = [s for s in defs(lw, e.args[1]) if s.kind == :logical]
        isempty(ds) && return me
        d1 = ds[1]
        plain = length(ds) == 1 && !d1.ragged && !(d1.lhs isa Symbol) && all(li -> li isa Symbol && li in loop_vars(d1), lhs_idx(d1))
        if plain
            sub = Dict{Symbol,Any}(li => ri for (li, ri) in zip(lhs_idx(d1), idx))
            return inline(lw, subst(d1.rhs, sub), depth + 1)
        end
        # case split over several defining statements / constant lhs subscripts (lowering.py _inline, Bones)
        out = Expr(:call, :__nodef__)
        for d in reverse(ds)
            li_all = lhs_idx(d)
            length(li_all) == length(idx) || fail("cannot match '$(e.args[1])' against its definition in statement $(d.sid)")
            sub = Dict{Symbol,Any}()
            solved = Set{Int}()
            for (q, (li, ri)) in enumerate(zip(li_all, idx))
                if li isa Symbol && li in loop_vars(d) && !haskey(sub, li)
                    sub[li] = ri
                    push!(solved, q)
                end
            end
            conds = Any[]
            for (q, (li, ri)) in enumerate(zip(li_all, idx))
                q in solved || push!(conds, Expr(:call, :__eq__, ri, subst(li, sub)))
            end
            Set(keys(sub)) == Set(loop_vars(d)) || fail("'$(e.args[1])': a loop variable of statement $(d.sid) is not a subscript of its lhs")
            for (v, lo, hi) in d.loops
                push!(conds, Expr(:call, :__le__, subst(lo, sub), sub[v]))
                push!(conds, Expr(:call, :__le__, sub[v], subst(hi, sub)))
            end
            cond = conds[1]
            for c in conds[2:end]
                cond = Expr(:call, :__and__, cond, c)
            end
            out = Expr(:call, :__select__, cond, inline(lw, subst(d.rhs, sub), depth + 1), out)
        end
        return out
    end
    if e isa Expr
        start = e.head == :call ? 2 : 1
        return Expr(e.head, e.args[1:(start - 1)]..., Any[inline(lw, a, depth) for a in e.args[start:end]]...)
    end
    fail("cannot inline $(e)")
end

# ---- linear predictor (lowering.py _Lin / _linearize) -----------------------------------------------------------------------
# const + sum_k coef_k * cov_k with symbolic data expressions; `terms` keeps insertion order (ref => cov expr)
struct Lin
    cst::Any                       # expr or nothing (== 0)
    terms::Vector{Pair{Any,Any}}   # ref => cov ; ref = (:g, label) | (:l, name, idx exprs)
end
Lin() = Lin(nothing, Pair{Any,Any}[])
addx(a, b) = a === nothing ? b : (b === nothing ? a : Expr(:call, :+, a, b))
function lin_add(x::Lin, o::Lin, sgn)
    terms = copy(x.terms)
    oc = (sgn > 0 || o.cst === nothing) ? o.cst : Expr(:call, :-, o.cst)
    for (k, v) in o.terms
        v2 = sgn > 0 ? v : Expr(:call, :-, v)
        q = findfirst(t -> t.first == k, terms)
        if q === nothing
            push!(terms, k => v2)
        else
            terms[q] = k => addx(terms[q].second, v2)
        end
    end
    return Lin(addx(x.cst, oc), terms)
end
function lin_scale(x::Lin, c, divide=false)
    f = divide ? (v -> Expr(:call, :/, v, c)) : (v -> Expr(:call, :*, v, c))
    return Lin(x.cst === nothing ? nothing : f(x.cst), Pair{Any,Any}[k => f(v) for (k, v) in x.terms])
end
function linearize(lw::Lowering, e, lvs)::Lin
    is_data(lw, e, lvs) && return Lin(e, Pair{Any,Any}[])
    e isa Symbol && return Lin(nothing, Pair{Any,Any}[(:g, string(e)) => 1.0])
    if is_ref(e)
        lab = elem_label(lw, e)
        lab !== nothing && haskey(lw.gval_ids, lab) && return Lin(nothing, Pair{Any,Any}[(:g, lab) => 1.0])
        return Lin(nothing, Pair{Any,Any}[(:l, e.args[1], e.args[2:end]) => 1.0])
    end
    is_neg(e) && return lin_add(Lin(), linearize(lw, e.args[2], lvs), -1)
    if is_binop(e)
        op, a, b = e.args
        if op == :+ || op == :-
            return lin_add(linearize(lw, a, lvs), linearize(lw, b, lvs), op == :+ ? 1 : -1)
        elseif op == :*
            is_data(lw, a, lvs) && return lin_scale(linearize(lw, b, lvs), a)
            is_data(lw, b, lvs) && return lin_scale(linearize(lw, a, lvs), b)
            fail("the linear predictor is not linear in the parameters")
        elseif op == :/ && is_data(lw, b, lvs)
            return lin_scale(linearize(lw, a, lvs), b, true)
        end
    end
    fail("unsupported expression in a linear predictor: $(e)")
end

gval_of(lw::Lowering, name) = haskey(lw.gval_ids, name) ? lw.gval_ids[name] : fail("'$(name)' is not a scalar parameter or scalar logical node")

"prior / aux argument: CONST | GVAL | data array indexed by the plate's loop variables (lowering.py _scalar_arg)"
function scalar_arg(lw::Lowering, e, lvs, shp)
    is_data(lw, e, lvs) && return data_arg(lw, e, lvs, shp)
    e isa Symbol && haskey(lw.gval_ids, string(e)) && return Arg(ARG_GVAL, lw.gval_ids[string(e)])
    lab = elem_label(lw, e)
    lab !== nothing && haskey(lw.gval_ids, lab) && return Arg(ARG_GVAL, lw.gval_ids[lab])
    fail("distribution argument $(e) must be a constant, data, or a scalar parameter")
end

"a data expression over the plate: CONST | ONE | DATA_I | DATA_J | DATA_IJ (lowering.py _data_arg)"
function data_arg(lw::Lowering, e, lvs, shp)
    used = [v for v in free_vars(e) if v in lvs]
    if isempty(used)
        v = Float64(ev(lw.de, e, Dict{Symbol,Int}()))
        return (v == 1.0 && e isa Number) ? one_arg() : const_arg(v)
    end
    ext = lw.cur_extents
    key = string(e, "|", lvs, "|", ext)
    if length(lvs) == 1
        return Arg(ARG_DATA_I, slot!(lw, key, grid(lw.de, e, lvs, ext)))
    end
    i_name, j_name = lvs
    if !(j_name in used)
        return Arg(ARG_DATA_I, slot!(lw, key * "|I", grid(lw.de, e, [i_name], ext[1:1])))
    elseif !(i_name in used)
        return Arg(ARG_DATA_J, slot!(lw, key * "|J", grid(lw.de, e, [j_name], ext[2:2])))
    end
    return Arg(ARG_DATA_IJ, slot!(lw, key * "|IJ", grid(lw.de, e, lvs, ext); rank=2, dim0=shp[1]))   # column-major already
end

"Bijectors.bijector(dist): support-based (identity when the model is untransformed) (lowering.py _transform_of)"
function transform_of(lw::Lowering, fam, args)
    lb, ub = -Inf, Inf
    if fam in (FAM_GAMMA, FAM_EXPONENTIAL, FAM_LOGNORMAL, FAM_WEIBULL, FAM_CHISQ)
        lb = 0.0
    elseif fam == FAM_BETA
        lb, ub = 0.0, 1.0
    elseif fam == FAM_UNIFORM
        (is_data(lw, args[1], Symbol[]) && is_data(lw, args[2], Symbol[])) || fail("dunif bounds must be constants")
        lb, ub = Float64(ev(lw.de, args[1], Dict{Symbol,Int}())), Float64(ev(lw.de, args[2], Dict{Symbol,Int}()))
    end
    (!lw.transformed || fam in DISCRETE) && return TR_IDENTITY, lb, ub
    isfinite(lb) && isfinite(ub) && return TR_LOGIT, lb, ub
    isfinite(lb) && return TR_LOG, lb, ub
    return TR_IDENTITY, lb, ub
end
"""`x ~ truncated(dnorm(mu, tau), lower, upper)` for a scalar parameter with constant arguments and bounds (lowering.py
_truncation): the bounds are the prior's support and pick the bijector; the evaluators fold log(cdf(upper) - cdf(lower))
into the prior's constant (include/jbugs_plan.h, jbugs_global.lb / ub)"""
function truncation(lw::Lowering, s::Stmt, fam, dargs, sub)
    fam in (FAM_NORMAL, FAM_LOGNORMAL, FAM_LOGISTIC, FAM_LAPLACE, FAM_EXPONENTIAL, FAM_GAMMA, FAM_WEIBULL, FAM_CHISQ, FAM_UNIFORM,
            FAM_BETA, FAM_T) ||
        fail("'$(name_of(s))': T(,) is lowered for dnorm, dlnorm, dlogis, ddexp, dexp, dgamma, dweib, dchisqr, dunif, dbeta and dt priors")
    lw.transformed || fail("'$(name_of(s))': truncated priors are evaluated in transformed space only")
    all(a -> is_data(lw, a, Symbol[]), dargs) || fail("'$(name_of(s))': a truncated prior needs constant arguments")
    bounds = Float64[]
    for (b, none) in zip(s.trunc, (-Inf, Inf))
        if b === :nothing || b === nothing
            push!(bounds, none)
            continue
        end
        b = subst(b, sub)
        is_data(lw, b, Symbol[]) || fail("'$(name_of(s))': truncation bounds must be constants")
        push!(bounds, Float64(ev(lw.de, b, Dict{Symbol,Int}())))
    end
    _, nlo, nhi = transform_of(lw, fam, dargs)        # the family's own support
    lb, ub = max(bounds[1], nlo), min(bounds[2], nhi)
    lb < ub || fail("'$(name_of(s))': truncation needs lower < upper inside the support of the distribution")
    isfinite(lb) && isfinite(ub) && return TR_LOGIT, lb, ub
    isfinite(lb) && return TR_LOG, lb, ub
    isfinite(ub) && return TR_UPPER, lb, ub
    return TR_IDENTITY, lb, ub
end
function family(lw::Lowering, call)
    f = call.args[1]
    haskey(FAMILY_OF, f) || fail("distribution $(f) is outside the supported univariate families")
    fam = FAMILY_OF[f]
    length(call.args) - 1 == FAMILY_NARGS[fam] || fail("$(f) expects $(FAMILY_NARGS[fam]) arguments")
    return fam
end

"compress a theta-index array to base + stride_i * i + stride_j * j when exact (lowering.py _affine_or_index)"
function affine_or_index(t::Array{Int64})
    ndims(t) == 0 && return t[], 0, 0, false
    if ndims(t) == 1
        base = t[1]
        si = length(t) > 1 ? t[2] - t[1] : 0
        return all(t[k] == base + si * (k - 1) for k in eachindex(t)) ? (base, si, 0, false) : (0, 0, 0, true)
    end
    base = t[1, 1]
    si = size(t, 1) > 1 ? t[2, 1] - t[1, 1] : 0
    sj = size(t, 2) > 1 ? t[1, 2] - t[1, 1] : 0
    ok = all(t[i, j] == base + si * (i - 1) + sj * (j - 1) for i in axes(t, 1), j in axes(t, 2))
    return ok ? (base, si, sj, false) : (0, 0, 0, true)
end

# ---- scalar logical nodes of globals -> gtape (lowering.py _emit_gtape / _gexpr) -----------------------------------------
function gpush!(lw::Lowering, op, a, b=ZERO_ARG)
    push!(lw.plan.gtape, GOp(op, a, b))
    return Arg(ARG_GVAL, length(lw.plan.globals) + length(lw.plan.gtape) - 1)
end
function gexpr(lw::Lowering, e)::Arg
    is_data(lw, e, Symbol[]) && return const_arg(ev(lw.de, e, Dict{Symbol,Int}()))
    e isa Symbol && return Arg(ARG_GVAL, gval_of(lw, string(e)))
    lab = elem_label(lw, e)
    lab !== nothing && return Arg(ARG_GVAL, gval_of(lw, lab))
    is_neg(e) && return gpush!(lw, OP_NEG, gexpr(lw, e.args[2]))
    if is_binop(e)
        op = Dict(:+ => OP_ADD, :- => OP_SUB, :* => OP_MUL, :/ => OP_DIV, :^ => OP_POW)[e.args[1]]
        a = gexpr(lw, e.args[2])
        return gpush!(lw, op, a, gexpr(lw, e.args[3]))
    end
    if e isa Expr && e.head == :call
        f = e.args[1]
        if f == :pow
            a = gexpr(lw, e.args[2])
            return gpush!(lw, OP_POW, a, gexpr(lw, e.args[3]))
        end
        haskey(UNARY_OF, f) && length(e.args) == 2 && return gpush!(lw, UNARY_OF[f], gexpr(lw, e.args[2]))
    end
    fail("unsupported scalar logical expression $(e)")
end
function emit_gtape!(lw::Lowering, name, e)
    a = gexpr(lw, e)
    if a.kind != ARG_GVAL || a.id < length(lw.plan.globals)
        push!(lw.plan.gtape, GOp(OP_IDENTITY, a, ZERO_ARG))     # alias / constant: its own slot
        a = Arg(ARG_GVAL, length(lw.plan.globals) + length(lw.plan.gtape) - 1)
    end
    lw.gval_ids[name] = a.id
end

# ---- parameter vectors whose elements become scalar globals (lowering.py _vector_globals) -------------------------------
function collect_refs!(acc, e)
    if is_ref(e)
        push!(acc, e)
        foreach(i -> collect_refs!(acc, i), e.args[2:end])
    elseif e isa Expr
        foreach(a -> collect_refs!(acc, a), e.args[(e.head == :call ? 2 : 1):end])
    end
    return acc
end
function vector_globals(lw::Lowering)
    found = Set{Int}()
    looped_param(d) = d.kind == :param && !isempty(d.loops) && !d.is_gq
    # (1) vectors indexed by the INNER loop variable next to parameters indexed by the outer one (LSAT)
    for s in lw.stmts
        (s.is_gq || s.kind != :obs || length(s.loops) != 2) && continue
        exprs = try
            [inline(lw, a) for a in s.rhs.args[2:end]]
        catch err
            err isa LoweringError || rethrow()
            continue
        end
        refs = Any[]
        foreach(e -> collect_refs!(refs, e), exprs)
        prm = [r for r in refs if any(looped_param, defs(lw, r.args[1]))]
        lv = loop_vars(s)
        outer = [r for r in prm if any(i -> i === lv[1], r.args[2:end])]
        inner = [r for r in prm if length(r.args) == 2 && r.args[2] === lv[2]]
        if !isempty(outer) && !isempty(inner)
            for r in inner, d in defs(lw, r.args[1])
                if d.kind == :param && length(d.loops) == 1 && d.extents == s.extents[2:2] && size_of(d) <= 10
                    push!(found, d.sid)
                end
            end
        end
        # ... and vectors picked through a data index of the inner loop variable next to parameters indexed by the outer
        # one (LeukFr: exp(beta * Z[i] + b[pair[i]]) * dL0[j]): read on the tape as gval[base + index] (OP_GVEC_AT)
        if !isempty(outer)
            for r in prm
                (length(r.args) == 2 && !(r.args[2] isa Symbol) && !is_range(r.args[2]) && !is_colon(r.args[2]) &&
                 is_data(lw, r.args[2], lv) && !is_data(lw, r.args[2], Symbol[]) && !(lv[1] in free_vars(r.args[2]))) || continue
                for d in defs(lw, r.args[1])
                    (d.kind == :param && length(d.loops) == 1 && size_of(d) <= 22) && push!(found, d.sid)
                end
            end
        end
    end
    # (2) elements referenced with constant subscripts by an observation or a scalar logical node (Stacks: beta[1] * z[i,1])
    for s in lw.stmts
        s.is_gq && continue
        (s.kind == :obs || (s.kind == :logical && isempty(s.loops))) || continue
        exprs = try
            s.kind == :obs ? [inline(lw, a) for a in s.rhs.args[2:end]] : [inline(lw, s.rhs)]
        catch err
            err isa LoweringError || rethrow()
            continue
        end
        refs = Any[]
        foreach(e -> collect_refs!(refs, e), exprs)
        for r in refs
            elem_label(lw, r) === nothing && continue
            for d in defs(lw, r.args[1])
                if looped_param(d)
                    size_of(d) > 48 && fail("'$(r.args[1])' has $(size_of(d)) elements, too many to expand into scalar globals; write the predictor as inprod(X[i, 1:P], beta[1:P]) for the dense path")
                    push!(found, d.sid)
                end
            end
        end
    end
    # (3) parameter vectors subscripted by a discrete latent that is summed out, directly (mu[z[i]]) or through a derived
    # subscript (x1[i] = x[i] + 1; phi[x1[i], d1[i]]), and the looped parameters behind a logical vector selected that way
    latents = Set{Symbol}(name_of(d) for d in lw.stmts if d.kind == :latent)
    if !isempty(latents)
        function mentions_latent(i)
            (is_range(i) || is_colon(i)) && return false
            e = try
                inline(lw, i)
            catch err
                err isa LoweringError || rethrow()
                return false
            end
            return any(r -> (r isa Symbol ? r : r.args[1]) in latents, latent_candidates(e))
        end
        function walk_latent(e, via::Bool)
            if is_ref(e)
                if via || any(mentions_latent, e.args[2:end])
                    for d in defs(lw, e.args[1])
                        if looped_param(d)
                            size_of(d) > 48 && fail("'$(e.args[1])' has too many elements to expand into scalar globals")
                            push!(found, d.sid)
                        elseif d.kind == :logical && !isempty(d.loops)
                            walk_latent(d.rhs, true)
                        end
                    end
                end
                foreach(i -> walk_latent(i, false), e.args[2:end])
            elseif e isa Expr && e.head == :call
                foreach(a -> walk_latent(a, via), e.args[2:end])
            end
        end
        for s in lw.stmts
            (s.kind == :obs && !s.is_gq) && foreach(a -> walk_latent(a, false), s.rhs.args[2:end])
        end
    end
    return found
end

"every symbol and `name[...]` reference inside `e` (candidates for a reference to a latent), subscripts included"
function latent_candidates(e, acc=Any[])
    if e isa Symbol
        push!(acc, e)
    elseif is_ref(e)
        push!(acc, e)
        foreach(i -> latent_candidates(i, acc), e.args[2:end])
    elseif e isa Expr && e.head == :call
        foreach(a -> latent_candidates(a, acc), e.args[2:end])
    end
    return acc
end
"`e` with every occurrence of the sub-expression `target` replaced by `new`"
function replace_sub(e, target, new)
    e == target && return new
    e isa Expr || return e
    return Expr(e.head, Any[k == 1 && (e.head == :call || e.head == :ref) ? a : replace_sub(a, target, new) for (k, a) in enumerate(e.args)]...)
end

# ---- one plate (lowering.py _build_plate ...) ------------------------------------------------------------------------------------
"(family, eta_arg, inlined args, link ops innermost-first reversed, predictor expression) of an observation"
function peel(lw::Lowering, s::Stmt)
    fam = family(lw, s.rhs)
    eta_arg = fam in (FAM_GAMMA, FAM_WEIBULL) ? 1 : 0
    args = Any[inline(lw, a) for a in s.rhs.args[2:end]]
    eta = args[eta_arg + 1]
    link = Int[]
    while eta isa Expr && eta.head == :call && haskey(LINK_OF, eta.args[1]) && length(eta.args) == 2 && !is_data(lw, eta, loop_vars(s))
        push!(link, LINK_OF[eta.args[1]])
        eta = eta.args[2]
    end
    return fam, eta_arg, args, reverse(link), eta
end

function local_index!(lw::Lowering, plate::Plate, obs::Stmt, ref, used::Set{Int})
    _, name, idx = ref
    k = findfirst(l -> l.name == string(name), plate.locals)
    k !== nothing && return k - 1
    ds = [d for d in defs(lw, name) if d.kind == :param]
    length(ds) == 1 || fail("'$(name)' must be defined by exactly one stochastic statement")
    d = ds[1]
    d.is_gq && fail("'$(name)' feeds an observation but was classified as generated quantity")
    d.sid in used && fail("'$(name)' is a parent of more than one observation plate")
    (length(d.loops) > length(obs.loops) || isempty(d.loops)) && fail("'$(name)': unsupported nesting")
    d.extents == obs.extents[1:length(d.loops)] || fail("'$(name)': loop ranges differ from the observation plate's")
    dl = lhs_idx(d)
    (length(dl) == length(d.loops) && length(idx) == length(dl)) || fail("'$(name)': lhs must be indexed by its loop variables")
    for (pos, (li, ri)) in enumerate(zip(dl, idx))
        li === loop_vars(d)[pos] || fail("'$(name)': lhs index $(pos) must be loop variable '$(loop_vars(d)[pos])'")
        ri === loop_vars(obs)[pos] || fail("'$(name)': must be referenced with the plate's own loop variables")
    end
    level = length(d.loops) == 2 ? LEVEL_OBS : LEVEL_GROUP
    fam = family(lw, d.rhs)
    tr, lb, ub = transform_of(lw, fam, d.rhs.args[2:end])
    tidx = theta_index(lw, d)
    base, si, sj, need_idx = affine_or_index(tidx)
    idx_slot = need_idx ? slot!(lw, "theta_index|$(name)", tidx; rank=ndims(tidx), dim0=size(tidx, 1)) : -1
    saved = lw.cur_extents
    lw.cur_extents = d.extents
    pargs = Arg[scalar_arg(lw, a, loop_vars(d), shape_of(d)) for a in d.rhs.args[2:end]]
    lw.cur_extents = saved
    push!(plate.locals, Local(string(name), level, tr, base, si, sj, fam, pargs, lb, ub, idx_slot))
    push!(used, d.sid)
    return length(plate.locals) - 1
end

function apply_obs_mask!(lw::Lowering, plate::Plate, s::Stmt)
    s.obs_mask === nothing && return
    (plate.family == FAM_T || plate.family == FAM_BETABIN) && fail("dt observations with missing cells are not supported")
    m = Float64.(s.obs_mask)
    plate.weight = ndims(m) == 1 ? Arg(ARG_DATA_I, slot!(lw, "obs_mask|$(name_of(s))", m)) :
                   Arg(ARG_DATA_IJ, slot!(lw, "obs_mask|$(name_of(s))", m; rank=2, dim0=size(m, 1)))
    plate.n_real = Int(sum(m))
    ys = lw.plan.data[plate.y.id + 1]
    good = findfirst(!isnan, ys.values)
    fill = good === nothing ? 0.0 : ys.values[good]
    ys.values = [isnan(v) ? fill : v for v in ys.values]
end

is_gvec_ref(lw::Lowering, s::Stmt, ref) =
    length(s.loops) == 2 && length(ref[3]) == 1 && ref[3][1] === loop_vars(s)[2] &&
    haskey(lw.gval_ids, label(ref[2], [s.extents[2][1]]))

function build_linear_plate!(lw::Lowering, s0::Stmt, used::Set{Int})
    s = s0
    length(s.loops) > 2 && fail("observation plates deeper than two loops are outside the plate class")
    if length(s.loops) == 2
        # random effects indexed by the INNER loop variable only (Equiv): build the plate with the loops swapped
        firsts = try
            _, _, _, _, eta = peel(lw, s)
            Set(r[3][1] for (r, _) in linearize(lw, eta, loop_vars(s)).terms if r[1] == :l && !isempty(r[3]) && r[3][1] isa Symbol)
        catch err
            err isa LoweringError || rethrow()
            Set{Any}()
        end
        if firsts == Set{Any}([loop_vars(s)[2]])
            s = Stmt(s.sid, s.stoch, s.lhs, s.rhs, reverse(s.loops), reverse(s.extents), s.kind, s.is_gq, s.ragged,
                     s.obs_mask === nothing ? nothing : permutedims(s.obs_mask), s.zobs === nothing ? nothing : permutedims(s.zobs), s.trunc)
        end
    end
    lvs, shp = loop_vars(s), shape_of(s)
    lw.cur_extents = s.extents
    N = isempty(shp) ? 1 : shp[1]
    T = length(shp) == 2 ? shp[2] : 1
    fam = family(lw, s.rhs)
    eta_arg = fam in (FAM_GAMMA, FAM_WEIBULL) ? 1 : 0
    fam in (FAM_UNIFORM, FAM_BETA, FAM_FLAT) && fail("$(s.rhs.args[1]) as an observation family is outside the plate class")
    plate = Plate(N, T; family=fam, eta_arg=eta_arg, name="$(name_of(s))~$(s.rhs.args[1])")
    yarg = isempty(shp) ? const_arg(lw.de.data[name_of(s)]) : data_arg(lw, s.lhs, lvs, shp)
    (yarg.kind == ARG_CONST && !isempty(shp)) && fail("observation does not vary over its plate")
    plate.y = yarg
    args = Any[inline(lw, a) for a in s.rhs.args[2:end]]
    eta = args[eta_arg + 1]
    link = Int[]
    while eta isa Expr && eta.head == :call && haskey(LINK_OF, eta.args[1]) && length(eta.args) == 2 && !is_data(lw, eta, lvs)
        push!(link, LINK_OF[eta.args[1]])
        eta = eta.args[2]
    end
    plate.link = reverse(link)
    if is_call(eta, :inprod)
        build_dense!(lw, s, fam, plate, eta, args, used)
        return nothing
    end
    fam == FAM_CATEGORICAL && fail("dcat needs an expression plate")
    lin = linearize(lw, eta, lvs)
    any(r -> r[1] == :l && is_gvec_ref(lw, s, r), first.(lin.terms)) && fail("a parameter vector indexed by the inner loop variable")
    gidx = gather_index(lw, s, lin)
    gidx !== nothing && return build_gathered_plate!(lw, s, fam, eta_arg, plate.link, args, lin, gidx, used)
    for (k, a) in enumerate(args)
        push!(plate.aux, k - 1 == eta_arg ? ZERO_ARG : scalar_arg(lw, a, lvs, shp))
    end
    for (ref, cov) in lin.terms
        ref[1] == :g && push!(plate.terms, Term(Arg(ARG_GVAL, gval_of(lw, ref[2])), data_arg(lw, cov, lvs, shp)))
    end
    for (ref, cov) in lin.terms
        ref[1] == :g && continue
        li = local_index!(lw, plate, s, ref, used)
        push!(plate.terms, Term(Arg(ARG_LOCAL, li), data_arg(lw, cov, lvs, shp)))
    end
    if lin.cst !== nothing
        c = data_arg(lw, lin.cst, lvs, shp)
        (c.kind == ARG_CONST && c.value == 0.0) || push!(plate.terms, Term(const_arg(1.0), c))
    end
    apply_obs_mask!(lw, plate, s)
    return plate
end

"linear-predictor plate when possible, else an expression plate (lowering.py _build_plate)"
function build_plate!(lw::Lowering, s::Stmt, used::Set{Int})
    before, n_slots = copy(used), length(lw.plan.data)
    try
        return build_linear_plate!(lw, s, used)
    catch first_err
        first_err isa LoweringError || rethrow()
        empty!(used); union!(used, before)
        for sl in lw.plan.data[(n_slots + 1):end]
            delete!(lw.slot_ids, sl.name)
        end
        resize!(lw.plan.data, n_slots)
        try
            return build_xplate!(lw, s, family(lw, s.rhs), used)
        catch second_err
            second_err isa LoweringError || rethrow()
            fail("$(second_err.msg) (as a linear-predictor plate: $(first_err.msg))")
        end
    end
end

# ---- expression plates (lowering.py _build_xplate / _tape) ---------------------------------------------------------------------
function build_xplate!(lw::Lowering, s::Stmt, fam, used::Set{Int})
    lvs, shp = loop_vars(s), shape_of(s)
    lw.cur_extents = s.extents
    plate = Plate(isempty(shp) ? 1 : shp[1], length(shp) == 2 ? shp[2] : 1; family=fam, eta_arg=0,
                  name="$(name_of(s))~$(s.rhs.args[1]) (expression tape)")
    if isempty(shp)
        # a scalar observation (y ~ dnorm(mu + delta[z], tau)): a plate of one cell
        plate.y = Arg(ARG_DATA_I, slot!(lw, "scalar_obs|$(name_of(s))", [Float64(ev(lw.de, s.lhs, Dict{Symbol,Int}()))]))
    else
        plate.y = data_arg(lw, s.lhs, lvs, shp)
    end
    plate.y.kind in (ARG_CONST, ARG_ONE, ARG_DATA_J) && fail("observation does not vary over its plate")
    args = Any[s.rhs.args[2:end]...]
    if fam == FAM_CATEGORICAL
        pa = args[1]
        (is_ref(pa) && (is_range(pa.args[end]) || is_colon(pa.args[end])) && !any(i -> is_range(i) || is_colon(i), pa.args[2:(end - 1)])) ||
            fail("dcat expects a probability vector p[..., 1:K]")
        args = Any[Expr(:ref, pa.args[1:(end - 1)]..., s.lhs)]      # p[.., y]: the element picked by the observed category
    end
    tape, memo = XOp[], Dict{String,Int}()
    lw.latent = nothing      # (statement of the discrete latent summed out in this plate, K), found while taping
    lw.de.lenient = true
    try
        for a in args
            e = inline(lw, a)
            if is_data(lw, e, lvs)
                push!(plate.aux, data_arg(lw, e, lvs, shp))
            elseif e isa Symbol && haskey(lw.gval_ids, string(e))
                push!(plate.aux, Arg(ARG_GVAL, lw.gval_ids[string(e)]))
            else
                push!(plate.aux, Arg(ARG_XVAL, tape!(lw, e, s, plate, tape, memo, used)))
            end
        end
    finally
        lw.de.lenient = false
    end
    ((fam == FAM_T || fam == FAM_BETABIN) && plate.aux[3].kind != ARG_CONST) && fail("the degrees of freedom of dt must be a constant")
    isempty(tape) && fail("internal: an expression plate without a tape")
    if lw.latent !== nothing
        z, K = lw.latent
        wa = z.rhs.args[2]
        push_raw(op::XOp) = (push!(tape, op); length(tape) > MAX_XOPS && fail("the predictor needs more than $(MAX_XOPS) tape ops"); length(tape) - 1)
        lw.de.lenient = true
        try
            if z.rhs.args[1] == :dbern
                # log p(z = 0) = log(1 - theta), log p(z = 1) = log(theta): two candidates picked by the state
                c0 = tape!(lw, Expr(:call, :log, Expr(:call, :-, 1.0, wa)), s, plate, tape, memo, used)
                c1 = tape!(lw, Expr(:call, :log, wa), s, plate, tape, memo, used)
                first = length(tape)
                push_raw(XOp(OP_IDENTITY, c0)); push_raw(XOp(OP_IDENTITY, c1))
                plate.logprior_slot = push_raw(XOp(OP_PICK_K, first))
            else
                # log p(z = k): the latent's own dcat(w[1:K]) picked at the state
                pk = Expr(:ref, wa.args[1:(end - 1)]..., z.lhs)
                plate.logprior_slot = tape!(lw, Expr(:call, :log, pk), s, plate, tape, memo, used)
            end
            for t in get(lw.latent_sibs, s.sid, Stmt[])
                sl = tape!(lw, inline(lw, sibling_logpdf(t, s)), s, plate, tape, memo, used)
                plate.logprior_slot = push_raw(XOp(OP_ADD, plate.logprior_slot, sl))
                plate.name *= " + $(name_of(t))~$(t.rhs.args[1])"
            end
        finally
            lw.de.lenient = false
        end
        plate.n_states = K
        if z.zobs !== nothing
            (length(z.loops) in (1, 2) && length(z.loops) == length(s.loops)) ||
                fail("'$(name_of(z))': a partially observed latent must be indexed like its observation")
            zo = z.zobs
            any(v -> v != floor(v) || v < 0 || v > K, zo) && fail("'$(name_of(z))': observed states must be integers in 1..$(K)")
            leaf = ndims(zo) == 1 ? Arg(ARG_DATA_I, slot!(lw, "latent_obs|$(name_of(z))", copy(zo))) :
                   Arg(ARG_DATA_IJ, slot!(lw, "latent_obs|$(name_of(z))", copy(zo); rank=2, dim0=size(zo, 1)))
            plate.zobs_slot = push_raw(XOp(OP_LEAF, 0, 0, 0, leaf))
        end
        plate.name *= ", $(name_of(z)) summed out over $(K) states"
        push!(used, z.sid)
        lw.latent = nothing
    elseif !isempty(get(lw.latent_sibs, s.sid, Stmt[]))
        fail("'$(name_of(s))': internal: the shared latent was not found while taping")
    end
    plate.xops = tape
    apply_obs_mask!(lw, plate, s)
    return plate
end

"(latent statement, K) when `ie` is a reference `z[i]` / `z` to a discrete latent of this plate that is being summed out, else nothing (lowering.py _latent_ref)"
function latent_ref(lw::Lowering, ie, s::Stmt)
    (ie isa Symbol || is_ref(ie)) || return nothing
    nm = ie isa Symbol ? ie : ie.args[1]
    ds = [d for d in defs(lw, nm) if d.kind == :latent]
    isempty(ds) && return nothing
    d = ds[1]
    (length(ds) == 1 && length(d.loops) <= length(s.loops) && d.loops == s.loops[1:length(d.loops)] && !d.is_gq) ||
        fail("'$(nm)': a discrete latent must share the loop nest of the observation it selects for")
    idx = ie isa Symbol ? Any[] : ie.args[2:end]
    string.(idx) == string.(lhs_idx(d)) || fail("'$(nm)' must be referenced with the subscripts of its own statement")
    K = 2   # dbern: states 0, 1 (value of state k: k); a dcat latent has the states 1..K (value of state k: k + 1)
    if d.rhs.args[1] != :dbern
        wa = d.rhs.args[2]
        (is_ref(wa) && is_range(wa.args[end]) && is_data(lw, wa.args[end].args[2], Symbol[]) && is_data(lw, wa.args[end].args[3], Symbol[]) &&
         Int(ev(lw.de, wa.args[end].args[2], Dict{Symbol,Int}())) == 1) ||
            fail("the latent's dcat needs a probability vector w[..., 1:K] with constant K")
        K = Int(ev(lw.de, wa.args[end].args[3], Dict{Symbol,Int}()))
        1 <= K <= MAX_STATES || fail("between 1 and $(MAX_STATES) states can be summed out per observation")
    end
    (lw.latent !== nothing && lw.latent[1].sid != d.sid) && fail("one discrete latent per observation plate")
    lw.latent = (d, K)
    return d, K
end

"sids of the summed-out latent statements the arguments of observation statement `s` depend on (lowering.py _latents_of)"
function latents_of(lw::Lowering, s::Stmt)
    latents = Dict{Symbol,Int}(name_of(d) => d.sid for d in lw.stmts if d.kind == :latent)
    out = Int[]
    isempty(latents) && return out
    for a in s.rhs.args[2:end]
        e = try
            inline(lw, a)
        catch err
            err isa LoweringError || rethrow()
            continue
        end
        for r in latent_candidates(e)
            nm = r isa Symbol ? r : r.args[1]
            (haskey(latents, nm) && !(latents[nm] in out)) && push!(out, latents[nm])
        end
    end
    return out
end

"""log density of observation statement `t` -- a second child of the latent summed out in the plate of `s` -- as an
expression for the tape (lowering.py _sibling_logpdf); the family's argument checks are not reproduced for a folded statement"""
function sibling_logpdf(t::Stmt, s::Stmt)
    (t.loops == s.loops && !t.ragged && t.obs_mask === nothing) ||
        fail("'$(name_of(t))': a second observation of a summed-out latent must be fully observed and share the loop nest of the first")
    y, fn, a = t.lhs, t.rhs.args[1], t.rhs.args[2:end]
    c(f, xs...) = Expr(:call, f, xs...)
    fn == :dbern && return c(:__select__, y, c(:log, a[1]), c(:log, c(:-, 1.0, a[1])))
    if fn == :dbin
        norm = c(:-, c(:-, c(:loggam, c(:+, a[2], 1.0)), c(:loggam, c(:+, y, 1.0))), c(:loggam, c(:+, c(:-, a[2], y), 1.0)))
        return c(:+, c(:+, c(:*, y, c(:log, a[1])), c(:*, c(:-, a[2], y), c(:log, c(:-, 1.0, a[1])))), norm)
    end
    if fn == :dnorm
        r = c(:-, y, a[1])
        return c(:-, c(:-, c(:*, 0.5, c(:log, a[2])), c(:*, c(:*, 0.5, a[2]), c(:*, r, r))), 0.5 * log(2.0 * pi))
    end
    fn == :dpois && return c(:-, c(:-, c(:*, y, c(:log, a[1])), a[1]), c(:logfact, y))
    fail("'$(name_of(t))': a second observation of a summed-out latent must be dbern, dbin, dnorm or dpois")
end

"(every leaf of `e` is a global value or a constant, at least one is a global value) (lowering.py _global_only)"
function global_only(lw::Lowering, e)
    e isa Number && return true, false
    if e isa Symbol
        haskey(lw.gval_ids, string(e)) && return true, true
        return is_data(lw, e, Symbol[]), false
    end
    if is_ref(e)
        lab = elem_label(lw, e)
        (lab !== nothing && haskey(lw.gval_ids, lab)) && return true, true
        return (!any(i -> is_range(i) || is_colon(i), e.args[2:end]) && is_data(lw, e, Symbol[])), false
    end
    if e isa Expr && e.head == :call && (is_neg(e) || is_binop(e) || e.args[1] == :pow || haskey(UNARY_OF, e.args[1])) && length(e.args) >= 2
        rs = [global_only(lw, a) for a in e.args[2:end]]
        return all(r -> r[1], rs), any(r -> r[2], rs)
    end
    return false, false
end

"""(first global value slot, number of elements, 0-based index expression) when `e` is `name[idx]` with `name` a parameter
vector whose elements were expanded into consecutive globals and `idx` a data expression that varies over the plate
(lowering.py _gvec_at), else nothing"""
function gvec_at(lw::Lowering, e, s::Stmt)
    lvs = loop_vars(s)
    idx = e.args[2:end]
    (length(idx) == 1 && !is_range(idx[1]) && !is_colon(idx[1]) && is_data(lw, idx[1], lvs) && !is_data(lw, idx[1], Symbol[])) || return nothing
    ds = [d for d in defs(lw, e.args[1]) if d.kind == :param && !d.is_gq]
    (length(ds) == 1 && length(ds[1].loops) == 1) || return nothing
    lo, hi = ds[1].extents[1]
    ids = [get(lw.gval_ids, label(e.args[1], [k]), -1) for k in lo:hi]
    (all(>=(0), ids) && ids == collect(ids[1]:(ids[1] + length(ids) - 1))) || return nothing
    vals = grid(lw.de, idx[1], lvs, s.extents)
    seen = s.obs_mask === nothing ? trues(size(vals)) : s.obs_mask
    v = vals[seen]
    (any(isnan, v) || any(x -> x != floor(x), v) || minimum(v) < lo || maximum(v) > hi) &&
        fail("'$(e)': the data index leaves the parameter vector")
    return ids[1], length(ids), Expr(:call, :-, idx[1], Float64(lo))
end

"""`phi[2, d1[i]]`: an element of a short parameter array whose elements are globals, picked by data subscripts that vary
over the plate -> nested selects over the index tuples that occur (lowering.py _global_pick)"""
function global_pick(lw::Lowering, e, s::Stmt)
    prefix = string(e.args[1], "[")
    any(k -> startswith(k, prefix), keys(lw.gval_ids)) || return nothing
    lvs = loop_vars(s)
    any(i -> is_range(i) || is_colon(i) || !is_data(lw, i, lvs), e.args[2:end]) && return nothing
    cols = [vec(grid(lw.de, i, lvs, s.extents)) for i in e.args[2:end]]
    seen = s.obs_mask === nothing ? trues(length(cols[1])) : vec(s.obs_mask)
    any(col -> any(isnan, col[seen]), cols) && fail("'$(e)': a subscript is missing for an observed cell")
    tuples = sort(unique([Tuple(Int(col[k]) for col in cols) for k in eachindex(seen) if seen[k]]))
    (isempty(tuples) || length(tuples) > 8) && fail("'$(e)': too many distinct elements picked by data subscripts")
    out = nothing
    for t in tuples
        elem = Expr(:ref, e.args[1], (Float64(v) for v in t)...)
        haskey(lw.gval_ids, something(elem_label(lw, elem), "")) || fail("'$(elem)' is not a parameter element")
        if out === nothing
            out = elem
            continue
        end
        cond = nothing
        for (i, v) in zip(e.args[2:end], t)
            ci = Expr(:call, :equals, i, Float64(v))
            cond = cond === nothing ? ci : Expr(:call, :*, cond, ci)
        end
        out = Expr(:call, :__select__, cond, elem, out)
    end
    return out
end

"append the ops computing `e` for one cell of the plate (common subexpressions shared); returns the 0-based slot"
function tape!(lw::Lowering, e, s::Stmt, plate::Plate, tape::Vector{XOp}, memo::Dict{String,Int}, used::Set{Int})
    key = string(e)
    haskey(memo, key) && return memo[key]
    function push_op(op::XOp)
        push!(tape, op)
        length(tape) > MAX_XOPS && fail("the predictor needs more than $(MAX_XOPS) tape ops")
        memo[key] = length(tape) - 1
        return memo[key]
    end
    lvs, shp = loop_vars(s), shape_of(s)
    rec(x) = tape!(lw, x, s, plate, tape, memo, used)
    if e isa Expr && e.head == :call
        # a sub-expression of globals and constants only (log(1 - q)) does not vary over the plate: it becomes a scalar
        # logical node on the gtape -- once per chain in the prologue -- and the tape reads its value (lowering.py _tape)
        only, has = global_only(lw, e)
        if only && has
            n_gt = length(lw.plan.gtape)
            a = try
                gexpr(lw, e)
            catch err
                err isa LoweringError || rethrow()
                resize!(lw.plan.gtape, n_gt)
                nothing
            end
            a === nothing || return push_op(XOp(OP_LEAF, 0, 0, 0, a))
        end
    end
    if is_call(e, :__select__)
        cond, a, b = e.args[2:4]
        m = grid(lw.de, cond, lvs, s.extents) .!= 0
        live = s.obs_mask === nothing ? m : (m .| .!s.obs_mask)
        dead = s.obs_mask === nothing ? .!m : (.!m .| .!s.obs_mask)
        all(live) && return (memo[key] = rec(a))
        all(dead) && return (memo[key] = rec(b))
        ia, ib = rec(a), rec(b)
        ic = rec(Expr(:call, :__mask__, cond))
        return push_op(XOp(OP_SELECT, ic, ia, ib))
    end
    if is_call(e, :__mask__)
        m = Float64.(grid(lw.de, e.args[2], lvs, s.extents) .!= 0)
        k = "mask|$(name_of(s))|$(e.args[2])"
        leaf = ndims(m) == 1 ? Arg(ARG_DATA_I, slot!(lw, k, m)) : Arg(ARG_DATA_IJ, slot!(lw, k, m; rank=2, dim0=shp[1]))
        return push_op(XOp(OP_LEAF, 0, 0, 0, leaf))
    end
    is_call(e, :__nodef__) && return push_op(XOp(OP_LEAF, 0, 0, 0, const_arg(NaN)))
    is_data(lw, e, lvs) && return push_op(XOp(OP_LEAF, 0, 0, 0, data_arg(lw, e, lvs, shp)))
    # K candidate slots gathered into consecutive identity copies behind a PICK_K op
    function pick_k(slots::Vector{Int})
        first = length(tape)
        for sl in slots
            push!(tape, XOp(OP_IDENTITY, sl))
        end
        return push_op(XOp(OP_PICK_K, first))
    end
    if lw.marginalize && (e isa Symbol || is_ref(e)) && latent_ref(lw, e, s) !== nothing
        # the latent's own value in arithmetic (mu + delta * z[i]): a constant per state
        z, K = latent_ref(lw, e, s)
        v0 = z.rhs.args[1] == :dbern ? 0 : 1
        first = length(tape)
        for k in 0:(K - 1)
            push!(tape, XOp(OP_LEAF, 0, 0, 0, const_arg(Float64(v0 + k))))
        end
        return push_op(XOp(OP_PICK_K, first))
    end
    e isa Symbol && return push_op(XOp(OP_LEAF, 0, 0, 0, Arg(ARG_GVAL, gval_of(lw, string(e)))))
    if is_ref(e)
        idx = e.args[2:end]
        if lw.marginalize
            lat = [q for (q, ie) in enumerate(idx) if latent_ref(lw, ie, s) !== nothing]
            length(lat) > 1 && fail("a reference may use the discrete latent in one subscript only")
            if length(lat) == 1
                # name[.., z[i], ..]: the K candidates name[.., k, ..] in consecutive slots, picked by the state
                z, K = latent_ref(lw, idx[lat[1]], s)
                return pick_k(Int[rec(inline(lw, Expr(:ref, e.args[1], idx[1:(lat[1] - 1)]..., Float64(k), idx[(lat[1] + 1):end]...))) for k in 1:K])
            end
            # a subscript that is a FUNCTION of the latent (x1[i] = x[i] + 1; phi[x1[i], ..]): evaluated per state
            for (q, ie) in enumerate(idx)
                (is_range(ie) || is_colon(ie) || is_data(lw, ie, lvs)) && continue
                ie2 = inline(lw, ie)
                refs = [r for r in latent_candidates(ie2) if latent_ref(lw, r, s) !== nothing]
                isempty(refs) && continue
                z, K = latent_ref(lw, refs[1], s)
                v0 = z.rhs.args[1] == :dbern ? 0 : 1
                slots = Int[]
                for k in 0:(K - 1)
                    sub = replace_sub(ie2, refs[1], Float64(v0 + k))
                    is_data(lw, sub, Symbol[]) || fail("a subscript built from the discrete latent must not depend on anything else that varies")
                    cidx = Float64(Int(ev(lw.de, sub, Dict{Symbol,Int}())))
                    push!(slots, rec(inline(lw, Expr(:ref, e.args[1], idx[1:(q - 1)]..., cidx, idx[(q + 1):end]...))))
                end
                return pick_k(slots)
            end
        end
        lab = elem_label(lw, e)
        lab !== nothing && haskey(lw.gval_ids, lab) && return push_op(XOp(OP_LEAF, 0, 0, 0, Arg(ARG_GVAL, lw.gval_ids[lab])))
        ref = (:l, e.args[1], e.args[2:end])
        is_gvec_ref(lw, s, ref) &&
            return push_op(XOp(OP_LEAF, 0, 0, 0, Arg(ARG_GVEC_J, lw.gval_ids[label(e.args[1], [s.extents[2][1]])])))
        at = gvec_at(lw, e, s)
        if at !== nothing
            # b[pair[i]]: an element of a parameter vector (consecutive global values) picked by a data index
            base, cnt, idx0 = at
            ia = rec(idx0)
            lf = tape[ia + 1]
            if lf.op == OP_LEAF && lf.leaf.kind in (ARG_CONST, ARG_ONE)      # the same element everywhere
                k0 = Int(lf.leaf.kind == ARG_ONE ? 1.0 : lf.leaf.value)
                return push_op(XOp(OP_LEAF, 0, 0, 0, Arg(ARG_GVAL, base + k0)))
            end
            return push_op(XOp(OP_GVEC_AT, ia, 0, 0, Arg(ARG_GVAL, Int32(base), Float64(cnt))))
        end
        picked = global_pick(lw, e, s)
        picked !== nothing && return (memo[key] = rec(picked))
        return push_op(XOp(OP_LEAF, 0, 0, 0, Arg(ARG_LOCAL, local_index!(lw, plate, s, ref, used))))
    end
    is_neg(e) && return push_op(XOp(OP_NEG, rec(e.args[2])))
    if is_binop(e)
        op = Dict(:+ => OP_ADD, :- => OP_SUB, :* => OP_MUL, :/ => OP_DIV, :^ => OP_POW)[e.args[1]]
        ia = rec(e.args[2]); ib = rec(e.args[3])
        return push_op(XOp(op, ia, ib))
    end
    if e isa Expr && e.head == :call
        f = e.args[1]
        if f == :pow && length(e.args) == 3
            ia = rec(e.args[2]); ib = rec(e.args[3])
            return push_op(XOp(OP_POW, ia, ib))
        end
        haskey(UNARY_OF, f) && length(e.args) == 2 && return push_op(XOp(UNARY_OF[f], rec(e.args[2])))
        if f == :logit && length(e.args) == 2
            # LogExpFunctions.logit(x) = log(x / (1 - x)), written out in tape arithmetic (lowering.py _tape)
            x = e.args[2]
            return (memo[key] = rec(Expr(:call, :log, Expr(:call, :/, x, Expr(:call, :-, 1.0, x)))))
        end
    end
    fail("unsupported expression in a predictor: $(e)")
end

# ---- random effects picked by a data index (lowering.py _gather_index / _build_gathered_plate) ------------------------------
function gather_index(lw::Lowering, s::Stmt, lin::Lin)
    refs = [r for (r, _) in lin.terms if r[1] == :l]
    (isempty(refs) || length(s.loops) != 1) && return nothing
    i_name = loop_vars(s)[1]
    plain = [length(r[3]) == 1 && r[3][1] === i_name for r in refs]
    (all(plain) || any(r -> length(r[3]) != 1, refs)) && return nothing
    any(plain) && fail("random effects indexed both by the loop variable and by a data index in one predictor")
    exprs = Set(string(r[3][1]) for r in refs)
    (length(exprs) == 1 && is_data(lw, refs[1][3][1], loop_vars(s))) ||
        fail("random effects of one predictor must share one data index expression (b[group[i]])")
    return refs[1][3][1]
end

"an argument that takes the value cols[k][i] in column k: CONST, ONE, DATA_J, DATA_I or DATA_IJ (lowering.py _columns_arg)"
function columns_arg(lw::Lowering, key, cols::Vector{Vector{Float64}}, N)
    K = length(cols)
    M = Matrix{Float64}(undef, N, K)
    for k in 1:K
        M[:, k] .= cols[k]
    end
    if N > 0 && all(M .== M[1:1, :])
        row = M[1, :]
        all(row .== row[1]) && return row[1] == 1.0 ? one_arg() : const_arg(row[1])
        return Arg(ARG_DATA_J, slot!(lw, key * "|J", row))
    end
    all(M .== M[:, 1:1]) && return Arg(ARG_DATA_I, slot!(lw, key * "|I", M[:, 1]))
    return Arg(ARG_DATA_IJ, slot!(lw, key * "|IJ", M; rank=2, dim0=N))
end

function build_gathered_plate!(lw::Lowering, s::Stmt, fam, eta_arg, link, args, lin::Lin, gidx_expr, used::Set{Int})
    n = shape_of(s)[1]
    lvs = loop_vars(s)
    grp = Int.(vec(grid(lw.de, gidx_expr, lvs, s.extents)))
    dmap = Dict{Any,Stmt}()
    for (ref, _) in lin.terms
        ref[1] == :l || continue
        ds = [d for d in defs(lw, ref[2]) if d.kind == :param]
        (length(ds) == 1 && length(ds[1].loops) == 1 && !ds[1].is_gq) || fail("'$(ref[2])' must be defined by one stochastic statement in a single loop")
        ds[1].sid in used && fail("'$(ref[2])' is a parent of more than one observation plate")
        dmap[ref[2]] = ds[1]
    end
    exts = Set(d.extents for d in values(dmap))
    length(exts) == 1 || fail("random effects picked by one data index must share their loop range")
    lo, hi = first(exts)[1]
    (minimum(grp) < lo || maximum(grp) > hi) && fail("data index outside the range of the random effect")
    G = hi - lo + 1
    g0 = grp .- lo                                           # 0-based group of each observation
    counts = zeros(Int, G)
    foreach(g -> counts[g + 1] += 1, g0)
    (n == 0 || minimum(counts) == 0) && fail("every element of a random effect picked by a data index needs at least one observation")
    Tmax = maximum(counts)
    G * Tmax > 8 * n + 1024 && fail("group sizes too unbalanced to pad into a rectangular plate")
    order = sortperm(g0; alg=MergeSort)                      # stable, like numpy's kind="stable"
    start = cumsum(vcat(0, counts[1:(end - 1)]))
    cell = Matrix{Int}(undef, G, Tmax)                       # observation shown in each cell (padding: the group's first)
    for g in 1:G
        cell[g, :] .= order[start[g] + 1]
    end
    w = zeros(G, Tmax)
    fillpos = zeros(Int, G)
    for o in order
        g = g0[o] + 1
        fillpos[g] += 1
        cell[g, fillpos[g]] = o
        w[g, fillpos[g]] = 1.0
    end
    lw.cur_extents = s.extents
    cols(e) = (v = vec(grid(lw.de, e, lvs, s.extents)); [v[cell[:, j]] for j in 1:Tmax])
    key = "gather|$(name_of(s))|$(gidx_expr)"
    plate = Plate(G, Tmax; family=fam, eta_arg=eta_arg, name="$(name_of(s))~$(s.rhs.args[1]) by $(gidx_expr)")
    plate.link = link
    plate.n_real = n
    ycols = cols(s.lhs)
    plate.y = Arg(ARG_DATA_IJ, slot!(lw, key * "|y", hcat(ycols...); rank=2, dim0=G))
    for (q, a) in enumerate(args)
        if q - 1 == eta_arg
            push!(plate.aux, ZERO_ARG)
        elseif is_data(lw, a, lvs)
            push!(plate.aux, columns_arg(lw, "$(key)|aux$(q - 1)", cols(a), G))
        else
            push!(plate.aux, scalar_arg(lw, a, lvs, shape_of(s)))
        end
    end
    (!all(w .== 1.0) && (fam == FAM_T || fam == FAM_BETABIN)) && fail("dt / dbetabin observations in a padded (data-indexed) plate are not supported")
    all(w .== 1.0) || (plate.weight = Arg(ARG_DATA_IJ, slot!(lw, key * "|w", w; rank=2, dim0=G)))
    for (ref, cov) in lin.terms
        ref[1] == :g && push!(plate.terms, Term(Arg(ARG_GVAL, gval_of(lw, ref[2])), columns_arg(lw, "$(key)|cov|$(ref[2])", cols(cov), G)))
    end
    for (ref, cov) in lin.terms
        ref[1] == :l || continue
        d = dmap[ref[2]]
        pfam = family(lw, d.rhs)
        tr, lb, ub = transform_of(lw, pfam, d.rhs.args[2:end])
        tidx = theta_index(lw, d)
        base, si, sj, need_idx = affine_or_index(tidx)
        idx_slot = need_idx ? slot!(lw, "theta_index|$(ref[2])", tidx; rank=ndims(tidx), dim0=size(tidx, 1)) : -1
        lw.cur_extents = d.extents
        pargs = Arg[scalar_arg(lw, a, loop_vars(d), shape_of(d)) for a in d.rhs.args[2:end]]
        lw.cur_extents = s.extents
        push!(plate.locals, Local(string(ref[2]), LEVEL_GROUP, tr, base, si, sj, pfam, pargs, lb, ub, idx_slot))
        push!(used, d.sid)
        push!(plate.terms, Term(Arg(ARG_LOCAL, length(plate.locals) - 1), columns_arg(lw, "$(key)|cov|$(ref[2])", cols(cov), G)))
    end
    if lin.cst !== nothing
        c = columns_arg(lw, "$(key)|offset", cols(lin.cst), G)
        (c.kind == ARG_CONST && c.value == 0.0) || push!(plate.terms, Term(const_arg(1.0), c))
    end
    return plate
end

# ---- sibling observation statements sharing their random effects (lowering.py _siblings / _build_merged_plate) -------------
function local_refs(lw::Lowering, s::Stmt)
    try
        _, _, _, _, eta = peel(lw, s)
        is_call(eta, :inprod) && return Set{Any}()
        return Set{Any}(r[2] for (r, _) in linearize(lw, eta, loop_vars(s)).terms if r[1] == :l)
    catch err
        err isa LoweringError || rethrow()
        return Set{Any}()
    end
end
function siblings(lw::Lowering, s::Stmt, others::Vector{Stmt})
    length(s.loops) == 1 || return Stmt[]
    mine = local_refs(lw, s)
    isempty(mine) && return Stmt[]
    fam, _, _, link, _ = peel(lw, s)
    sibs = Stmt[]
    grew = true
    while grew
        grew = false
        for t in others
            (t in sibs || t.loops != s.loops) && continue
            refs = local_refs(lw, t)
            isempty(intersect(refs, mine)) && continue
            tf, _, _, tl, _ = peel(lw, t)
            (tf == fam && tl == link) || fail("'$(name_of(s))' and '$(name_of(t))' share a random effect but differ in family or link")
            push!(sibs, t)
            union!(mine, refs)
            grew = true
        end
    end
    return sibs
end
function build_merged_plate!(lw::Lowering, stmts::Vector{Stmt}, used::Set{Int})
    s0 = stmts[1]
    lvs = loop_vars(s0)
    lw.cur_extents = s0.extents
    N, K = shape_of(s0)[1], length(stmts)
    fam, eta_arg, _, link, _ = peel(lw, s0)
    fam in (FAM_UNIFORM, FAM_BETA, FAM_FLAT) && fail("$(s0.rhs.args[1]) as an observation family is outside the plate class")
    names = join([string(name_of(t)) for t in stmts], "+")
    plate = Plate(N, K; family=fam, eta_arg=eta_arg, name="$(names)~$(s0.rhs.args[1])")
    plate.link = link
    peeled = [peel(lw, t) for t in stmts]
    evc(e) = vec(grid(lw.de, e, lvs, s0.extents))
    ycols = [evc(t.lhs) for t in stmts]
    plate.y = columns_arg(lw, "merged|y|$(names)", ycols, N)
    if plate.y.kind in (ARG_CONST, ARG_ONE, ARG_DATA_J)
        plate.y = Arg(ARG_DATA_IJ, slot!(lw, "merged|y|$(names)|IJ", hcat(ycols...); rank=2, dim0=N))
    end
    for q in 0:(FAMILY_NARGS[fam] - 1)
        if q == eta_arg
            push!(plate.aux, ZERO_ARG)
            continue
        end
        exprs = [p[3][q + 1] for p in peeled]
        if all(e -> is_data(lw, e, lvs), exprs)
            push!(plate.aux, columns_arg(lw, "merged|aux$(q)|$(names)", [evc(e) for e in exprs], N))
        else
            a0 = scalar_arg(lw, exprs[1], lvs, shape_of(s0))
            for e in exprs[2:end]
                scalar_arg(lw, e, lvs, shape_of(s0)) == a0 || fail("merged observation statements must share their non-data arguments")
            end
            push!(plate.aux, a0)
        end
    end
    lins = [linearize(lw, p[5], lvs) for p in peeled]
    refs = Any[]
    for lin in lins, (r, _) in lin.terms
        r in refs || push!(refs, r)
    end
    zero = zeros(N)
    cov_of(lin, ref) = (q = findfirst(t -> t.first == ref, lin.terms); q === nothing ? nothing : lin.terms[q].second)
    cov_cols(ref) = [(c = cov_of(lin, ref); c === nothing ? zero : evc(c)) for lin in lins]
    for ref in refs
        ref[1] == :g && push!(plate.terms, Term(Arg(ARG_GVAL, gval_of(lw, ref[2])), columns_arg(lw, "merged|cov|$(names)|$(ref[2])", cov_cols(ref), N)))
    end
    for ref in refs
        ref[1] == :g && continue
        owner = stmts[findfirst(lin -> cov_of(lin, ref) !== nothing, lins)]
        li = local_index!(lw, plate, owner, ref, used)
        push!(plate.terms, Term(Arg(ARG_LOCAL, li), columns_arg(lw, "merged|cov|$(names)|$(ref[2])", cov_cols(ref), N)))
    end
    if any(lin -> lin.cst !== nothing, lins)
        c = columns_arg(lw, "merged|offset|$(names)", [lin.cst === nothing ? zero : evc(lin.cst) for lin in lins], N)
        (c.kind == ARG_CONST && c.value == 0.0) || push!(plate.terms, Term(const_arg(1.0), c))
    end
    return plate
end

# ---- dense linear predictor inprod(X[i, 1:P], beta[1:P]) (lowering.py _build_dense) ----------------------------------------------
function build_dense!(lw::Lowering, s::Stmt, fam, plate::Plate, eta, args, used::Set{Int})
    (length(s.loops) == 1 && plate.link == [OP_LOGISTIC] && fam in (FAM_BERNOULLI, FAM_BINOMIAL)) ||
        fail("inprod predictors are lowered for dbern/dbin with a logit link in a single loop")
    lvs = loop_vars(s)
    i_name = lvs[1]
    a, b = eta.args[2], eta.args[3]
    if is_data(lw, b, lvs) && !is_data(lw, a, lvs)
        a, b = b, a
    end
    ok = is_ref(a) && length(a.args) == 3 && a.args[2] === i_name && (is_range(a.args[3]) || is_colon(a.args[3])) && is_data(lw, a, lvs) &&
         is_ref(b) && length(b.args) == 2 && (is_range(b.args[2]) || is_colon(b.args[2]))
    ok || fail("inprod must have the form inprod(X[i, 1:P], beta[1:P]) with X data")
    ds = [d for d in defs(lw, b.args[1]) if d.kind == :param]
    (length(ds) == 1 && length(ds[1].loops) == 1 && !ds[1].is_gq) || fail("'$(b.args[1])' must be defined by one stochastic statement in a single loop")
    d = ds[1]
    d.sid in used && fail("'$(b.args[1])' is a parent of more than one plate")
    Pn, N = shape_of(d)[1], shape_of(s)[1]
    X = Matrix{Float64}(undef, N, Pn)
    for i in 1:N
        X[i, :] .= vec(ev(lw.de, a, Dict{Symbol,Int}(i_name => s.extents[1][1] + i - 1)))
    end
    x_slot = slot!(lw, "dense|$(a)", X; rank=2, dim0=N)
    y_slot = slot!(lw, "dense|y|$(s.lhs)", vec(grid(lw.de, s.lhs, lvs, s.extents)))
    n_slot = fam == FAM_BINOMIAL ? slot!(lw, "dense|n|$(args[2])", vec(grid(lw.de, args[2], lvs, s.extents))) : -1
    base, si, _, need_idx = affine_or_index(theta_index(lw, d))
    need_idx && fail("the coefficient vector of a dense predictor needs an affine theta-slot map")
    pfam = family(lw, d.rhs)
    tr, lb, ub = transform_of(lw, pfam, d.rhs.args[2:end])
    tr == TR_IDENTITY || fail("coefficients of a dense predictor must live on the real line")
    lw.cur_extents = d.extents
    pargs = Arg[scalar_arg(lw, x, loop_vars(d), shape_of(d)) for x in d.rhs.args[2:end]]
    lw.cur_extents = s.extents
    prior = Plate(Pn, 1; family=FAM_FLAT, has_obs=false, name="$(b.args[1])~$(d.rhs.args[1]) (prior)")
    push!(prior.locals, Local(string(b.args[1]), LEVEL_GROUP, tr, base, si, 0, pfam, pargs, lb, ub, -1))
    push!(lw.plan.plates, prior)
    push!(used, d.sid)
    push!(lw.plan.dense, Dense(N, Pn, base, si, x_slot, y_slot, n_slot, fam, OP_LOGISTIC))
end

# ---- the whole plan (lowering.py _build_plan) ---------------------------------------------------------------------------------------
function build_plan!(lw::Lowering)
    lw.plan = Plan(lw.D, lw.transformed, Global[], GOp[], DataSlot[], Plate[], Dense[])
    vec_sids = vector_globals(lw)
    for s in lw.stmts
        (s.trunc !== nothing && !(s.kind == :param && !s.is_gq && (isempty(s.loops) || s.sid in vec_sids))) &&
            fail("'$(name_of(s))': T(,) is lowered for scalar parameters only (not for observations, generated quantities or the parameters of a plate)")
    end
    gitems = Tuple{String,Stmt,Dict{Symbol,Any},Int}[]
    for s in lw.stmts
        (s.kind == :param && !s.is_gq) || continue
        if isempty(s.loops)
            push!(gitems, (instances(lw, s)[1], s, Dict{Symbol,Any}(), 1))
        elseif s.sid in vec_sids
            lvs = loop_vars(s)
            for (k, I) in enumerate(CartesianIndices(Tuple(shape_of(s))))
                sub = Dict{Symbol,Any}(lvs[d] => s.extents[d][1] + I[d] - 1 for d in 1:length(lvs))
                push!(gitems, (instances(lw, s)[k], s, sub, k))
            end
        end
    end
    scalar_logicals = [s for s in lw.stmts if s.kind == :logical && !s.is_gq && isempty(s.loops)]
    for (lab, _, _, _) in gitems
        haskey(lw.gval_ids, lab) && fail("'$(lab)' is defined twice")
        lw.gval_ids[lab] = length(lw.gval_ids)
    end
    for (lab, s, sub, k) in gitems
        fam = family(lw, s.rhs)
        dargs = Any[subst(a, sub) for a in s.rhs.args[2:end]]
        tr, lb, ub = transform_of(lw, fam, dargs)
        s.trunc === nothing || ((tr, lb, ub) = truncation(lw, s, fam, dargs, sub))
        push!(lw.plan.globals, Global(lab, vec(theta_index(lw, s))[k], tr, fam, Arg[], lb, ub))
    end
    for s in scalar_logicals
        emit_gtape!(lw, instances(lw, s)[1], s.rhs)
    end
    for (g, (lab, s, sub, k)) in zip(lw.plan.globals, gitems)
        g.args = Arg[scalar_arg(lw, subst(a, sub), Symbol[], Int[]) for a in s.rhs.args[2:end]]
    end
    used = Set{Int}(vec_sids)
    merged = Set{Int}()
    obs_stmts = [s for s in lw.stmts if s.kind == :obs]
    # observation statements hanging off the same summed-out latent are ONE cell of the sum over its states: the first is
    # the plate's observation, the others are folded into the state's log weight (sibling_logpdf)
    empty!(lw.latent_sibs)
    folded = Set{Int}()
    if lw.marginalize
        by_latent = Dict{Int,Vector{Stmt}}()
        order = Int[]
        for s in obs_stmts
            s.is_gq && continue
            for zs in latents_of(lw, s)
                haskey(by_latent, zs) || (by_latent[zs] = Stmt[]; push!(order, zs))
                push!(by_latent[zs], s)
            end
        end
        for zs in order
            ch = by_latent[zs]
            length(ch) > 1 || continue
            (haskey(lw.latent_sibs, ch[1].sid) || any(t -> t.sid in folded, ch)) &&
                fail("observations shared between two discrete latents are outside the plate class")
            lw.latent_sibs[ch[1].sid] = ch[2:end]
            union!(folded, [t.sid for t in ch[2:end]])
        end
    end
    for s in obs_stmts
        (s.sid in merged || s.sid in folded) && continue
        sibs = siblings(lw, s, [t for t in obs_stmts if !(t.sid in merged) && t.sid != s.sid && !(t.sid in folded)])
        pl = if !isempty(sibs)
            union!(merged, [t.sid for t in sibs])
            build_merged_plate!(lw, vcat([s], sibs), used)
        else
            build_plate!(lw, s, used)
        end
        pl === nothing || push!(lw.plan.plates, pl)
    end
    for s in lw.stmts
        (s.kind == :param && !s.is_gq && !isempty(s.loops) && !(s.sid in used)) &&
            fail("parameter '$(name_of(s))' is not a direct parent of an observation plate (latent-only layers are outside the plate class)")
    end
    return lw.plan
end

"""
    lower_program(program::Expr, data::Dict{Symbol,Any}, node_info; transformed=true) -> (Plan, Lowering)

`program`: the reconstructed model definition (second return value of `JuliaBUGS._generate_lowered_model_def`);
`data`: data and transformed data by symbol (arrays with `missing`/NaN for unobserved cells); `node_info`: per
variable label `(is_stochastic, is_observed, is_generated_quantity)` from `model.graph_evaluation_data`;
`marginalize`: the reference's `UseAutoMarginalization` (evaluation.jl:739-961) -- unobserved `dcat` / `dbern` nodes are
summed out per observation cell instead of being entries of theta (lowering.py, `marginalize=True`).
"""
function lower_program(program::Expr, data::Dict{Symbol,Any}, node_info::Dict{String,Tuple{Bool,Bool,Bool}};
                       transformed::Bool=true, marginalize::Bool=false)
    stmts = flatten!(Stmt[], program, Tuple{Symbol,Any,Any}[])
    lw = Lowering(stmts, DataEval(data, false), transformed, node_info, Dict{Int,Tuple{Int,Vector{Int}}}(), 0,
                  Plan(0, transformed, Global[], GOp[], DataSlot[], Plate[], Dense[]), Dict{String,Int}(), Dict{String,Int}(),
                  Tuple{Int,Int}[], marginalize, nothing, Dict{Int,Vector{Stmt}}())
    bounds!(lw)
    classify!(lw)
    order_parameters!(lw)
    build_plan!(lw)
    return lw.plan, lw
end

end # module JBugsLowering
