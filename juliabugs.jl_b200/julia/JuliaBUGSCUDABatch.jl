# JuliaBUGSCUDABatch.jl -- the Julia side of the drop-in: what a JuliaBUGS maintainer adds (as a package
# extension or `src/model/cuda_batch.jl`) to bind libjbugs_cuda.  SHIPPED UNEXECUTED: Julia is not installed
# in the build image; the same ABI is exercised from Python (`jbugs/cuda.py`) by the test-suite.
#
# Reference plug-in points (paths relative to JuliaBUGS/):
#   EvaluationMode subtypes            src/model/bugsmodel.jl:248-252
#   set_evaluation_mode                src/model/bugsmodel.jl:736-801
#   _eval_logdensity dispatch          src/model/logdensityproblems.jl:3-17
#   LogDensityProblems methods         src/model/logdensityproblems.jl:19-55, 168-174, 219-232
#   cache invalidation on condition    src/model/abstractppl.jl:796-799
#   set_observed_values!               src/model/abstractppl.jl:872-888
module JuliaBUGSCUDABatch

using JuliaBUGS
using JuliaBUGS: BUGSModel
using JuliaBUGS.Model: EvaluationMode
using LogDensityProblems

const libjbugs = get(ENV, "JBUGS_CUDA_LIB", "libjbugs_cuda")

# ---------------------------------------------------------------------------------------------- ABI (include/jbugs_cuda.h)
const LAYOUT_CHAIN_FAST = Cint(0)   # element (d, c) at d*ld + c : a Julia Matrix{Float64}(C, D)  (ld = C)
const LAYOUT_PARAM_FAST = Cint(1)   # element (d, c) at c*ld + d : a Julia Matrix{Float64}(D, C)  (ld = D)

struct JBugsError <: Exception
    code::Cint
    msg::String
end
_check(rc) = rc == 0 ? nothing : throw(JBugsError(rc, unsafe_string(ccall((:jbugs_last_error, libjbugs), Cstring, ()))))

mutable struct PlanHandle
    ptr::Ptr{Cvoid}
    function PlanHandle(blob::Vector{UInt8}, device::Integer)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        _check(ccall((:jbugs_plan_create, libjbugs), Cint, (Ptr{UInt8}, Csize_t, Cint, Ref{Ptr{Cvoid}}),
                     blob, length(blob), device, out))
        h = new(out[])
        finalizer(h -> (h.ptr == C_NULL || ccall((:jbugs_plan_destroy, libjbugs), Cint, (Ptr{Cvoid},), h.ptr); h.ptr = C_NULL), h)
        return h
    end
end

dim(h::PlanHandle) = Int(ccall((:jbugs_dim, libjbugs), Int64, (Ptr{Cvoid},), h.ptr))

function upload!(h::PlanHandle, slot::Integer, values::Vector{Float64})
    _check(ccall((:jbugs_plan_upload_data, libjbugs), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Csize_t, Cint),
                 h.ptr, slot, values, sizeof(values), 0))
end
function upload!(h::PlanHandle, slot::Integer, values::Vector{Int64})
    _check(ccall((:jbugs_plan_upload_data, libjbugs), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Csize_t, Cint),
                 h.ptr, slot, values, sizeof(values), 1))
end

"""
    eval_batch!(h, θ::Matrix{Float64}, logp, grad, status; want_grad=true)

`θ` is `D × C` (one column per chain, Julia's native layout == PARAM_FAST).  Host arrays are staged by the
library; pass `CuArray` pointers (CHAIN_FAST, see `batch_alloc`) to evaluate in place on the device.
"""
function eval_batch!(h::PlanHandle, θ::Matrix{Float64}, logp::Vector{Float64}, grad::Matrix{Float64},
                     status::Vector{Int32}; want_grad::Bool=true)
    D, C = size(θ)
    _check(ccall((:jbugs_eval_batch, libjbugs), Cint,
                 (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Cint, Cint, Ptr{Cvoid}),
                 h.ptr, θ, D, C, LAYOUT_PARAM_FAST, logp, grad, status, want_grad, 0, C_NULL))
    return logp, grad, status
end

# ---------------------------------------------------------------------------------------------- the evaluation mode
include("lower_to_plan.jl")
using .JBugsLowering

"""
    UseCUDABatch(; device=0, marginalize=false)

New `EvaluationMode`: the model is lowered at statement/plate granularity to a flat kernel plan
(`include/jbugs_plan.h`) and evaluated by libjbugs_cuda for a whole batch of chains per launch.

The device state lives IN the mode object (a mutable struct): `BUGSModel` is an immutable struct, so it cannot key a
`WeakKeyDict`, and the reference resets `evaluation_mode` to `UseGraph()` whenever `condition` / `fix` / `decondition`
change the graph (src/model/abstractppl.jl:789-797, next to `log_density_computation_function => nothing`) -- the old
mode object, and with it the plan handle, is dropped exactly where the generated function is.
"""
mutable struct UseCUDABatch <: EvaluationMode
    device::Int
    marginalize::Bool      # the reference's UseAutoMarginalization on the device: dcat / dbern latents summed out per cell
    handle::Union{Nothing,PlanHandle}
    plan::Union{Nothing,JBugsLowering.Plan}
end
UseCUDABatch(; device=0, marginalize=false) = UseCUDABatch(device, marginalize, nothing, nothing)

_label(vn) = replace(string(vn), " " => "")          # "Y[1, 2]" -> "Y[1,2]", the label format of the lowering

"""
    lower_to_plan(model::BUGSModel) -> (plan::JBugsLowering.Plan, reconstructed_model_def::Expr)

Statement-level lowering on top of `_generate_lowered_model_def` (src/source_gen.jl:459-552): the reference fissions,
sorts and regroups the statements and removes transformed data; `JBugsLowering.lower_program` (lower_to_plan.jl) turns
the reconstructed program into plate records.  Observed / parameter / generated-quantity classes come from
`model.graph_evaluation_data` (src/model/bugsmodel.jl:118-246), data from `model.data` and the transformed data the
reference folded into `model.evaluation_env` (src/JuliaBUGS.jl:175-187).
"""
function lower_to_plan(model::BUGSModel; marginalize::Bool=false)
    gd = model.graph_evaluation_data
    lowered, reconstructed = JuliaBUGS._generate_lowered_model_def(
        model.model_def, model.g, model.evaluation_env;
        generated_quantities=Set(gd.generated_quantities), fixed_parameters=Set(gd.fixed_parameters))
    lowered === nothing && throw(JBugsLowering.LoweringError("the reference could not order the statements of this model sequentially"))
    gq = Set(_label(vn) for vn in gd.generated_quantities)
    node_info = Dict{String,Tuple{Bool,Bool,Bool}}()
    for (k, vn) in enumerate(gd.sorted_nodes)
        lab = _label(vn)
        node_info[lab] = (gd.is_stochastic_vals[k], gd.is_observed_vals[k], lab in gq)
    end
    # data = what the user passed (missing -> NaN) + the transformed data the reference computed: every symbol of the
    # evaluation environment that no statement of the reconstructed program defines
    defined = Set{Symbol}()
    function lhs_symbols(block)
        for st in block.args
            st isa Expr || continue
            if st.head == :for
                lhs_symbols(st.args[2])
            elseif st.head == :block
                lhs_symbols(st)
            elseif st.head == :(=) || (st.head == :call && st.args[1] == :~)
                l = st.head == :(=) ? st.args[1] : st.args[2]
                push!(defined, l isa Symbol ? l : l.args[1])
            end
        end
    end
    lhs_symbols(reconstructed)
    tofloat(v) = v isa AbstractArray ? map(x -> ismissing(x) ? NaN : Float64(x), v) : (ismissing(v) ? NaN : Float64(v))
    data = Dict{Symbol,Any}()
    for (k, v) in pairs(model.evaluation_env)
        k in defined || (data[k] = tofloat(v))
    end
    for (k, v) in pairs(model.data)
        data[k] = tofloat(v)
    end
    plan, _ = JBugsLowering.lower_program(reconstructed, data, node_info; transformed=model.transformed, marginalize=marginalize)
    return plan, reconstructed
end

function plan_handle(model::BUGSModel)
    mode = model.evaluation_mode::UseCUDABatch
    if mode.handle === nothing
        plan, _ = lower_to_plan(model; marginalize=mode.marginalize)
        h = PlanHandle(JBugsLowering.to_bytes(plan), mode.device)
        foreach(((k, sl),) -> upload!(h, k - 1, sl.values), enumerate(plan.data))
        mode.handle, mode.plan = h, plan
    end
    return mode.handle
end

# set_evaluation_mode(model, UseCUDABatch()): the parameter order is rebuilt from the reconstructed program exactly as the
# generated mode does (src/model/bugsmodel.jl:770-799) so that getparams / dimension / chain naming follow the plan's
# theta layout.  Works in both spaces: `settrans(model, false)` gives identity transforms in the plan.
function JuliaBUGS.set_evaluation_mode(model::BUGSModel, mode::UseCUDABatch)
    _, reconstructed = lower_to_plan(model; marginalize=mode.marginalize)
    pass = JuliaBUGS.CollectSortedNodes(model.evaluation_env)
    JuliaBUGS.analyze_block(pass, reconstructed)
    gd = model.graph_evaluation_data
    sorted_nodes = filter(node -> node in gd.sorted_nodes, pass.sorted_nodes)
    new_gd = JuliaBUGS.Model.GraphEvaluationData(model.g, sorted_nodes;
                                                 generated_quantities=Set(gd.generated_quantities),
                                                 fixed_parameters=Set(gd.fixed_parameters))
    model = JuliaBUGS.BangBang.setproperty!!(model, :graph_evaluation_data, new_gd)
    if mode.marginalize
        # the reference's own marginalising mode first (bugsmodel.jl:802-873): it fills `marginalization_cache` for the new
        # node order -- dimension / getparams then count the continuous parameters only, in the plan's theta order
        model = JuliaBUGS.set_evaluation_mode(model, JuliaBUGS.UseAutoMarginalization())
        model.evaluation_mode isa JuliaBUGS.UseAutoMarginalization ||
            throw(JBugsLowering.LoweringError("the reference could not prepare auto-marginalisation for this model"))
    end
    # (the same call the reference ends its own set_evaluation_mode with, bugsmodel.jl:876)
    return JuliaBUGS.BangBang.setproperty!!(model, :evaluation_mode, UseCUDABatch(mode.device, mode.marginalize, nothing, nothing))
end

# set_observed_values! keeps the structure (abstractppl.jl:872-888): re-lower for the new values, re-upload, do not re-plan
function reupload!(model::BUGSModel)
    h = plan_handle(model)
    plan, _ = lower_to_plan(model; marginalize=model.evaluation_mode.marginalize)
    JBugsLowering.to_bytes(plan) == JBugsLowering.to_bytes(model.evaluation_mode.plan) ||
        error("the new observations change the structure of the plan (e.g. the missingness pattern): set the evaluation mode again")
    foreach(((k, sl),) -> upload!(h, k - 1, sl.values), enumerate(plan.data))
    model.evaluation_mode.plan = plan
    return model
end

# ---------------------------------------------------------------------------------------------- LogDensityProblems
function JuliaBUGS.Model._eval_logdensity(model::BUGSModel, ::UseCUDABatch, x::AbstractVector)
    logp = Vector{Float64}(undef, 1)
    eval_batch!(plan_handle(model), reshape(collect(Float64, x), :, 1), logp, Matrix{Float64}(undef, 0, 0),
                Vector{Int32}(undef, 1); want_grad=false)
    return logp[1]          # -Inf on DomainError, exactly like logdensityproblems.jl:22-26
end

# the CUDA mode carries its own reverse pass: first-order capability without an AD backend
struct CUDABatchModel{M<:BUGSModel}
    base_model::M
end
LogDensityProblems.capabilities(::Type{<:CUDABatchModel}) = LogDensityProblems.LogDensityOrder{1}()
LogDensityProblems.dimension(m::CUDABatchModel) = LogDensityProblems.dimension(m.base_model)
LogDensityProblems.logdensity(m::CUDABatchModel, x::AbstractVector) = LogDensityProblems.logdensity(m.base_model, x)

function LogDensityProblems.logdensity_and_gradient(m::CUDABatchModel, x::AbstractVector)
    θ = reshape(collect(Float64, x), :, 1)
    logp, grad, _ = eval_batch!(plan_handle(m.base_model), θ, Vector{Float64}(undef, 1), similar(θ), Vector{Int32}(undef, 1))
    return logp[1], vec(grad)            # fresh copy; (-Inf, NaN...) on DomainError (logdensityproblems.jl:226-229)
end

"""
    logdensity_and_gradient_batch(m, θ::AbstractMatrix) -> (logp::Vector, grad::Matrix)

The new batched entry point (the reference has none: one θ per call, logdensityproblems.jl:19,219):
`θ` is `D × C`, one column per chain.
"""
function logdensity_and_gradient_batch(m::CUDABatchModel, θ::AbstractMatrix)
    θ = Matrix{Float64}(θ)
    C = size(θ, 2)
    logp, grad, status = eval_batch!(plan_handle(m.base_model), θ, Vector{Float64}(undef, C), similar(θ), Vector{Int32}(undef, C))
    return logp, grad
end

# --- run-time specialisation: a kernel compiled for this model's plate shapes (nvcc, cached by shape); the reference
# generates and compiles Julia code per model at the same point (set_evaluation_mode, bugsmodel.jl:736-801)
specialize!(h::PlanHandle; cache_dir::Union{Nothing,String}=nothing) =
    _check(ccall((:jbugs_plan_specialize, libjbugs), Cint, (Ptr{Cvoid}, Cstring), h.ptr, cache_dir === nothing ? C_NULL : cache_dir))

# --- observation sharding (one process per GPU; the unique id travels over MPI / Distributed, not shown) -------------
set_shard!(h::PlanHandle, plate::Integer, i0::Integer, i1::Integer) =
    _check(ccall((:jbugs_plan_set_shard, libjbugs), Cint, (Ptr{Cvoid}, Cint, Int64, Int64), h.ptr, plate, i0, i1))
set_dense_shard!(h::PlanHandle, i0::Integer, i1::Integer) =
    _check(ccall((:jbugs_plan_set_dense_shard, libjbugs), Cint, (Ptr{Cvoid}, Int64, Int64), h.ptr, i0, i1))
function comm_init!(h::PlanHandle, nranks::Integer, rank::Integer, unique_id::Vector{UInt8})
    @assert length(unique_id) == 128
    _check(ccall((:jbugs_comm_init, libjbugs), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), h.ptr, nranks, rank, unique_id))
end

# --- the sampler on the device: what `AbstractMCMC.sample(model, NUTS(0.8), n)` does through AdvancedHMC
# (ext/JuliaBUGSAdvancedHMCExt.jl:40-66), for `n_chains` lock-step chains; draws come back as (chains, monitored, draws)
function sample_nuts(model::BUGSModel, n_samples::Integer; n_adapts::Integer=1000, n_chains::Integer=128,
                     max_depth::Integer=10, seed::Integer=1234, monitor::Vector{Int64}=collect(Int64, 0:LogDensityProblems.dimension(model) - 1),
                     step_size::Real=0.01, jitter::Real=0.1)
    h = plan_handle(model)
    hmc = Ref{Ptr{Cvoid}}(C_NULL)
    _check(ccall((:jbugs_hmc_create, libjbugs), Cint, (Ptr{Cvoid}, Int64, UInt64, Ref{Ptr{Cvoid}}), h.ptr, n_chains, seed, hmc))
    try
        θ0 = Vector{Float64}(JuliaBUGS.getparams(model))
        _check(ccall((:jbugs_hmc_init, libjbugs), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cdouble, Cdouble), hmc[], θ0, jitter, step_size))
        acc = Ref{Cdouble}(0.0)
        _check(ccall((:jbugs_hmc_run_nuts, libjbugs), Cint,
                     (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int64}, Cint, Ptr{Float64}, Ref{Cdouble}),
                     hmc[], n_adapts, max_depth, 1, C_NULL, 0, C_NULL, acc))
        draws = Array{Float64}(undef, n_chains, length(monitor), n_samples)   # unconstrained theta rows
        _check(ccall((:jbugs_hmc_run_nuts, libjbugs), Cint,
                     (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Int64}, Cint, Ptr{Float64}, Ref{Cdouble}),
                     hmc[], n_samples, max_depth, 0, monitor, length(monitor), draws, acc))
        return draws, acc[]
    finally
        ccall((:jbugs_hmc_destroy, libjbugs), Cint, (Ptr{Cvoid},), hmc[])
    end
end

end # module
