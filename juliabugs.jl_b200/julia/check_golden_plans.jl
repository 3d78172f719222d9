# check_golden_plans.jl -- run on a machine that has Julia + JuliaBUGS (not possible in the build image):
#
#     julia --project=<JuliaBUGS checkout> juliabugs.jl_b200/julia/check_golden_plans.jl
#
# For every example the reference ships whose plan is committed under tests/golden/plans/, compile the model with the
# reference, lower it with the Julia pass (lower_to_plan.jl through JuliaBUGSCUDABatch.lower_to_plan) and compare
#   * the plan blob, byte for byte, with tests/golden/plans/<name>.bin (written by the Python pass,
#     tests/golden/make_plans.py), and
#   * every data slot with the sha256 digests of tests/golden/plans/manifest.json,
#   * the theta labels (`getparams` order after `set_evaluation_mode(model, UseCUDABatch())`) with the manifest's.
# No GPU is needed: nothing here calls libjbugs_cuda.
using JuliaBUGS, SHA, JSON
include(joinpath(@__DIR__, "JuliaBUGSCUDABatch.jl"))
using .JuliaBUGSCUDABatch: lower_to_plan, JBugsLowering

const GOLD = normpath(joinpath(@__DIR__, "..", "..", "tests", "golden", "plans"))
const MANIFEST = JSON.parsefile(joinpath(GOLD, "manifest.json"))
# golden name => field of JuliaBUGS.BUGSExamples (surgical: the two model definitions of 05_Surgical.jl)
const EXAMPLES = Dict("rats" => :rats, "pumps" => :pumps, "dogs" => :dogs, "seeds" => :seeds,
                      "surgical_realistic" => :surgical_realistic, "surgical_simple" => :surgical_simple, "epil" => :epil,
                      "blockers" => :blockers, "oxford" => :oxford, "salm" => :salm, "equiv" => :equiv, "dyes" => :dyes,
                      "stacks" => :stacks, "bones" => :bones, "dugongs" => :dugongs, "lsat" => :lsat, "air" => :air,
                      "leuk" => :leuk, "leukfr" => :leukfr)

# plans lowered with `marginalize=true` (the reference's UseAutoMarginalization): Cervix sums its 1929 missing x[i] out
const MARGINALISED = Dict("cervix_marginalised" => :cervix)

function check(name, field; marginalize=false)
    ex = getfield(JuliaBUGS.BUGSExamples, field)
    model = compile(ex.model_def, ex.data, ex.inits)
    plan, _ = lower_to_plan(model; marginalize=marginalize)
    blob = JBugsLowering.to_bytes(plan)
    gold = read(joinpath(GOLD, name * ".bin"))
    want = MANIFEST[name]
    ok = blob == gold && bytes2hex(sha256(gold)) == want["sha256"]
    for (sl, w) in zip(plan.data, want["data_slots"])
        ok &= bytes2hex(sha256(reinterpret(UInt8, sl.values))) == w["sha256"]
    end
    ok &= length(plan.data) == length(want["data_slots"])
    println(rpad(name, 22), ok ? "identical" : "DIFFERS", "  (", length(blob), " bytes, ", length(plan.data), " data slots)")
    return ok
end

results = [check(n, f) for (n, f) in sort(collect(EXAMPLES))]
append!(results, [check(n, f; marginalize=true) for (n, f) in sort(collect(MARGINALISED))])
println(count(results), " of ", length(results), " plans byte-identical to the Python pass")
exit(all(results) ? 0 : 1)
