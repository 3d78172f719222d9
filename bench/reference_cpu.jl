# bench/reference_cpu.jl -- the reference's own CPU timing of the hot path, for machines that have Julia + JuliaBUGS.
#
# Julia is NOT installed in the build image nor on the GPU box, so this script ships unexecuted; `bench.py --impl reference`
# therefore times the C restatement (oracle/jbugs_oracle.c) and says so in its JSON line.  On a machine with Julia:
#
#     julia --project=/path/to/JuliaBUGS -t auto bench/reference_cpu.jl rats 30
#
# It follows the reference's benchmark harness (JuliaBUGS/benchmark/juliabugs.jl:76-87): generated log-density mode,
# `BUGSModelWithGradient` with Mooncake, Chairmarks medians, and reports the metric of bench.py
# (obs-evals/s = observed scalar nodes x chains / time of one logdensity_and_gradient; one chain per call).
using JuliaBUGS, LogDensityProblems, Chairmarks, ADTypes, Mooncake, Random, Printf

function rats_data(N; T=5, seed=1234)
    rng = Random.MersenneTwister(seed)
    x = [8.0, 15.0, 22.0, 29.0, 36.0][1:T]
    alpha = 242 .+ 14 .* randn(rng, N)
    beta = 6.2 .+ 0.5 .* randn(rng, N)
    Y = [alpha[i] + beta[i] * (x[j] - 22) + 6 * randn(rng) for i in 1:N, j in 1:T]
    return (; x, xbar=22, N, T, Y)
end

function main(args)
    name = Symbol(get(args, 1, "rats"))
    ex = getfield(JuliaBUGS.BUGSExamples, name)
    data = name == :rats && length(args) >= 2 ? rats_data(parse(Int, args[2])) : ex.data
    model = compile(ex.model_def, data)            # the unrolled graph compile limits N (SURVEY 3.1)
    model = JuliaBUGS.set_evaluation_mode(model, JuliaBUGS.UseGeneratedLogDensityFunction())
    ad_model = JuliaBUGS.BUGSModelWithGradient(model, AutoMooncake(; config=nothing))
    theta = JuliaBUGS.getparams(model)
    n_obs = count(vn -> model.g[vn].is_observed, JuliaBUGS.Model.labels(model.g))
    t = Chairmarks.@be LogDensityProblems.logdensity_and_gradient($ad_model, $theta)
    sec = Chairmarks.median(t).time
    @printf("{\"impl\": \"reference\", \"metric\": \"logdensity+gradient obs-evals/s\", \"value\": %.6e, \"unit\": \"obs-evals/s\", \"ms_per_step\": %.6f, \"cores\": 1, \"kind\": \"reference\", \"sample\": \"%s, one chain per call, Mooncake\"}\n",
            n_obs / sec, sec * 1e3, name)
end

main(ARGS)
