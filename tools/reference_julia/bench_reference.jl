# bench_reference.jl -- throughput of the REAL LegendDSP.dsp_icpc on the host cores (SURVEY.md section 8d, "CPU baseline").
#
#   julia -t auto --project=<env with LegendDSP> tools/reference_julia/bench_reference.jl case_dir [repeats]
#
# NOT RUN in this repository's image (no Julia toolchain).  Uses the case exported by export_case.py; events are split
# into one contiguous chunk per Julia thread (the reference itself is single-threaded per call, src/dsp_icpc.jl:62-230).
include(joinpath(@__DIR__, "dump_reference.jl"))     # builds `data`, `config`, runs once (compilation excluded below)

repeats = length(ARGS) > 1 ? parse(Int, ARGS[2]) : 3
nt = Threads.nthreads()
chunks = [r for r in Iterators.partition(1:n_ev, cld(n_ev, nt))]
τ = T(case["tau_ns"])
best = Inf
for _ in 1:repeats
    t0 = time_ns()
    Threads.@threads for r in chunks
        dsp_icpc(data[r], config, τ, PropDict())
    end
    global best = min(best, (time_ns() - t0) / 1e9)
end
println("{\"impl\": \"LegendDSP.jl dsp_icpc\", \"threads\": ", nt, ", \"events\": ", n_ev, ", \"seconds\": ", best,
        ", \"waveforms_per_s\": ", n_ev / best, "}")
