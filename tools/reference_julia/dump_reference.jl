# dump_reference.jl -- run the REAL LegendDSP.dsp_icpc on an exported parity case (SURVEY.md section 8c: "substitute truth").
#
#   python tools/reference_julia/export_case.py case_dir 2000
#   julia --project=<env with LegendDSP> tools/reference_julia/dump_reference.jl case_dir
#   python tools/reference_julia/compare_dump.py case_dir
#
# NOT RUN in this repository's image (no Julia toolchain); it only uses the reference's public API:
# DSPConfig(::PropDict) (src/types.jl), dsp_icpc(data, config, τ, pars_filter) (src/dsp_icpc.jl:62-230).
# Output: case_dir/reference_rows.bin, n_events x 49 Float64, row = event, columns in the order of case.json["columns"],
# every quantity stripped in the unit case.json["units"] names (dimensionless otherwise).
using LegendDSP, RadiationDetectorSignals, Unitful, TypedTables, PropDicts, IntervalSets, ArraysOfArrays
import JSON

dir = ARGS[1]
case = JSON.parsefile(joinpath(dir, "case.json"))
n_ev, n = case["n_events"], case["n_samples"]
const UNIT = Dict("ns" => u"ns", "us" => u"µs", "ms" => u"ms", "s" => u"s")
T(q::AbstractDict) = Float64(q["val"]) * UNIT[q["unit"]]    # quantities keep the unit they were written in (rounding ties)
T(x::Real) = Float64(x) * u"ns"

# ---- config: plain numbers -> the Unitful PropDict DSPConfig expects ----
c = case["config"]
win(w) = PropDict(:min => T(w["min"]), :max => T(w["max"]))
grid(g) = PropDict(:start => T(g["start"]), :stop => T(g["stop"]), :step => T(g["step"]))
rtft(g) = PropDict(:rt => grid(g["rt"]), :ft => grid(g["ft"]))
fd = c["flt_defaults"]; kw = c["kwargs_pars"]
pd = PropDict(
    :enc_pickoff_trap => T(c["enc_pickoff_trap"]), :enc_pickoff_zac => T(c["enc_pickoff_zac"]),
    :enc_pickoff_cusp => T(c["enc_pickoff_cusp"]),
    :bl_window => win(c["bl_window"]), :tail_window => win(c["tail_window"]), :current_window => win(c["current_window"]),
    :auxbl1_window => win(c["auxbl1_window"]), :auxbl2_window => win(c["auxbl2_window"]),
    :auxpz1_window => win(c["auxpz1_window"]), :auxpz2_window => win(c["auxpz2_window"]),
    :flt_length_cusp => T(c["flt_length_cusp"]), :flt_length_zac => T(c["flt_length_zac"]),
    :t0_threshold => Float64(c["t0_threshold"]), :inTraceCut_std_threshold => Float64(c["inTraceCut_std_threshold"]),
    :sg_flt_degree => Int(c["sg_flt_degree"]),
    :qdrift_int_length => [T(q) for q in c["qdrift_int_length"]], :lq_int_length => [T(q) for q in c["lq_int_length"]],
    :e_grid_trap => rtft(c["e_grid_trap"]), :e_grid_zac => rtft(c["e_grid_zac"]), :e_grid_cusp => rtft(c["e_grid_cusp"]),
    :a_grid_wl_sg => grid(c["a_grid_wl_sg"]),
    :flt_defaults => PropDict(
        :sg => T(fd["sg"]),
        :trap => PropDict(:rt => T(fd["trap"]["rt"]), :ft => T(fd["trap"]["ft"])),
        :zac => PropDict(:rt => T(fd["zac"]["rt"]), :ft => T(fd["zac"]["ft"])),
        :cusp => PropDict(:rt => T(fd["cusp"]["rt"]), :ft => T(fd["cusp"]["ft"]))),
    :kwargs_pars => PropDict(
        :fc_bit_depth => Int(kw["fc_bit_depth"]),
        :t0_flt_pars => [T(q) for q in kw["t0_flt_pars"]],
        :t0_mintot => T(kw["t0_mintot"]), :tx_mintot => T(kw["tx_mintot"]), :intrace_mintot => T(kw["intrace_mintot"]),
        :int_interpolation_order => Int(kw["int_interpolation_order"]),
        :int_interpolation_length => T(kw["int_interpolation_length"]),
        :sig_interpolation_order => Int(kw["sig_interpolation_order"]),
        :sig_interpolation_length => T(kw["sig_interpolation_length"])),
)
# NOTE: if this DSPConfig constructor expects a different nesting for a key (it is defined by _create_dsp_config in the
# installed LegendDSP version), adapt the PropDict above; the VALUES are the contract of this dump.
config = DSPConfig(pd)

# ---- data: dense n_samples x n_events UInt16 matrix -> ArrayOfRDWaveforms (each waveform contiguous) ----
raw = Array{UInt16}(undef, n, n_ev)
read!(joinpath(dir, "wf_u16.bin"), raw)
t = range(0.0u"ns", step = T(case["step_ns"]), length = n)
wvfs = ArrayOfRDWaveforms((fill(t, n_ev), nestedview(raw)))
data = Table(waveform = wvfs, baseline = fill(0.0f0, n_ev), timestamp = fill(UInt64(0), n_ev),
             eventnumber = UInt32.(1:n_ev), daqenergy = fill(UInt16(0), n_ev))

result = dsp_icpc(data, config, T(case["tau_ns"]), PropDict())

units = case["units"]
unit_of = Dict("us" => u"µs", "ns" => u"ns", "1/ns" => u"ns^-1")
rows = Array{Float64}(undef, length(case["columns"]), n_ev)      # column-major: one event per column = row-major file
for (j, name) in enumerate(case["columns"])
    col = getproperty(result, Symbol(name))
    for e in 1:n_ev
        v = col[e]
        rows[j, e] = v isa Quantity ? Float64(ustrip(unit_of[units[name]], v)) : Float64(v)
    end
end
write(joinpath(dir, "reference_rows.bin"), rows)
println("wrote ", joinpath(dir, "reference_rows.bin"), " (", n_ev, " events x ", size(rows, 1), " columns)")
