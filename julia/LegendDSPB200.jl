# LegendDSPB200.jl -- drop-in for LegendDSP.dsp_icpc on top of liblgdsp_b200.so (include/lgdsp_b200.h).
#
# UNTESTED in this repository's environment: the build image has no Julia toolchain.  The struct layout below is
# the one tests/test_abi.py verifies for the Python ctypes mirror (same offsets); the index arithmetic uses the
# reference's own Unitful expressions (src/tailstats.jl:16-18, src/dsp_routines.jl:9-25, src/dsp_icpc.jl:87-99).
#
# Usage:   using LegendDSP, LegendDSPB200
#          tbl = LegendDSPB200.dsp_icpc(data, config, τ, pars_filter)      # same Table as LegendDSP.dsp_icpc
module LegendDSPB200

using Unitful, TypedTables, PropDicts, ArraysOfArrays, RadiationDetectorSignals
import LegendDSP: DSPConfig, get_fltpars

const LIB = get(ENV, "LGDSP_B200_LIB", "liblgdsp_b200.so")

const MAX_DNI, MAX_SG, MAX_FIR, NCOL = 64, 33, 4096, 49
const PARAMS_VERSION = UInt32(4)      # LGDSP_PARAMS_VERSION

struct Trap; navg::Int32; ngap::Int32; navg2::Int32; reserved::Int32; end
struct Dni;  n_w::Int32; degree::Int32; A::NTuple{MAX_DNI * 4, Float64}; end
struct Sg;   n_taps::Int32; offset::Int32; h::NTuple{MAX_SG, Float64}; end
struct CuspZac
    n_taps::Int32; flat::Int32; sigma::Float64; tau::Float64; beta::Float64; coeffs::NTuple{MAX_FIR, Float64}
end
struct IcpcParams          # == lgdsp_icpc_params, include/lgdsp_b200.h
    struct_size::UInt32; version::UInt32; n_samples::Int32; groups::UInt32
    t_first_ns::Float64; dt_ns::Float64
    sat_low::Int64; sat_high::Int64
    bl_from::Int32; bl_until::Int32; tail_from::Int32; tail_until::Int32
    pz_km1::Float64
    t0_trap::Trap; t0inv_trap::Trap; t0_threshold::Float64; t0_min_n::Int32; tx_min_n::Int32
    tx_frac::NTuple{5, Float64}
    qdrift_first_ns::Float64; qdrift_last_ns::Float64; lq_first_ns::Float64; lq_last_ns::Float64
    int_dni::Dni; sig_dni::Dni
    trap_10410::Trap; trap_535::Trap; trap_313::Trap; trap_e::Trap
    trap_pickoff_ns::Float64; cusp_pickoff_ns::Float64; zac_pickoff_ns::Float64
    sg::NTuple{3, Sg}
    cur_from::NTuple{4, Int32}; cur_until::NTuple{4, Int32}
    intrace_nsigma::Float64; intrace_min_n::Int32; intrace_bl_from::Int32; intrace_bl_until::Int32
    cuspzac_direct::Int32; reserved0::Int32
    cusp::CuspZac; zac::CuspZac
end

mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(device::Integer = 0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:lgdsp_create, LIB), Cint, (Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), device, C_NULL, out)
        rc == 0 || error("lgdsp_create: " * unsafe_string(ccall((:lgdsp_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        h = new(out[])
        finalizer(x -> ccall((:lgdsp_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.ptr), h)
        h
    end
end
lasterr(h::Handle) = unsafe_string(ccall((:lgdsp_last_error, LIB), Cstring, (Ptr{Cvoid},), h.ptr))

# the reference's own rounding: round(Int, ...) is ties-to-even on the Unitful quotient
_idx(t, first_x, step_x) = round(Int, ustrip(NoUnits, (t - first_x) / step_x))           # 0-based (src/tailstats.jl:16-18)
_cnt(t, step) = round(Int, ustrip(NoUnits, t / step))
_trap(a, g, a2, dt) = Trap(_cnt(a, dt), _cnt(g, dt), _cnt(a2, dt), 0)

function _dni(order::Integer, len, dt)
    n_w = _cnt(len, dt)
    A = zeros(Float64, MAX_DNI * 4)
    ccall((:lgdsp_lsq_fit_matrix, LIB), Cint, (Int32, Int32, Ptr{Float64}), n_w, order, A) == 0 || error("lsq_fit_matrix")
    Dni(n_w, order, Tuple(A))
end
function _sg(len, degree, dt)
    n = _cnt(len, dt); n += iseven(n)                      # RddspPolicy.sg_even_length = "up" (DESIGN.md section 2)
    h = zeros(Float64, MAX_SG)
    ccall((:lgdsp_sg_coeffs, LIB), Cint, (Int32, Int32, Int32, Ptr{Float64}), n, degree, 1, h) == 0 || error("sg_coeffs")
    Sg(n, (n - 1) ÷ 2, Tuple(h))
end
function _cz(sym::Symbol, rt, ft, τcz, len, dt)
    L = _cnt(len, dt); c = zeros(Float64, MAX_FIR)
    σ, τs, β = ustrip(NoUnits, rt / dt), ustrip(NoUnits, τcz / dt), ustrip(NoUnits, len / dt)   # src/dsp_icpc.jl:88,90
    ccall((sym, LIB), Cint, (Float64, Int32, Float64, Int32, Float64, Ptr{Float64}), σ, _cnt(ft, dt), τs, L, β, c) == 0 || error(string(sym))
    CuspZac(L, _cnt(ft, dt), σ, τs, β, Tuple(c))
end

"""resolve DSPConfig + τ + pars_filter into sample-domain parameters (mirrors legenddsp.jl_b200/config.py)"""
function resolve_params(wvfs, config::DSPConfig, τ, pars_filter::PropDict)
    t = wvfs[1].time; dt = step(t); t1 = first(t); n = length(t)
    kw = config.kwargs_pars
    trap_rt, trap_ft = get_fltpars(pars_filter, :trap, config)
    cusp_rt, cusp_ft = get_fltpars(pars_filter, :cusp, config)
    zac_rt, zac_ft = get_fltpars(pars_filter, :zac, config)
    sg_wl = get_fltpars(pars_filter, :sg, config)
    t0p = get(kw, :t0_flt_pars, [40u"ns", 100u"ns", 2000u"ns"])
    sg0 = _sg(sg_wl, config.sg_flt_degree, dt); sg1 = _sg(60u"ns", config.sg_flt_degree, dt); sg2 = _sg(100u"ns", config.sg_flt_degree, dt)
    curw(s, k) = (max(0, _idx(leftendpoint(config.current_window), t1 + s.offset * dt, dt)),
                  min(n - s.n_taps, _idx(rightendpoint(config.current_window), t1 + s.offset * dt, dt)))
    cw = (curw(sg0, 0), curw(sg1, 1), curw(sg2, 2), (_idx(leftendpoint(config.current_window), t1, dt), _idx(rightendpoint(config.current_window), t1, dt)))
    RC = ustrip(NoUnits, τ / dt)
    IcpcParams(UInt32(sizeof(IcpcParams)), PARAMS_VERSION, n, 0x7f,                      # LGDSP_GROUP_ALL
        ustrip(u"ns", t1), ustrip(u"ns", dt),
        0, 2^kw.fc_bit_depth - kw.fc_bit_depth,                                                  # src/dsp_icpc.jl:93-94
        _idx(leftendpoint(config.bl_window), t1, dt), _idx(rightendpoint(config.bl_window), t1, dt),
        _idx(leftendpoint(config.tail_window), t1, dt), _idx(rightendpoint(config.tail_window), t1, dt),
        (RC + 1) / RC - 1,                                                                        # 1/alpha - 1
        _trap(t0p[1], t0p[2], t0p[3], dt), _trap(40u"ns", 100u"ns", 2000u"ns", dt),              # src/dsp_routines.jl:9
        config.t0_threshold, max(1, _cnt(kw.t0_mintot, dt)), max(1, _cnt(kw.tx_mintot, dt)),
        (0.1, 0.5, 0.8, 0.9, 0.99),
        ustrip(u"ns", first(config.qdrift_int_length)), ustrip(u"ns", last(config.qdrift_int_length)),
        ustrip(u"ns", first(config.lq_int_length)), ustrip(u"ns", last(config.lq_int_length)),
        _dni(kw.int_interpolation_order, kw.int_interpolation_length, dt), _dni(kw.sig_interpolation_order, kw.sig_interpolation_length, dt),
        _trap(10u"μs", 4u"μs", 10u"μs", dt), _trap(5u"μs", 3u"μs", 5u"μs", dt), _trap(3u"μs", 1u"μs", 3u"μs", dt), _trap(trap_rt, trap_ft, trap_rt, dt),
        ustrip(u"ns", trap_rt + trap_ft / 2), ustrip(u"ns", config.flt_length_cusp / 2), ustrip(u"ns", config.flt_length_zac / 2),
        (sg0, sg1, sg2), Int32.(first.(cw)), Int32.(last.(cw)),
        config.inTraceCut_std_threshold, max(1, _cnt(kw.intrace_mintot, dt)),
        _idx(leftendpoint(config.bl_window) + first(t) , t1 + sg0.offset * dt, dt), _idx(rightendpoint(config.bl_window), t1 + sg0.offset * dt, dt),  # src/dsp_routines.jl:75
        0, 0,
        _cz(:lgdsp_cusp_coeffs, cusp_rt, cusp_ft, 10000000.0u"μs", config.flt_length_cusp, dt),
        _cz(:lgdsp_zac_coeffs, zac_rt, zac_ft, 10000000.0u"μs", config.flt_length_zac, dt))
end

const COLS = (:blmean, :blsigma, :blslope, :bloffset, :tailmean, :tailsigma, :tailslope, :tailoffset, :qc_label,
    :t0, :t10, :t50, :t80, :t90, :t99, :t50_current, :drift_time, :tail_τ, :tail_mean, :tail_sigma, :e_max, :e_min,
    :e_10410, :e_535, :e_313, :e_10410_inv, :e_313_inv, :t0_inv, :e_trap, :e_cusp, :e_zac, :e_trap_max, :e_cusp_max, :e_zac_max,
    :t_trap_max, :t_cusp_max, :t_zac_max, :qdrift, :lq, :a_sg, :a_60, :a_100, :a_raw,
    :inTrace_intersect, :inTrace_n, :n_sat_low, :n_sat_high, :n_sat_low_cons, :n_sat_high_cons)
const UNIT = Dict(:blslope => u"ns^-1", :tailslope => u"ns^-1", :drift_time => u"ns", :tail_τ => u"ns", :t_trap_max => u"ns",
    :t_cusp_max => u"ns", :t_zac_max => u"ns", :inTrace_intersect => u"ns",
    (c => u"μs" for c in (:t0, :t10, :t50, :t80, :t90, :t99, :t50_current, :t0_inv))...)
const INTCOLS = (:qc_label, :inTrace_n, :n_sat_low, :n_sat_high, :n_sat_low_cons, :n_sat_high_cons)

const _handle = Ref{Union{Nothing, Handle}}(nothing)
handle() = something(_handle[], (_handle[] = Handle(parse(Int, get(ENV, "LGDSP_B200_DEVICE", "0")))))

"""
    dsp_icpc(data::Table, config::DSPConfig, τ, pars_filter::PropDict; f_evaluate_qc = missing)

Same signature and output table as `LegendDSP.dsp_icpc` (src/dsp_icpc.jl:62-230); the arithmetic runs in liblgdsp_b200.
"""
function dsp_icpc(data, config::DSPConfig, τ::Quantity, pars_filter::PropDict; f_evaluate_qc = missing)
    ismissing(f_evaluate_qc) || throw(ArgumentError("f_evaluate_qc is not supported by the B200 path (qc_label = -1)"))
    wvfs = data.waveform
    sig = wvfs.signal
    flat = sig isa ArrayOfSimilarVectors{UInt16} ? flatview(sig) : reduce(hcat, (collect(UInt16, s) for s in sig))   # n_samples x n_events
    n_samples, n_events = size(flat)
    p = Ref(resolve_params(wvfs, config, τ, pars_filter))
    rows = Matrix{Float64}(undef, NCOL, n_events)                   # column-major: one 49-double row per event
    h = handle()
    GC.@preserve flat rows p begin
        rc = ccall((:lgdsp_icpc_run, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{IcpcParams}, Ptr{UInt16}, Int64, Int64, Ptr{Float64}),
                   h.ptr, p, flat, n_events, stride(flat, 2), rows)
    end
    rc == 0 || (rc == -1 ? throw(ArgumentError(lasterr(h))) : error("lgdsp_icpc_run: " * lasterr(h)))
    cols = map(enumerate(COLS)) do (i, c)
        v = rows[i, :]
        c => (c in INTCOLS ? Int.(v) : haskey(UNIT, c) ? v .* UNIT[c] : v)
    end
    TypedTables.Table(; cols..., blfc = data.baseline, timestamp = data.timestamp, eventID_fadc = data.eventnumber, e_fc = data.daqenergy)
end

end # module
