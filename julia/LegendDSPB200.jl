# LegendDSPB200.jl -- drop-in for the LegendDSP.jl entry points of the `dsp_icpc` hot path on top of liblgdsp_b200.so
# (include/lgdsp_b200.h):
#
#     dsp_icpc(data, config, τ, pars_filter)                     src/dsp_icpc.jl:62
#     dsp_icpc_compressed(data, config, τ, pars_filter)          src/dsp_icpc.jl:293   (encoded waveforms are decoded ON THE GPU)
#     dsp_trap_rt_optimization(wvfs, config, τ; ft)              src/dsp_filter_optimization.jl:102
#     dsp_trap_ft_optimization(wvfs, config, τ, rt)              src/dsp_filter_optimization.jl:241
#
# UNTESTED in this repository's environment: the build image has no Julia toolchain.  What IS checked here:
# tests/test_abi.py keeps julia/lgdsp_offsets.jl (struct layouts measured by the C compiler), the column list and the constants
# below in sync with the header and the Python ctypes mirror.  Parameter blocks are filled as raw byte buffers at those
# offsets -- no 66 KB isbits structs with 4096-element tuples.  The index arithmetic uses the reference's own Unitful
# expressions (src/tailstats.jl:16-18, src/dsp_routines.jl:9-25, src/dsp_icpc.jl:87-99).
#
# Usage:   using LegendDSP, LegendDSPB200
#          tbl = LegendDSPB200.dsp_icpc(data, config, τ, pars_filter)      # same Table as LegendDSP.dsp_icpc
module LegendDSPB200

using Unitful, TypedTables, PropDicts, ArraysOfArrays, RadiationDetectorSignals, IntervalSets
import LegendDSP: DSPConfig, get_fltpars

include("lgdsp_offsets.jl")

const LIB = get(ENV, "LGDSP_B200_LIB", "liblgdsp_b200.so")
const NCOL = LGDSP_NCOL

# ---------------------------------------------------------------------------------------------------
# handle
# ---------------------------------------------------------------------------------------------------
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
const _handle = Ref{Union{Nothing, Handle}}(nothing)
handle() = something(_handle[], (_handle[] = Handle(parse(Int, get(ENV, "LGDSP_B200_DEVICE", "0")))))
_check(h, rc, what) = rc == 0 || (rc == -1 ? throw(ArgumentError(lasterr(h))) : error(what * ": " * lasterr(h)))

"""page-lock an array that is reused across calls (e.g. the read buffer of an LH5 reader): the host entry points then copy
straight from it instead of staging through the library's pinned ring.  Call `unpin!` before the array is freed."""
pin!(a::Array) = (ccall((:lgdsp_host_register, LIB), Cint, (Ptr{Cvoid}, Int64), a, sizeof(a)) == 0 || error("lgdsp_host_register"); a)
unpin!(a::Array) = (ccall((:lgdsp_host_unregister, LIB), Cint, (Ptr{Cvoid},), a); a)

# ---------------------------------------------------------------------------------------------------
# parameter blocks as byte buffers
# ---------------------------------------------------------------------------------------------------
put!(buf::Vector{UInt8}, off::Integer, v::T) where {T} = GC.@preserve buf unsafe_store!(Ptr{T}(pointer(buf) + off), v)
putv!(buf::Vector{UInt8}, off::Integer, v::AbstractVector{T}) where {T} = for (i, x) in enumerate(v); put!(buf, off + (i - 1) * sizeof(T), x); end

# the reference's own rounding: round(Int, ...) is ties-to-even on the Unitful quotient
_idx(t, first_x, step_x) = round(Int, ustrip(NoUnits, (t - first_x) / step_x))           # 0-based (src/tailstats.jl:16-18)
_cnt(t, step) = round(Int, ustrip(NoUnits, t / step))
_minn(t, step) = max(1, _cnt(t, step))

function put_trap!(buf, off, a, g, a2, dt)
    put!(buf, off + OFF_LGDSP_TRAP.navg, Int32(_cnt(a, dt))); put!(buf, off + OFF_LGDSP_TRAP.ngap, Int32(_cnt(g, dt)))
    put!(buf, off + OFF_LGDSP_TRAP.navg2, Int32(_cnt(a2, dt)))
end
function put_dni!(buf, off, order::Integer, len, dt)
    n_w = _cnt(len, dt)
    (order + 1 <= n_w <= LGDSP_MAX_DNI) || throw(ArgumentError("PolynomialDNI($order, $len) does not fit the sampling step"))
    A = zeros(Float64, LGDSP_MAX_DNI * 4)
    ccall((:lgdsp_lsq_fit_matrix, LIB), Cint, (Int32, Int32, Ptr{Float64}), n_w, order, A) == 0 || error("lgdsp_lsq_fit_matrix")
    put!(buf, off + OFF_LGDSP_DNI.n_w, Int32(n_w)); put!(buf, off + OFF_LGDSP_DNI.degree, Int32(order)); putv!(buf, off + OFF_LGDSP_DNI.A, A)
end
"""SavitzkyGolayFilter(len, degree, 1): returns (n_taps, offset) of the kernel written at `off`"""
function put_sg!(buf, off, len, degree, dt)
    n = _cnt(len, dt); n += iseven(n)                      # RddspPolicy.sg_even_length = "up" (DESIGN.md section 2)
    h = zeros(Float64, LGDSP_MAX_SG)
    ccall((:lgdsp_sg_coeffs, LIB), Cint, (Int32, Int32, Int32, Ptr{Float64}), n, degree, 1, h) == 0 || error("lgdsp_sg_coeffs")
    put!(buf, off + OFF_LGDSP_SG.n_taps, Int32(n)); put!(buf, off + OFF_LGDSP_SG.offset, Int32((n - 1) ÷ 2)); putv!(buf, off + OFF_LGDSP_SG.h, h)
    (n, (n - 1) ÷ 2)
end
function put_cz!(buf, off, sym::Symbol, rt, ft, τcz, len, dt)
    L = _cnt(len, dt); c = zeros(Float64, LGDSP_MAX_FIR)
    σ, τs, β = ustrip(NoUnits, rt / dt), ustrip(NoUnits, τcz / dt), ustrip(NoUnits, len / dt)   # src/dsp_icpc.jl:88,90
    ccall((sym, LIB), Cint, (Float64, Int32, Float64, Int32, Float64, Ptr{Float64}), σ, _cnt(ft, dt), τs, L, β, c) == 0 || error(string(sym))
    O = OFF_LGDSP_CUSPZAC
    put!(buf, off + O.n_taps, Int32(L)); put!(buf, off + O.flat, Int32(_cnt(ft, dt))); put!(buf, off + O.sigma, σ)
    put!(buf, off + O.tau, τs); put!(buf, off + O.beta, β); putv!(buf, off + O.coeffs, c)
end
function _window(w, first_x, dt, n, what)
    a, b = _idx(leftendpoint(w), first_x, dt), _idx(rightendpoint(w), first_x, dt)
    (0 <= a <= b <= n - 1) || throw(AssertionError("$what: index range $(a + 1):$(b + 1) outside 1:$n"))   # src/tailstats.jl:23-25
    (Int32(a), Int32(b))
end

"""resolve DSPConfig + τ + pars_filter for a waveform time axis into an `lgdsp_icpc_params` byte buffer
(mirrors legenddsp.jl_b200/config.py `resolve_icpc_params`, role "full")"""
function icpc_params(t::AbstractRange, config::DSPConfig, τ, pars_filter::PropDict; groups::Integer = LGDSP_GROUP_ALL)
    dt = step(t); t1 = first(t); n = length(t)
    kw = config.kwargs_pars
    buf = zeros(UInt8, SIZEOF_LGDSP_ICPC_PARAMS)
    O = OFF_LGDSP_ICPC_PARAMS
    put!(buf, O.struct_size, UInt32(SIZEOF_LGDSP_ICPC_PARAMS)); put!(buf, O.version, UInt32(LGDSP_PARAMS_VERSION))
    put!(buf, O.n_samples, Int32(n)); put!(buf, O.groups, UInt32(groups))
    put!(buf, O.t_first_ns, Float64(ustrip(u"ns", t1))); put!(buf, O.dt_ns, Float64(ustrip(u"ns", dt)))
    put!(buf, O.sat_low, Int64(0)); put!(buf, O.sat_high, Int64(2^kw.fc_bit_depth - kw.fc_bit_depth))         # src/dsp_icpc.jl:93-94
    bl = _window(config.bl_window, t1, dt, n, "bl_window"); tl = _window(config.tail_window, t1, dt, n, "tail_window")
    put!(buf, O.bl_from, bl[1]); put!(buf, O.bl_until, bl[2]); put!(buf, O.tail_from, tl[1]); put!(buf, O.tail_until, tl[2])
    RC = ustrip(NoUnits, τ / dt)
    put!(buf, O.pz_km1, Float64((RC + 1) / RC - 1))                                                             # 1/alpha - 1
    t0p = get(kw, :t0_flt_pars, [40u"ns", 100u"ns", 2000u"ns"])
    put_trap!(buf, O.t0_trap, t0p[1], t0p[2], t0p[3], dt)
    put_trap!(buf, O.t0inv_trap, 40u"ns", 100u"ns", 2000u"ns", dt)                                             # src/dsp_routines.jl:9
    put!(buf, O.t0_threshold, Float64(config.t0_threshold))
    put!(buf, O.t0_min_n, Int32(_minn(kw.t0_mintot, dt))); put!(buf, O.tx_min_n, Int32(_minn(kw.tx_mintot, dt)))
    putv!(buf, O.tx_frac, [0.1, 0.5, 0.8, 0.9, 0.99])                                                          # src/dsp_icpc.jl:132-136
    put!(buf, O.qdrift_first_ns, Float64(ustrip(u"ns", first(config.qdrift_int_length))))
    put!(buf, O.qdrift_last_ns, Float64(ustrip(u"ns", last(config.qdrift_int_length))))
    put!(buf, O.lq_first_ns, Float64(ustrip(u"ns", first(config.lq_int_length))))
    put!(buf, O.lq_last_ns, Float64(ustrip(u"ns", last(config.lq_int_length))))
    put_dni!(buf, O.int_dni, kw.int_interpolation_order, kw.int_interpolation_length, dt)
    put_dni!(buf, O.sig_dni, kw.sig_interpolation_order, kw.sig_interpolation_length, dt)
    trap_rt, trap_ft = get_fltpars(pars_filter, :trap, config)
    cusp_rt, cusp_ft = get_fltpars(pars_filter, :cusp, config)
    zac_rt, zac_ft = get_fltpars(pars_filter, :zac, config)
    sg_wl = get_fltpars(pars_filter, :sg, config)
    put_trap!(buf, O.trap_10410, 10u"μs", 4u"μs", 10u"μs", dt); put_trap!(buf, O.trap_535, 5u"μs", 3u"μs", 5u"μs", dt)
    put_trap!(buf, O.trap_313, 3u"μs", 1u"μs", 3u"μs", dt); put_trap!(buf, O.trap_e, trap_rt, trap_ft, trap_rt, dt)
    put!(buf, O.trap_pickoff_ns, Float64(ustrip(u"ns", trap_rt + trap_ft / 2)))
    put!(buf, O.cusp_pickoff_ns, Float64(ustrip(u"ns", config.flt_length_cusp / 2)))
    put!(buf, O.zac_pickoff_ns, Float64(ustrip(u"ns", config.flt_length_zac / 2)))
    sg = [put_sg!(buf, O.sg + (k - 1) * SIZEOF_LGDSP_SG, wl, config.sg_flt_degree, dt) for (k, wl) in enumerate((sg_wl, 60u"ns", 100u"ns"))]
    for k in 1:3      # current_window in the index space of every SG trace (first time = t1 + offset*dt), then of the waveform
        w = _window(config.current_window, t1 + sg[k][2] * dt, dt, n - sg[k][1] + 1, "current_window on sg[$k]")
        put!(buf, O.cur_from + 4 * (k - 1), w[1]); put!(buf, O.cur_until + 4 * (k - 1), w[2])
    end
    w = _window(config.current_window, t1, dt, n, "current_window")
    put!(buf, O.cur_from + 12, w[1]); put!(buf, O.cur_until + 12, w[2])
    put!(buf, O.intrace_nsigma, Float64(config.inTraceCut_std_threshold)); put!(buf, O.intrace_min_n, Int32(_minn(kw.intrace_mintot, dt)))
    first_sg = t1 + sg[1][2] * dt                                                                              # src/dsp_routines.jl:75
    ia, ib = _idx(leftendpoint(config.bl_window) + first_sg, first_sg, dt), _idx(rightendpoint(config.bl_window), first_sg, dt)
    (0 <= ia <= ib <= n - sg[1][1]) || throw(AssertionError("in-trace sigma window outside the SG trace"))
    put!(buf, O.intrace_bl_from, Int32(ia)); put!(buf, O.intrace_bl_until, Int32(ib))
    put!(buf, O.cuspzac_direct, Int32(0))
    put_cz!(buf, O.cusp, :lgdsp_cusp_coeffs, cusp_rt, cusp_ft, 10000000.0u"μs", config.flt_length_cusp, dt)
    put_cz!(buf, O.zac, :lgdsp_zac_coeffs, zac_rt, zac_ft, 10000000.0u"μs", config.flt_length_zac, dt)
    buf
end

"""`lgdsp_sweep_params` byte buffer (legenddsp.jl_b200/config.py `resolve_sweep_params`)"""
function sweep_params(t::AbstractRange, config::DSPConfig, τ; out_f64::Bool)
    dt = step(t); t1 = first(t); n = length(t); kw = config.kwargs_pars
    buf = zeros(UInt8, SIZEOF_LGDSP_SWEEP_PARAMS); O = OFF_LGDSP_SWEEP_PARAMS
    put!(buf, O.struct_size, UInt32(SIZEOF_LGDSP_SWEEP_PARAMS)); put!(buf, O.version, UInt32(LGDSP_PARAMS_VERSION))
    put!(buf, O.n_samples, Int32(n)); put!(buf, O.tx_min_n, Int32(_minn(kw.tx_mintot, dt)))
    put!(buf, O.t_first_ns, Float64(ustrip(u"ns", t1))); put!(buf, O.dt_ns, Float64(ustrip(u"ns", dt)))
    bl = _window(config.bl_window, t1, dt, n, "bl_window")
    put!(buf, O.bl_from, bl[1]); put!(buf, O.bl_until, bl[2])
    RC = ustrip(NoUnits, τ / dt); put!(buf, O.pz_km1, Float64((RC + 1) / RC - 1))
    put_dni!(buf, O.sig_dni, kw.sig_interpolation_order, kw.sig_interpolation_length, dt)
    put!(buf, O.out_f64, Int32(out_f64))
    buf
end
"""array of `lgdsp_sweep_variant` (kind 0: trapezoid) for (rt, ft) pairs; mode 0: fixed pick-off, 1: t50 + pick-off"""
function trap_variants(pairs, dt, pickoffs, mode::Integer)
    buf = zeros(UInt8, SIZEOF_LGDSP_SWEEP_VARIANT * length(pairs)); O = OFF_LGDSP_SWEEP_VARIANT
    for (i, ((rt, ft), pk)) in enumerate(zip(pairs, pickoffs))
        off = (i - 1) * SIZEOF_LGDSP_SWEEP_VARIANT
        put!(buf, off + O.kind, Int32(0)); put!(buf, off + O.pickoff_mode, Int32(mode))
        put!(buf, off + O.pickoff_ns, Float64(ustrip(u"ns", pk)))
        put_trap!(buf, off + O.trap, rt, ft, rt, dt)
    end
    buf
end

# ---------------------------------------------------------------------------------------------------
# tables
# ---------------------------------------------------------------------------------------------------
const COLS = (:blmean, :blsigma, :blslope, :bloffset, :tailmean, :tailsigma, :tailslope, :tailoffset, :qc_label,
    :t0, :t10, :t50, :t80, :t90, :t99, :t50_current, :drift_time, :tail_τ, :tail_mean, :tail_sigma, :e_max, :e_min,
    :e_10410, :e_535, :e_313, :e_10410_inv, :e_313_inv, :t0_inv, :e_trap, :e_cusp, :e_zac, :e_trap_max, :e_cusp_max, :e_zac_max,
    :t_trap_max, :t_cusp_max, :t_zac_max, :qdrift, :lq, :a_sg, :a_60, :a_100, :a_raw,
    :inTrace_intersect, :inTrace_n, :n_sat_low, :n_sat_high, :n_sat_low_cons, :n_sat_high_cons)
const UNIT = Dict(:blslope => u"ns^-1", :tailslope => u"ns^-1", :drift_time => u"ns", :tail_τ => u"ns", :t_trap_max => u"ns",
    :t_cusp_max => u"ns", :t_zac_max => u"ns", :inTrace_intersect => u"ns",
    (c => u"μs" for c in (:t0, :t10, :t50, :t80, :t90, :t99, :t50_current, :t0_inv))...)
const INTCOLS = (:qc_label, :inTrace_n, :n_sat_low, :n_sat_high, :n_sat_low_cons, :n_sat_high_cons)
_col(c, v) = c in INTCOLS ? Int.(v) : haskey(UNIT, c) ? v .* UNIT[c] : v
_colidx(c::Symbol) = findfirst(==(c), COLS)

"""dense `n_samples x n_events` matrix of the samples (each waveform contiguous, as `flatview` of an LH5 column)"""
_flat(sig, ::Type{T}) where {T} = sig isa ArrayOfSimilarVectors{T} ? flatview(sig) : reduce(hcat, (collect(T, s) for s in sig))

"""
    dsp_icpc(data::Table, config::DSPConfig, τ, pars_filter::PropDict; f_evaluate_qc = missing)

Same signature and output table as `LegendDSP.dsp_icpc` (src/dsp_icpc.jl:62-230); the arithmetic runs in liblgdsp_b200.
"""
function dsp_icpc(data, config::DSPConfig, τ::Quantity, pars_filter::PropDict; f_evaluate_qc = missing)
    ismissing(f_evaluate_qc) || throw(ArgumentError("f_evaluate_qc is not supported by the B200 path (qc_label = -1)"))
    wvfs = data.waveform
    flat = _flat(wvfs.signal, UInt16)                                   # n_samples x n_events
    n_samples, n_events = size(flat)
    p = icpc_params(wvfs[1].time, config, τ, pars_filter)
    rows = Matrix{Float64}(undef, NCOL, n_events)                       # column-major: one 49-double row per event
    h = handle()
    GC.@preserve flat rows p begin
        rc = ccall((:lgdsp_icpc_run, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt16}, Int64, Int64, Ptr{Float64}),
                   h.ptr, p, flat, n_events, stride(flat, 2), rows)
    end
    _check(h, rc, "lgdsp_icpc_run")
    cols = [c => _col(c, rows[i, :]) for (i, c) in enumerate(COLS)]
    TypedTables.Table(; cols..., blfc = data.baseline, timestamp = data.timestamp, eventID_fadc = data.eventnumber, e_fc = data.daqenergy)
end

# ---- dsp_icpc_compressed: src/dsp_icpc.jl:293-499 ----
"""encoded waveform column (`VectorOfEncodedArrays` of LegendDataTypes): (codec id, bytes, element pointers, shift)"""
function _encoded(col)
    codec = col.codec
    name = string(nameof(typeof(codec)))
    id = occursin("Radware", name) ? LGDSP_CODEC_RADWARE : LGDSP_CODEC_ULEB128ZZD
    shift = id == LGDSP_CODEC_RADWARE ? Int32(getfield(codec, :shift)) : Int32(0)
    vv = col.encoded                                             # VectorOfVectors{UInt8}
    (Int32(id), flatview(vv), Int64.(vv.elem_ptr .- first(vv.elem_ptr)), shift)
end
_isencoded(sig) = hasproperty(sig, :codec) && hasproperty(sig, :encoded)

"""
    dsp_icpc_compressed(data::Table, config::DSPConfig, τ, pars_filter::PropDict)

Same table as `LegendDSP.dsp_icpc_compressed`.  When `data.waveform_presummed.signal` / `data.waveform_windowed.signal` are still
ENCODED (the form LH5 files hold), the reference's `decode_data` calls (:313-314) run on the GPU: only the encoded bytes cross
the host link.
"""
function dsp_icpc_compressed(data, config::DSPConfig, τ::Quantity, pars_filter::PropDict; f_evaluate_qc = missing)
    ismissing(f_evaluate_qc) || throw(ArgumentError("f_evaluate_qc is not supported by the B200 path (qc_label = -1)"))
    wp, ww = data.waveform_presummed, data.waveform_windowed
    presum = only(unique(data.presum_rate))                                                    # :324
    n_events = length(wp)
    h = handle()
    # the two parameter blocks: the presummed pass (energies, tail, saturation, in-trace) and the windowed pass (timing, currents)
    ppre = icpc_params(wp[1].time, _pre_config(config, presum), τ, pars_filter; groups = 0x01 | 0x02 | 0x04 | 0x10 | 0x40)
    put!(ppre, OFF_LGDSP_ICPC_PARAMS.sat_high, Int64((2^config.kwargs_pars.fc_bit_depth - config.kwargs_pars.fc_bit_depth) * presum))   # :334
    pwdw = icpc_params(ww[1].time, config, τ, pars_filter; groups = 0x01 | 0x02 | 0x08 | 0x20)
    tp = wp[1].time
    aux = Int32[]                                                                              # auxbl1, auxbl2, auxpz1, auxpz2  (:338-339, :365-366)
    for w in (config.auxbl1_window, config.auxbl2_window, config.auxpz1_window, config.auxpz2_window)
        append!(aux, _window(w, first(tp), step(tp), length(tp), "aux window"))
    end
    rows_pre = Matrix{Float64}(undef, NCOL, n_events); rows_wdw = similar(rows_pre)
    stats = Array{Float64, 3}(undef, LGDSP_NSTAT, 5, n_events)
    if _isencoded(wp.signal) && _isencoded(ww.signal)
        cp, bp, op, sp = _encoded(wp.signal); cw, bw, ow, sw = _encoded(ww.signal)
        GC.@preserve bp op bw ow ppre pwdw aux rows_pre rows_wdw stats begin
            rc = ccall((:lgdsp_icpc_compressed_run_encoded, LIB), Cint,
                       (Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt8}, Int32, Ptr{UInt8}, Ptr{Int64}, Int32, Int32, Int32, Ptr{UInt8}, Ptr{Int64}, Int32, Int32,
                        Float64, Ptr{Int32}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                       h.ptr, ppre, pwdw, cp, bp, op, sp, Int32(4), cw, bw, ow, sw, Int32(2), Float64(presum), aux, n_events, rows_pre, rows_wdw, stats)
        end
        _check(h, rc, "lgdsp_icpc_compressed_run_encoded")
    else
        fp = _flat(wp.signal, UInt32); fw = _flat(ww.signal, UInt16)
        GC.@preserve fp fw ppre pwdw aux rows_pre rows_wdw stats begin
            rc = ccall((:lgdsp_icpc_compressed_run, LIB), Cint,
                       (Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt8}, Ptr{Cvoid}, Int32, Int64, Ptr{Cvoid}, Int32, Int64, Float64, Ptr{Int32}, Int64,
                        Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                       h.ptr, ppre, pwdw, fp, Int32(4), stride(fp, 2), fw, Int32(2), stride(fw, 2), Float64(presum), aux, n_events, rows_pre, rows_wdw, stats)
        end
        _check(h, rc, "lgdsp_icpc_compressed_run")
    end
    _compressed_table(data, rows_pre, rows_wdw, stats)
end
# the in-trace / current-rise filter of the presummed waveform is SavitzkyGolayFilter(sg_wl * presum_rate / 2)  (:439): the
# presummed pass resolves with that window length as its sg default
function _pre_config(config::DSPConfig, presum)
    d = deepcopy(config.default_flt_param); d.sg.wl = d.sg.wl * presum / 2
    DSPConfig{typeof(config.t0_threshold)}((f == :default_flt_param ? d : getfield(config, f) for f in fieldnames(DSPConfig))...)
end
# which waveform every column of the reference's result table comes from: src/dsp_icpc.jl:463-499
const FROM_WDW = (:t0, :t10, :t50, :t80, :t90, :t99, :drift_time, :qdrift, :lq, :a_sg, :a_60, :a_100, :a_raw, :t50_current, :t0_inv)
function _compressed_table(data, rp, rw, st)
    cols = Pair{Symbol, Any}[]
    for c in COLS
        src = c in FROM_WDW ? rw : rp
        push!(cols, c => _col(c, src[_colidx(c), :]))
    end
    push!(cols, :e_max_pre => rp[_colidx(:e_max), :], :e_min_pre => rp[_colidx(:e_min), :], :t50_pre => rp[_colidx(:t50), :] .* u"μs")
    push!(cols, :e_max => rw[_colidx(:e_max), :], :e_min => rw[_colidx(:e_min), :])
    for (w, name) in enumerate((:auxbl1, :auxbl2, :bl, :auxpz1, :auxpz2)), (k, f) in ((1, :mean), (2, :sigma), (5, :slope_sigma))
        name == :bl && f != :slope_sigma && continue
        push!(cols, Symbol(name, :_, f) => st[k, w, :])
    end
    TypedTables.Table(; cols..., blfc = data.baseline, timestamp = data.timestamp, eventID_fadc = data.eventnumber, e_fc = data.daqenergy)
end

# ---- trapezoid sweeps: src/dsp_filter_optimization.jl:102-133, 241-274 ----
function _trap_sweep(wvfs, config, τ, pairs, pickoffs, mode, ::Type{T}) where {T <: Union{Float32, Float64}}
    flat = _flat(wvfs.signal, UInt16); n_samples, n_events = size(flat)
    t = wvfs[1].time
    sp = sweep_params(t, config, τ; out_f64 = T === Float64)
    var = trap_variants(pairs, step(t), pickoffs, mode)
    out = Matrix{T}(undef, length(pairs), n_events)                     # == the reference's (n_grid x n_events) column-major matrix
    h = handle()
    GC.@preserve flat sp var out begin
        rc = ccall((:lgdsp_sweep_run, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt16}, Int64, Int64, Ptr{UInt8}, Int32, Ptr{Cvoid}, Ptr{Float64}),
                   h.ptr, sp, flat, n_events, stride(flat, 2), var, Int32(length(pairs)), out, C_NULL)
    end
    _check(h, rc, "lgdsp_sweep_run")
    out
end

"""`dsp_trap_rt_optimization(wvfs, config, τ; ft)` -> Matrix{Float64}(n_rt, n_events): ENC grid at the fixed pick-off
`config.enc_pickoff_trap` (src/dsp_filter_optimization.jl:102-133)"""
function dsp_trap_rt_optimization(wvfs, config::DSPConfig, τ::Quantity; ft = 2.0u"μs")
    rts = collect(config.e_grid_rt_trap)
    _trap_sweep(wvfs, config, τ, [(rt, ft) for rt in rts], fill(config.enc_pickoff_trap, length(rts)), 0, Float64)
end
"""`dsp_trap_ft_optimization(wvfs, config, τ, rt)` -> Matrix{Float32}(n_ft, n_events): energy grid at t50 + rt + ft/2
(src/dsp_filter_optimization.jl:241-274)"""
function dsp_trap_ft_optimization(wvfs, config::DSPConfig, τ::Quantity, rt)
    fts = collect(config.e_grid_ft_trap)
    _trap_sweep(wvfs, config, τ, [(rt, ft) for ft in fts], [rt + ft / 2 for ft in fts], 1, Float32)
end

end # module
