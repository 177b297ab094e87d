"""ctypes wrapper around oracle/liblgdsp_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this module.
PARITY STATUS: "parity unpinned" for the RadiationDetectorDSP.jl parts (see lgdsp_oracle.c header).
"""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblgdsp_oracle.so")
_SO_FAST = os.path.join(_HERE, "liblgdsp_oracle_fast.so")   # -O3 -march=native build of the same sources: the TIMED CPU arm

_abi = importlib.import_module("legenddsp.jl_b200._abi")

ORC_NIDX = 16
IDX_NAMES = ("t0", "t10", "t50", "t80", "t90", "t99", "t50_current", "t0_inv", "trap_max", "cusp_max",
             "zac_max", "intrace")

_dp = C.POINTER(C.c_double)


def _cpu_signature():
    """model + ISA flags of this host: the fast build uses -march=native, so it has to be compiled where it runs"""
    import hashlib
    try:
        txt = open("/proc/cpuinfo").read()
        keep = sorted({l.split(":", 1)[1].strip() for l in txt.splitlines() if l.startswith(("flags", "model name"))})
        return hashlib.md5("|".join(keep).encode()).hexdigest()
    except OSError:
        return "unknown"


def build(force=False):
    deps = [os.path.join(_HERE, f) for f in ("lgdsp_oracle.c", "lgdsp_codec_oracle.c", "lgdsp_synth_oracle.c", "Makefile")]
    deps.append(os.path.join(_HERE, "..", "include", "lgdsp_b200.h"))
    newest = max(os.path.getmtime(d) for d in deps)
    stamp = os.path.join(_HERE, ".build_host")
    sig = _cpu_signature()
    same_host = os.path.exists(stamp) and open(stamp).read().strip() == sig
    if force or not same_host or not all(os.path.exists(f) and os.path.getmtime(f) >= newest for f in (_SO, _SO_FAST)):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
        with open(stamp, "w") as f:
            f.write(sig + "\n")
    return _SO


_lib = None
_fast = False


def use_fast_build(on=True):
    """switch this module to the -O3 -march=native build (bench.py's CPU arms); the tests stay on the strict build"""
    global _lib, _fast
    if on != _fast:
        _lib = None
        _fast = on


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO_FAST if _fast else _SO)
        L.orc_lsq_fit_matrix.argtypes = [C.c_int, C.c_int, _dp]
        L.orc_sg_coeffs.argtypes = [C.c_int, C.c_int, C.c_int, _dp]
        L.orc_cusp_coeffs.argtypes = [C.c_double, C.c_int, C.c_double, C.c_int, C.c_double, _dp]
        L.orc_zac_coeffs.argtypes = [C.c_double, C.c_int, C.c_double, C.c_int, C.c_double, _dp]
        L.orc_saturation.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]
        L.orc_saturation.restype = None
        for name in ("orc_signalstats", "orc_tailstats", "orc_extremestats"):
            f = getattr(L, name)
            f.argtypes = [_dp, C.c_double, C.c_double, C.c_int, C.c_int, _dp]
            f.restype = None
        L.orc_get_wvf_maximum.argtypes = [_dp, C.c_int, C.c_int]
        L.orc_get_wvf_maximum.restype = C.c_double
        L.orc_derivative.argtypes = [_dp, C.c_int, C.c_double, _dp]
        L.orc_derivative.restype = None
        L.orc_invcr.argtypes = [_dp, C.c_int, C.c_double, _dp]
        L.orc_invcr.restype = None
        L.orc_integrator.argtypes = [_dp, C.c_int, C.c_double, _dp]
        L.orc_integrator.restype = None
        L.orc_trap.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
        L.orc_trap_bruteforce.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
        L.orc_fir_valid.argtypes = [_dp, C.c_int, _dp, C.c_int, _dp]
        L.orc_corr_valid.argtypes = [_dp, C.c_int, _dp, C.c_int, _dp]
        L.orc_intersect.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                    C.POINTER(C.c_int64), C.POINTER(C.c_int)]
        L.orc_intersect.restype = C.c_double
        L.orc_dni.argtypes = [C.POINTER(_abi.Dni), _dp, C.c_int, C.c_double]
        L.orc_dni.restype = C.c_double
        L.orc_dsp_icpc.argtypes = [C.POINTER(_abi.IcpcParams), C.c_void_p, C.c_int64, C.c_int64, _dp,
                                   C.POINTER(C.c_int32), C.c_int]
        L.orc_trap_sweep.argtypes = [C.POINTER(_abi.SweepParams), C.c_void_p, C.c_int64, C.c_int64,
                                     C.POINTER(_abi.TrapVariant), C.c_int, C.POINTER(C.c_float), C.c_int]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


class OracleBuilders:
    """the oracle's own (independent) filter-coefficient construction, same interface as config.LibBuilders"""

    def lsq_fit_matrix(self, n, degree):
        A = np.zeros((n, degree + 1))
        if lib().orc_lsq_fit_matrix(n, degree, A.ctypes.data_as(_dp)) != 0:
            raise ValueError("orc_lsq_fit_matrix failed")
        return A

    def sg_coeffs(self, n_taps, degree, derivative):
        h = np.zeros(n_taps)
        if lib().orc_sg_coeffs(n_taps, degree, derivative, h.ctypes.data_as(_dp)) != 0:
            raise ValueError("orc_sg_coeffs failed")
        return h

    def cusp_coeffs(self, sigma, flat, tau, L, beta):
        c = np.zeros(L)
        if lib().orc_cusp_coeffs(sigma, flat, tau, L, beta, c.ctypes.data_as(_dp)) != 0:
            raise ValueError("orc_cusp_coeffs failed")
        return c

    def zac_coeffs(self, sigma, flat, tau, L, beta):
        c = np.zeros(L)
        if lib().orc_zac_coeffs(sigma, flat, tau, L, beta, c.ctypes.data_as(_dp)) != 0:
            raise ValueError("orc_zac_coeffs failed")
        return c


def saturation(y_u16, low, high):
    y = np.ascontiguousarray(y_u16, dtype=np.uint16)
    out = (C.c_int64 * 4)()
    lib().orc_saturation(y.ctypes.data, y.size, int(low), int(high), out)
    return dict(low=out[0], high=out[1], max_cons_low=out[2], max_cons_high=out[3])


def _stats(fn, y, t0, dt, frm, until, n):
    y, p = _d(y)
    out = np.zeros(n)
    getattr(lib(), fn)(p, float(t0), float(dt), int(frm), int(until), out.ctypes.data_as(_dp))
    return out


def signalstats(y, t0, dt, frm, until):
    o = _stats("orc_signalstats", y, t0, dt, frm, until, 4)
    return dict(mean=o[0], sigma=o[1], slope=o[2], offset=o[3])


def tailstats(y, t0, dt, frm, until):
    o = _stats("orc_tailstats", y, t0, dt, frm, until, 3)
    return dict(mean=o[0], sigma=o[1], tau=o[2])


def extremestats(y, t0, dt, frm, until):
    o = _stats("orc_extremestats", y, t0, dt, frm, until, 4)
    return dict(min=o[0], max=o[1], tmin=o[2], tmax=o[3])


def get_wvf_maximum(y, frm, until):
    y, p = _d(y)
    return lib().orc_get_wvf_maximum(p, int(frm), int(until))


def derivative(x, gain=1.0):
    x, p = _d(x)
    out = np.zeros_like(x)
    lib().orc_derivative(p, x.size, float(gain), out.ctypes.data_as(_dp))
    return out


def invcr(x, km1):
    x, p = _d(x)
    out = np.zeros_like(x)
    lib().orc_invcr(p, x.size, float(km1), out.ctypes.data_as(_dp))
    return out


def integrator(x, gain=1.0):
    x, p = _d(x)
    out = np.zeros_like(x)
    lib().orc_integrator(p, x.size, float(gain), out.ctypes.data_as(_dp))
    return out


def trap(y, navg, ngap, navg2, bruteforce=False):
    y, p = _d(y)
    out = np.zeros(max(y.size, 1))
    fn = lib().orc_trap_bruteforce if bruteforce else lib().orc_trap
    n = fn(p, y.size, navg, ngap, navg2, out.ctypes.data_as(_dp))
    return out[:n].copy()


def fir_valid(y, c):
    y, p = _d(y)
    c, pc = _d(c)
    out = np.zeros(max(y.size, 1))
    n = lib().orc_fir_valid(p, y.size, pc, c.size, out.ctypes.data_as(_dp))
    return out[:n].copy()


def corr_valid(y, h):
    y, p = _d(y)
    h, ph = _d(h)
    out = np.zeros(max(y.size, 1))
    n = lib().orc_corr_valid(p, y.size, ph, h.size, out.ctypes.data_as(_dp))
    return out[:n].copy()


def intersect(y, t0, dt, thr, min_n):
    y, p = _d(y)
    mult = C.c_int64(0)
    pos = C.c_int(-1)
    x = lib().orc_intersect(p, y.size, float(t0), float(dt), float(thr), int(min_n), C.byref(mult), C.byref(pos))
    return dict(x=x, multiplicity=mult.value, pos=pos.value)


def dni(dni_struct, y, p_frac):
    y, p = _d(y)
    return lib().orc_dni(C.byref(dni_struct), p, y.size, float(p_frac))


def dsp_icpc(params, wf_u16, n_threads=0, want_idx=False):
    """rows[n_events, NCOL] (and idx[n_events, 16]) of the oracle chain on wf_u16[n_events, ld]"""
    wf = np.ascontiguousarray(wf_u16, dtype=np.uint16)
    assert wf.ndim == 2 and wf.shape[1] >= params.n_samples
    n_ev, ld = wf.shape
    rows = np.zeros((n_ev, _abi.NCOL))
    idx = np.full((n_ev, ORC_NIDX), -1, dtype=np.int32) if want_idx else None
    used = lib().orc_dsp_icpc(C.byref(params), wf.ctypes.data, n_ev, ld, rows.ctypes.data_as(_dp),
                              idx.ctypes.data_as(C.POINTER(C.c_int32)) if want_idx else None, int(n_threads))
    return (rows, idx, used) if want_idx else (rows, used)


def signalstats5(y, t0, dt, frm, until):
    """signalstats incl. slope_residual_sigma (population sigma of the straight-line fit residuals; parity unpinned)"""
    y, p = _d(y)
    out = np.zeros(5)
    L = lib()
    L.orc_signalstats5.argtypes = [_dp, C.c_double, C.c_double, C.c_int, C.c_int, _dp]
    L.orc_signalstats5.restype = None
    L.orc_signalstats5(p, float(t0), float(dt), int(frm), int(until), out.ctypes.data_as(_dp))
    return out


def compressed_columns():
    L = lib()
    L.orc_compressed_columns.restype = C.c_char_p
    return tuple(c for c in L.orc_compressed_columns().decode().split(",") if c)


def dsp_icpc_compressed(p_pre, p_wdw, pre, wdw, presum_rate, aux_windows, n_threads=0):
    """oracle dsp_icpc_compressed (src/dsp_icpc.jl:293-499): dict column -> float64[n_events] of the computed columns.
    pre / wdw: uint16 or uint32 arrays [n_events, n_samples]; aux_windows: [(from, until)] x 4 (auxbl1, auxbl2, auxpz1,
    auxpz2), 0-based inclusive on the presummed axis."""
    def prep(a):
        a = np.asarray(a)
        assert a.ndim == 2 and a.dtype in (np.uint16, np.uint32)
        return np.ascontiguousarray(a)
    pre, wdw = prep(pre), prep(wdw)
    assert pre.shape[0] == wdw.shape[0] and pre.shape[1] >= p_pre.n_samples and wdw.shape[1] >= p_wdw.n_samples
    cols = compressed_columns()
    n_ev = pre.shape[0]
    rows = np.zeros((n_ev, len(cols)))
    aux = (C.c_int32 * 8)(*[int(v) for ab in aux_windows for v in ab])
    L = lib()
    L.orc_dsp_icpc_compressed.argtypes = [C.POINTER(_abi.IcpcParams), C.POINTER(_abi.IcpcParams), C.c_void_p, C.c_int, C.c_int64,
                                          C.c_void_p, C.c_int, C.c_int64, C.c_double, C.POINTER(C.c_int32), C.c_int64, _dp, C.c_int]
    L.orc_dsp_icpc_compressed(C.byref(p_pre), C.byref(p_wdw), pre.ctypes.data, pre.dtype.itemsize, pre.shape[1],
                              wdw.ctypes.data, wdw.dtype.itemsize, wdw.shape[1], float(presum_rate), aux, n_ev,
                              rows.ctypes.data_as(_dp), int(n_threads))
    return {c: rows[:, i].copy() for i, c in enumerate(cols)}


def thresholdstats(y, mn=-np.inf, mx=np.inf):
    y, p = _d(y)
    L = lib()
    L.orc_thresholdstats.argtypes = [_dp, C.c_int, C.c_double, C.c_double]
    L.orc_thresholdstats.restype = C.c_double
    return L.orc_thresholdstats(p, y.size, float(mn), float(mx))


def thresholdstats_mad(y, mn=-np.inf, mx=np.inf):
    y, p = _d(y)
    L = lib()
    L.orc_thresholdstats_mad.argtypes = [_dp, C.c_int, C.c_double, C.c_double]
    L.orc_thresholdstats_mad.restype = C.c_double
    return L.orc_thresholdstats_mad(p, y.size, float(mn), float(mx))


def intersect_maximum(y, t0, dt, thr, min_n, max_n, cap=256):
    """IntersectMaximum (src/intersect_maximum.jl:24-119): dict x, x_high, x_tot, max (arrays), multiplicity"""
    y, p = _d(y)
    out = [np.zeros(cap) for _ in range(4)]
    L = lib()
    L.orc_intersect_maximum.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                        _dp, _dp, _dp, _dp]
    L.orc_intersect_maximum.restype = C.c_int
    n = L.orc_intersect_maximum(p, y.size, float(t0), float(dt), float(thr), int(min_n), int(max_n), int(cap),
                                *[o.ctypes.data_as(_dp) for o in out])
    m = min(n, cap)
    return {"x": out[0][:m], "x_high": out[1][:m], "x_tot": out[2][:m], "max": out[3][:m], "multiplicity": n}


def dsp_sipm(params, wf, n_threads=0):
    """oracle dsp_sipm (src/dsp_sipm.jl:47-158): rows[n_events, SIPM_NCOL], trig[n_events, 4, 4, max_triggers]"""
    wf = np.ascontiguousarray(wf)
    assert wf.ndim == 2 and wf.dtype == (np.float32 if params.sample_kind == _abi.SAMPLE_F32 else np.uint16)
    n_ev, ld = wf.shape
    rows = np.zeros((n_ev, _abi.SIPM_NCOL))
    trig = np.zeros((n_ev, _abi.SIPM_NLIST, _abi.SIPM_NFIELD, params.max_triggers))
    L = lib()
    L.orc_dsp_sipm.argtypes = [C.POINTER(_abi.SipmParams), C.c_void_p, C.c_int64, C.c_int64, _dp, _dp, C.c_int]
    L.orc_dsp_sipm(C.byref(params), wf.ctypes.data, n_ev, ld, rows.ctypes.data_as(_dp), trig.ctypes.data_as(_dp), int(n_threads))
    return rows, trig


def qdrift_flt_optimization(params, wf_u16, blmean, n_threads=0):
    """oracle dsp_qdrift_flt_optimization (src/dsp_filter_optimization.jl:72-90): (qdrift[n_events], t0_us[n_events])"""
    wf = np.ascontiguousarray(wf_u16, dtype=np.uint16)
    bl = np.ascontiguousarray(blmean, dtype=np.float64)
    out = np.zeros((wf.shape[0], 2))
    L = lib()
    L.orc_qdrift_flt_optimization.argtypes = [C.POINTER(_abi.IcpcParams), C.c_void_p, C.c_int64, C.c_int64, _dp, _dp, C.c_int]
    L.orc_qdrift_flt_optimization(C.byref(params), wf.ctypes.data, wf.shape[0], wf.shape[1], bl.ctypes.data_as(_dp),
                                  out.ctypes.data_as(_dp), int(n_threads))
    return out[:, 0].copy(), out[:, 1].copy()


def multi_intersect(y, t0, dt, ratios, min_n, n=1, degree=1, rate=1):
    """MultiIntersect (src/multi_intersect.jl:10-121) on one trace: x[n_thr]; raises AssertionError like the reference (:85-88)"""
    y, p = _d(y)
    r = np.ascontiguousarray(ratios, dtype=np.float64)
    A = OracleBuilders().lsq_fit_matrix(2 * n, degree)
    A = np.ascontiguousarray(A, dtype=np.float64)
    x = np.zeros(len(r))
    L = lib()
    L.orc_multi_intersect.argtypes = [_dp, C.c_int, C.c_double, C.c_double, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp]
    rc = L.orc_multi_intersect(p, y.size, float(t0), float(dt), r.ctypes.data_as(_dp), len(r), int(min_n), int(n), int(degree),
                               int(rate), A.ctypes.data_as(_dp), x.ctypes.data_as(_dp))
    if rc != 0:
        raise AssertionError("cannot interpolate intersect on left boundary")
    return x


def trap_sweep(sparams, wf_u16, variants, n_threads=0):
    wf = np.ascontiguousarray(wf_u16, dtype=np.uint16)
    n_ev, ld = wf.shape
    nv = len(variants)
    out = np.zeros((n_ev, nv), dtype=np.float32)  # [event][variant] = Julia (n_variants x n_events) column-major
    lib().orc_trap_sweep(C.byref(sparams), wf.ctypes.data, n_ev, ld, variants, nv,
                         out.ctypes.data_as(C.POINTER(C.c_float)), int(n_threads))
    return out


def sweep(sparams, wf_u16, variants, n_threads=0, want_aux=False):
    """general sweep (orc_sweep): float64 out[n_events, n_variants] (+ aux[n_events, 4] = blmean, blslope, t50_us, 0)"""
    wf = np.ascontiguousarray(wf_u16, dtype=np.uint16)
    n_ev, ld = wf.shape
    nv = len(variants)
    out = np.zeros((n_ev, nv), dtype=np.float64)
    aux = np.zeros((n_ev, 4), dtype=np.float64) if want_aux else None
    L = lib()
    L.orc_sweep.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, _dp, _dp, C.c_int]
    L.orc_sweep(C.byref(sparams), wf.ctypes.data, n_ev, ld, C.cast(variants, C.c_void_p), nv, out.ctypes.data_as(_dp),
                aux.ctypes.data_as(_dp) if want_aux else None, int(n_threads))
    return (out, aux) if want_aux else out


def sweep_ext(sparams, wf, variants, baseline=None, n_threads=0, want_aux=False):
    """orc_sweep_ext: general sweep on uint16 / uint32 samples with an optional external per-event baseline"""
    wf = np.ascontiguousarray(wf)
    if wf.dtype not in (np.uint16, np.uint32):
        raise TypeError("uint16 or uint32 samples")
    n_ev, ld = wf.shape
    nv = len(variants)
    out = np.zeros((n_ev, nv), dtype=np.float64)
    aux = np.zeros((n_ev, 4), dtype=np.float64) if want_aux else None
    bl = None if baseline is None else np.ascontiguousarray(baseline, dtype=np.float64)
    L = lib()
    L.orc_sweep_ext.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _dp, C.c_int64, C.c_int64, C.c_void_p, C.c_int, _dp, _dp, C.c_int]
    L.orc_sweep_ext(C.byref(sparams), wf.ctypes.data, wf.dtype.itemsize, bl.ctypes.data_as(_dp) if bl is not None else None, n_ev, ld,
                    C.cast(variants, C.c_void_p), nv, out.ctypes.data_as(_dp), aux.ctypes.data_as(_dp) if want_aux else None,
                    int(n_threads))
    return (out, aux) if want_aux else out


def dsp_sg_optimization_compressed(S_pre, S_wdw, wf_pre, wf_wdw, energy_variant, sg_variants, presum_rate):
    """oracle dsp_sg_optimization_compressed (src/dsp_filter_optimization.jl:460-511) on resolved sweep parameters:
    baseline statistics, pole-zero correction, t50 and the trap(rt, ft) energy on the presummed waveform (:469-495), the
    Savitzky-Golay window-length sweep on the windowed waveform shifted by blmean / presum_rate (:477, :498-503)"""
    e, aux = sweep_ext(S_pre, wf_pre, energy_variant, want_aux=True)
    a = sweep_ext(S_wdw, wf_wdw, sg_variants, baseline=aux[:, 0] / float(presum_rate))
    with np.errstate(divide="ignore", invalid="ignore"):
        aoe = a / e[:, :1]
    return dict(aoe=aoe, energy=e[:, 0].copy(), blmean=aux[:, 0].copy(), blslope=aux[:, 1].copy(), t50=aux[:, 2].copy())


def num_threads():
    return lib().orc_num_threads()


# ---- waveform codecs of decode_data (lgdsp_codec_oracle.c) ----
def radware_encode(x_u16, shift=-32768, big_endian=True):
    """one waveform -> bytes (radware-sigcompress v1.0 restatement)"""
    L = lib()
    L.orc_radware_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64]
    L.orc_radware_encode.restype = C.c_int64
    x = np.ascontiguousarray(x_u16, dtype=np.uint16)
    out = np.empty(2 * x.size + 8 * (x.size // 48 + 2) + 16, dtype=np.uint8)
    nb = L.orc_radware_encode(x.ctypes.data, x.size, int(shift), 1 if big_endian else 0, out.ctypes.data, out.size)
    if nb < 0:
        raise ValueError(f"orc_radware_encode: {nb}")
    return out[:nb].copy()


def radware_decode(b_u8, n_max=65535, shift=-32768, big_endian=True):
    L = lib()
    L.orc_radware_decode.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int]
    b = np.ascontiguousarray(b_u8, dtype=np.uint8)
    out = np.zeros(n_max, dtype=np.uint16)
    n = L.orc_radware_decode(b.ctypes.data, b.size, int(shift), 1 if big_endian else 0, out.ctypes.data, n_max)
    if n < 0:
        raise ValueError("orc_radware_decode: malformed stream")
    return out[:n].copy()


def uleb128zzd_encode(x):
    L = lib()
    L.orc_uleb128zzd_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64]
    L.orc_uleb128zzd_encode.restype = C.c_int64
    x = np.ascontiguousarray(x)
    assert x.dtype in (np.uint16, np.uint32)
    out = np.empty(10 * x.size + 16, dtype=np.uint8)
    nb = L.orc_uleb128zzd_encode(x.ctypes.data, x.dtype.itemsize, x.size, out.ctypes.data, out.size)
    if nb < 0:
        raise ValueError("orc_uleb128zzd_encode")
    return out[:nb].copy()


def uleb128zzd_decode(b_u8, dtype=np.uint32, n_max=65536):
    L = lib()
    L.orc_uleb128zzd_decode.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int]
    b = np.ascontiguousarray(b_u8, dtype=np.uint8)
    out = np.zeros(n_max, dtype=dtype)
    n = L.orc_uleb128zzd_decode(b.ctypes.data, b.size, np.dtype(dtype).itemsize, out.ctypes.data, n_max)
    if n < 0:
        raise ValueError("orc_uleb128zzd_decode: malformed stream")
    return out[:n].copy()


def synth_generate(n_events, first_event=0, n_samples=8192, seed=20260101, mode=0, noise_sigma=3.0, tau_samples=31250.0):
    """the synthetic event stream (lgdsp_synth_oracle.c): the CPU arms' own generator, no product library involved"""
    L = lib()
    L.orc_synth_generate.argtypes = [C.POINTER(_abi.SynthParams), C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
    sp = _abi.SynthParams(seed, n_samples, mode, noise_sigma, tau_samples)
    out = np.zeros((int(n_events), n_samples), dtype=np.uint16)
    rc = L.orc_synth_generate(C.byref(sp), int(first_event), int(n_events), n_samples, out.ctypes.data)
    if rc != 0:
        raise ValueError("orc_synth_generate failed")
    return out


def pz_trap(params, wf_u16, n_threads=0):
    """BASELINE configs[1] on the CPU: columns blmean, t0, t50, e_trap, e_10410"""
    L = lib()
    L.orc_pz_trap.argtypes = [C.POINTER(_abi.IcpcParams), C.c_void_p, C.c_int64, C.c_int64, _dp, C.c_int]
    wf = np.ascontiguousarray(wf_u16, dtype=np.uint16)
    out = np.zeros((wf.shape[0], 5), dtype=np.float64)
    L.orc_pz_trap(C.byref(params), wf.ctypes.data, wf.shape[0], wf.shape[1], out.ctypes.data_as(_dp), int(n_threads))
    return out
