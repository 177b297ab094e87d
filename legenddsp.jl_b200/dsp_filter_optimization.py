"""Filter-optimisation sweeps: host-side mirror of /root/reference/src/dsp_filter_optimization.jl --
`dsp_trap_rt_optimization` (:102-133), `dsp_cusp_rt_optimization` (:145-182), `dsp_zac_rt_optimization` (:193-231),
`dsp_trap_ft_optimization` (:241-274), `dsp_cusp_ft_optimization` (:286-325), `dsp_zac_ft_optimization` (:336-375),
`dsp_sg_optimization` (:393-441), plus the batched (rt x ft) grid that callers of the reference build by looping over
the ft sweep (BASELINE.json config 4), `dsp_qc_flt_optimization` without a classifier (:9-70) and
`dsp_qdrift_flt_optimization` (:72-91).  The `_compressed` sweeps and the ML classifier are out of scope.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import _abi
from ._lib import Handle
from collections import OrderedDict

from .config import (DSPConfig, Q, RddspPolicy, DEFAULT_POLICY, cuspzac_sweep_variants, get_fltpars, grid_values,
                     resolve_sweep_params, sg_sweep_variants, trap_sweep_variants, trap_variants, us)
from .dsp_icpc import _as_waveforms, _signal_u16, get_handle


def _run(wvfs, config: DSPConfig, τ: Q, variants, device, handle: Optional[Handle]) -> np.ndarray:
    w = _as_waveforms(wvfs)
    sig = _signal_u16(w.signal)
    n_events, n_samples = sig.shape
    S = resolve_sweep_params(config, τ, n_samples=n_samples, t_first=w.t_first, step=w.step)
    h = handle or get_handle(device)
    out = np.zeros((n_events, len(variants)), dtype=np.float32)
    h.sweep_run_host(S, sig.ctypes.data, n_events, sig.strides[0] // 2, variants, out.ctypes.data)
    return out


def dsp_trap_rt_optimization(wvfs, config: DSPConfig, τ: Q, *, ft: Q = us(2.0), device: int = 0,
                             handle: Optional[Handle] = None) -> np.ndarray:
    """ENC noise grid for the trap rise-time grid at fixed flat-top `ft`, pick-off `enc_pickoff_trap`.
    Returns Float64[n_rt, n_events] like the reference (src/dsp_filter_optimization.jl:122)."""
    w = _as_waveforms(wvfs)
    rts = grid_values(config.e_grid_rt_trap)
    # (the general sweep entry in float64: the reference's grid is a Matrix{Float64}, :122)
    var = trap_sweep_variants(rts, [ft], w.step, mode="rt", pickoff=config.enc_pickoff_trap)
    out = _run_general(w, config, τ, var, f64=True, device=device, handle=handle)
    return np.ascontiguousarray(out.T)


def dsp_trap_ft_optimization(wvfs, config: DSPConfig, τ: Q, rt: Q, *, device: int = 0,
                             handle: Optional[Handle] = None) -> np.ndarray:
    """Energy grid for the trap flat-top grid at fixed rise time `rt`, pick-off t50 + rt + ft/2.
    Returns Float32[n_ft, n_events] like the reference (src/dsp_filter_optimization.jl:263)."""
    w = _as_waveforms(wvfs)
    fts = grid_values(config.e_grid_ft_trap)
    var = trap_variants([rt], fts, w.step, mode="ft")
    out = _run(w, config, τ, var, device, handle)
    return np.ascontiguousarray(out.T)


def dsp_trap_rtft_grid(wvfs, config: DSPConfig, τ: Q, rts: Optional[Sequence[Q]] = None,
                       fts: Optional[Sequence[Q]] = None, *, device: int = 0,
                       handle: Optional[Handle] = None) -> np.ndarray:
    """All (rt, ft) variants of the ft sweep in ONE pass over the waveforms: Float32[n_rt, n_ft, n_events]
    (equals stacking dsp_trap_ft_optimization(wvfs, config, τ, rt) over rt)."""
    w = _as_waveforms(wvfs)
    rts = list(rts) if rts is not None else grid_values(config.e_grid_rt_trap)
    fts = list(fts) if fts is not None else grid_values(config.e_grid_ft_trap)
    var = trap_variants(rts, fts, w.step, mode="ft")
    out = _run(w, config, τ, var, device, handle)
    return np.ascontiguousarray(out.T).reshape(len(rts), len(fts), -1)


# ----------------------------------------------------------------------------------------------
# general sweeps (lgdsp_sweep_run): CUSP / ZAC rise- and flat-top-time sweeps, Savitzky-Golay window-length sweep
# ----------------------------------------------------------------------------------------------
def _run_general(wvfs, config: DSPConfig, τ: Q, variants, *, f64: bool, want_aux: bool = False, device: int = 0,
                 handle: Optional[Handle] = None):
    w = _as_waveforms(wvfs)
    sig = _signal_u16(w.signal)
    n_events, n_samples = sig.shape
    S = resolve_sweep_params(config, τ, n_samples=n_samples, t_first=w.t_first, step=w.step, out_f64=f64)
    h = handle or get_handle(device)
    out = np.zeros((n_events, len(variants)), dtype=np.float64 if f64 else np.float32)
    aux = np.zeros((n_events, 4), dtype=np.float64) if want_aux else None
    h.gsweep_run_host(S, sig.ctypes.data, n_events, sig.strides[0] // 2, variants.array, out.ctypes.data,
                      aux.ctypes.data if want_aux else None)
    return (out, aux) if want_aux else out


def _cuspzac_rt(kind, wvfs, config, τ, ft, policy, device, handle):
    w = _as_waveforms(wvfs)
    rts = grid_values(config.e_grid_rt_cusp if kind == "cusp" else config.e_grid_rt_zac)
    var = cuspzac_sweep_variants(config, kind, rts, [ft], w.step, mode="rt", policy=policy)
    out = _run_general(w, config, τ, var, f64=True, device=device, handle=handle)
    return np.ascontiguousarray(out.T)


def _cuspzac_ft(kind, wvfs, config, τ, rt, policy, device, handle):
    w = _as_waveforms(wvfs)
    fts = grid_values(config.e_grid_ft_cusp if kind == "cusp" else config.e_grid_ft_zac)
    var = cuspzac_sweep_variants(config, kind, [rt], fts, w.step, mode="ft", policy=policy)
    out = _run_general(w, config, τ, var, f64=False, device=device, handle=handle)
    return np.ascontiguousarray(out.T)


def dsp_cusp_rt_optimization(wvfs, config: DSPConfig, τ: Q, *, ft: Q = us(2.0), policy: RddspPolicy = DEFAULT_POLICY,
                             device: int = 0, handle: Optional[Handle] = None) -> np.ndarray:
    """ENC noise grid for the CUSP rise-time grid at fixed flat-top `ft`, pick-off `enc_pickoff_cusp`:
    Float64[n_rt, n_events]  (src/dsp_filter_optimization.jl:145-182)"""
    return _cuspzac_rt("cusp", wvfs, config, τ, ft, policy, device, handle)


def dsp_zac_rt_optimization(wvfs, config: DSPConfig, τ: Q, *, ft: Q = us(2.0), policy: RddspPolicy = DEFAULT_POLICY,
                            device: int = 0, handle: Optional[Handle] = None) -> np.ndarray:
    """same for the ZAC filter  (src/dsp_filter_optimization.jl:193-231)"""
    return _cuspzac_rt("zac", wvfs, config, τ, ft, policy, device, handle)


def dsp_cusp_ft_optimization(wvfs, config: DSPConfig, τ: Q, rt: Q, *, policy: RddspPolicy = DEFAULT_POLICY,
                             device: int = 0, handle: Optional[Handle] = None) -> np.ndarray:
    """energy grid for the CUSP flat-top grid at fixed rise time `rt`, pick-off t50 + flt_length_cusp/2:
    Float32[n_ft, n_events]  (src/dsp_filter_optimization.jl:286-325)"""
    return _cuspzac_ft("cusp", wvfs, config, τ, rt, policy, device, handle)


def dsp_zac_ft_optimization(wvfs, config: DSPConfig, τ: Q, rt: Q, *, policy: RddspPolicy = DEFAULT_POLICY,
                            device: int = 0, handle: Optional[Handle] = None) -> np.ndarray:
    """same for the ZAC filter  (src/dsp_filter_optimization.jl:336-375)"""
    return _cuspzac_ft("zac", wvfs, config, τ, rt, policy, device, handle)


def dsp_sg_optimization(wvfs, config: DSPConfig, τ: Q, pars_filter, *, f_evaluate_qc=None,
                        policy: RddspPolicy = DEFAULT_POLICY, device: int = 0, handle: Optional[Handle] = None):
    """Savitzky-Golay window-length sweep  (src/dsp_filter_optimization.jl:393-441): table with the columns
    aoe[n_events, n_wl] (current maximum / energy), energy (trap(rt, ft) at t50 + rt + ft/2), blmean, blslope [1/ns],
    t50 [us], qc_label (-1)."""
    if f_evaluate_qc is not None:
        raise NotImplementedError("f_evaluate_qc is not supported; qc_label is -1 as in the reference without a model")
    w = _as_waveforms(wvfs)
    sig = _signal_u16(w.signal)
    rt, ft = get_fltpars(pars_filter, "trap", config)       # pars_filter.trap.rt / .ft  (:401-402)
    wls = grid_values(config.a_grid_wl_sg)
    sgv = sg_sweep_variants(config, wls, n_samples=sig.shape[1], t_first=w.t_first, step=w.step, policy=policy)
    ev = trap_sweep_variants([rt], [ft], w.step, mode="ft")
    # one pass: variant 0 = the energy estimate, variants 1.. = the window lengths
    from .config import SweepVariants
    allv = SweepVariants(1 + len(wls))
    allv.array[0] = ev.array[0]
    for i in range(len(wls)):
        allv.array[1 + i] = sgv.array[i]
    allv._keep = sgv._keep
    out, aux = _run_general(w, config, τ, allv, f64=True, want_aux=True, device=device, handle=handle)
    energy = out[:, 0].copy()
    with np.errstate(divide="ignore", invalid="ignore"):
        aoe = out[:, 1:] / energy[:, None]
    return OrderedDict(aoe=np.ascontiguousarray(aoe), energy=energy, blmean=aux[:, 0].copy(), blslope=aux[:, 1].copy(),
                       t50=aux[:, 2].copy(), qc_label=np.full(sig.shape[0], -1, dtype=np.int64))


def _int_samples(sig):
    """decoded integer samples -> (contiguous uint16 | uint32 array, bytes per sample)"""
    a = np.asarray(sig)
    if a.dtype == np.uint32 or (a.dtype.kind in "iu" and a.dtype.itemsize > 2 and a.size and int(a.max()) > 65535):
        return np.ascontiguousarray(a, dtype=np.uint32), 4
    return _signal_u16(a), 2


def dsp_sg_optimization_compressed(wvfs_wdw, wvfs_pre, config: DSPConfig, τ: Q, pars_filter, *, presum_rate: float = 8.0,
                                   f_evaluate_qc=None, policy: RddspPolicy = DEFAULT_POLICY, device: int = 0,
                                   handle: Optional[Handle] = None):
    """Savitzky-Golay window-length sweep on the compressed format (src/dsp_filter_optimization.jl:460-511): baseline,
    pole-zero correction, t50 and the trap(rt, ft) energy from the PRESUMMED waveform (:469-495), the current maxima from the
    WINDOWED waveform shifted by blmean / presum_rate (:477, :498-503).  Same table as `dsp_sg_optimization`
    (aoe, energy, blmean, blslope, t50, qc_label)."""
    if f_evaluate_qc is not None:
        raise NotImplementedError("f_evaluate_qc is not supported; qc_label is -1 as in the reference without a model")
    wp, ww = _as_waveforms(wvfs_pre), _as_waveforms(wvfs_wdw)
    sp, sbp = _int_samples(wp.signal)
    sw, sbw = _int_samples(ww.signal)
    if sp.shape[0] != sw.shape[0]:
        raise ValueError("presummed and windowed waveforms: different event counts")
    n_events = sp.shape[0]
    rt, ft = pars_filter["trap"]["rt"], pars_filter["trap"]["ft"]        # :467-468
    wls = grid_values(config.a_grid_wl_sg)
    h = handle or get_handle(device)
    # presummed waveform: one trapezoid variant at t50 + rt + ft/2 and the aux outputs (blmean, blslope, t50)
    S_pre = resolve_sweep_params(config, τ, n_samples=sp.shape[1], t_first=wp.t_first, step=wp.step, out_f64=True)
    ev = trap_sweep_variants([rt], [ft], wp.step, mode="ft")
    energy = np.zeros((n_events, 1), dtype=np.float64)
    aux = np.zeros((n_events, 4), dtype=np.float64)
    h.gsweep_run_ext_host(S_pre, sp.ctypes.data, sbp, None, n_events, sp.strides[0] // sbp, ev.array, energy.ctypes.data,
                          aux.ctypes.data)
    # windowed waveform, shifted by the presummed baseline / presum_rate: the window-length grid
    S_wdw = resolve_sweep_params(config, τ, n_samples=sw.shape[1], t_first=ww.t_first, step=ww.step, out_f64=True,
                                 external_baseline=True)
    sgv = sg_sweep_variants(config, wls, n_samples=sw.shape[1], t_first=ww.t_first, step=ww.step, policy=policy)
    bl = np.ascontiguousarray(aux[:, 0] / float(presum_rate))
    a = np.zeros((n_events, len(wls)), dtype=np.float64)
    h.gsweep_run_ext_host(S_wdw, sw.ctypes.data, sbw, bl.ctypes.data, n_events, sw.strides[0] // sbw, sgv.array, a.ctypes.data, None)
    with np.errstate(divide="ignore", invalid="ignore"):
        aoe = a / energy
    return OrderedDict(aoe=np.ascontiguousarray(aoe), energy=energy[:, 0].copy(), blmean=aux[:, 0].copy(),
                       blslope=aux[:, 1].copy(), t50=aux[:, 2].copy(), qc_label=np.full(n_events, -1, dtype=np.int64))


def dsp_qc_flt_optimization(wvfs, config: DSPConfig, τ: Q, f_evaluate_qc=None, *, device: int = 0,
                            handle: Optional[Handle] = None):
    """QC DSP for the filter optimisation without a classifier (src/dsp_filter_optimization.jl:12-14, 31-70): table with
    energy (trap(rt, ft) of the default filter parameters at t50 + rt + ft/2), blmean, blslope [1/ns], t50 [us],
    qc_label (-1)"""
    if f_evaluate_qc is not None:
        raise NotImplementedError("f_evaluate_qc is not supported; qc_label is -1 as in the reference without a model")
    w = _as_waveforms(wvfs)
    rt, ft = config.default_flt_param["trap"]["rt"], config.default_flt_param["trap"]["ft"]     # :57-58
    var = trap_sweep_variants([rt], [ft], w.step, mode="ft")
    out, aux = _run_general(w, config, τ, var, f64=True, want_aux=True, device=device, handle=handle)
    return OrderedDict(energy=out[:, 0].copy(), blmean=aux[:, 0].copy(), blslope=aux[:, 1].copy(), t50=aux[:, 2].copy(),
                       qc_label=np.full(out.shape[0], -1, dtype=np.int64))


def dsp_qc_flt_optimization_compressed(wvfs, config: DSPConfig, τ: Q, f_evaluate_qc=None, **kw):
    """`dsp_qc_flt_optimization_compressed(wvfs, config, τ, missing)` (src/dsp_filter_optimization.jl:26-28): without a
    classifier the same `_get_dsp_qc_flt_optimization(wvfs, config, τ, nothing)` as the uncompressed entry"""
    return dsp_qc_flt_optimization(wvfs, config, τ, f_evaluate_qc, **kw)


def dsp_qdrift_flt_optimization(wvfs, blmean, config: DSPConfig, τ: Q, *, device: int = 0, handle: Optional[Handle] = None,
                                builders=None) -> np.ndarray:
    """Q-drift for the filter optimisation (src/dsp_filter_optimization.jl:72-90): the waveforms are shifted by the GIVEN
    baseline means, pole-zero corrected, t0 and get_qdrift as in dsp_icpc.  Returns qdrift[n_events]."""
    from .config import resolve_icpc_params
    w = _as_waveforms(wvfs)
    sig = _signal_u16(w.signal)
    bl = np.ascontiguousarray(blmean, dtype=np.float64)
    if bl.shape != (sig.shape[0],):
        raise ValueError("blmean must hold one value per waveform")
    P = resolve_icpc_params(config, τ, None, n_samples=sig.shape[1], t_first=w.t_first, step=w.step,
                            groups=_abi.GROUP_BASE | _abi.GROUP_TIMING | _abi.GROUP_QDRIFT, builders=builders)
    h = handle or get_handle(device)
    rows = np.zeros((sig.shape[0], _abi.NCOL))
    h.icpc_run_ext_host(P, sig.ctypes.data, 2, bl.ctypes.data, sig.shape[0], sig.strides[0] // 2, rows.ctypes.data)
    return np.ascontiguousarray(rows[:, _abi.COL["qdrift"]])
