"""`dsp_icpc(data, config, τ, pars_filter)` -- host-side mirror of the reference entry point
(/root/reference/src/dsp_icpc.jl:62-230) on top of the C ABI (include/lgdsp_b200.h).

Same argument meaning, same output column names/order/units as the reference's TypedTables.Table
(src/dsp_icpc.jl:210-229).  All arithmetic runs in the CUDA library; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from dataclasses import dataclass
from typing import Any, Dict, Mapping, Optional

import numpy as np

from . import _abi
from ._lib import Handle
from .config import (DSPConfig, Q, RddspPolicy, DEFAULT_POLICY, ns, resolve_icpc_params,
                     resolve_compressed_params)


@dataclass
class RDWaveforms:
    """Minimal stand-in for `ArrayOfRDWaveforms` with a shared, uniformly sampled time axis:
    `signal[n_events, n_samples]` (each waveform contiguous, as `flatview(wvfs.signal)` of the LH5 column) and
    the time axis of the first waveform (the reference only consults `wvfs[1].time`, src/dsp_icpc.jl:88,90)."""
    signal: Any
    t_first: Q = ns(0.0)
    step: Q = ns(16.0)

    def __len__(self):
        return int(self.signal.shape[0])


# reference column order, src/dsp_icpc.jl:210-229 (pass-through columns included)
TABLE_COLUMNS = (
    "blmean", "blsigma", "blslope", "bloffset", "tailmean", "tailsigma", "tailslope", "tailoffset", "qc_label",
    "t0", "t10", "t50", "t80", "t90", "t99", "t50_current", "drift_time", "tail_τ", "tail_mean", "tail_sigma",
    "e_max", "e_min", "e_10410", "e_535", "e_313", "e_10410_inv", "e_313_inv", "t0_inv", "e_trap", "e_cusp", "e_zac",
    "e_trap_max", "e_cusp_max", "e_zac_max", "t_trap_max", "t_cusp_max", "t_zac_max", "qdrift", "lq",
    "a_sg", "a_60", "a_100", "a_raw", "blfc", "timestamp", "eventID_fadc", "e_fc",
    "inTrace_intersect", "inTrace_n", "n_sat_low", "n_sat_high", "n_sat_low_cons", "n_sat_high_cons",
)
_PASS_THROUGH = {"blfc": "baseline", "timestamp": "timestamp", "eventID_fadc": "eventnumber", "e_fc": "daqenergy"}
_RENAME = {"tail_τ": "tail_tau"}

_handles: Dict[int, Handle] = {}


def get_handle(device: int = 0) -> Handle:
    """process-wide handle per device (created on first use; raises without a CUDA device)"""
    h = _handles.get(device)
    if h is None or h._h is None:
        h = Handle(device)
        _handles[device] = h
    return h


def _as_waveforms(wvfs) -> RDWaveforms:
    if isinstance(wvfs, RDWaveforms):
        return wvfs
    if hasattr(wvfs, "signal") and hasattr(wvfs, "step"):
        return RDWaveforms(wvfs.signal, getattr(wvfs, "t_first", ns(0.0)), wvfs.step)
    return RDWaveforms(wvfs)


def _signal_u16(sig) -> np.ndarray:
    a = np.asarray(sig)
    if a.ndim != 2:
        raise ValueError("waveform signals must be a 2-D array [n_events, n_samples]")
    if a.dtype != np.uint16:
        if not np.issubdtype(a.dtype, np.integer):
            raise TypeError("this implementation processes raw ADC samples (UInt16); got dtype %s" % a.dtype)
        if a.size and (a.min() < 0 or a.max() > 65535):
            raise ValueError("samples outside the UInt16 range")
        a = a.astype(np.uint16)
    if a.strides[1] != 2:
        a = np.ascontiguousarray(a)
    return a


def rows_to_table(rows: np.ndarray, data: Optional[Mapping[str, Any]] = None) -> "OrderedDict[str, np.ndarray]":
    """double[n_events, 49] device rows -> reference-ordered table of columns"""
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    n = rows.shape[0]
    for name in TABLE_COLUMNS:
        if name in _PASS_THROUGH:
            src = _PASS_THROUGH[name]
            if data is not None and src in data:
                out[name] = np.asarray(data[src])
            continue
        col = rows[:, _abi.COL[_RENAME.get(name, name)]]
        if _RENAME.get(name, name) in _abi.INT_COLUMNS:
            out[name] = col.astype(np.int64)
        else:
            out[name] = np.ascontiguousarray(col)
    assert all(len(v) == n for v in out.values())
    return out


def dsp_icpc(data: Mapping[str, Any], config: DSPConfig, τ: Q, pars_filter: Optional[Dict[str, Any]] = None, *,
             f_evaluate_qc=None, device: int = 0, handle: Optional[Handle] = None,
             policy: RddspPolicy = DEFAULT_POLICY, groups: int = _abi.GROUP_ALL,
             cuspzac_direct: bool = False) -> "OrderedDict[str, np.ndarray]":
    """DSP routine for ICPC detectors: the reference's `dsp_icpc(data, config, τ, pars_filter)`.

    `data` needs the columns `waveform` (RDWaveforms or a UInt16 array [n_events, n_samples]) and, for the
    pass-through columns, `baseline`, `timestamp`, `eventnumber`, `daqenergy`.
    Returns an ordered mapping column name -> numpy array (times in µs / ns as in the reference, see _abi.UNITS).
    """
    if f_evaluate_qc is not None:
        # src/dsp_icpc.jl:108: the ML quality-cut classifier (LIBSVM) is outside the hot path
        raise NotImplementedError("f_evaluate_qc is not supported; qc_label is -1 as in the reference without a model")
    w = _as_waveforms(data["waveform"])
    sig = _signal_u16(w.signal)
    n_events, n_samples = sig.shape
    P = resolve_icpc_params(config, τ, pars_filter, n_samples=n_samples, t_first=w.t_first, step=w.step,
                            groups=groups, policy=policy, cuspzac_direct=cuspzac_direct)
    h = handle or get_handle(device)
    rows = np.zeros((n_events, _abi.NCOL), dtype=np.float64)
    h.icpc_run_host(P, sig.ctypes.data, n_events, sig.strides[0] // 2, rows.ctypes.data)
    return rows_to_table(rows, data)


def dsp_icpc_rows(signal_u16: np.ndarray, params: _abi.IcpcParams, *, device: int = 0,
                  handle: Optional[Handle] = None) -> np.ndarray:
    """lower-level entry used by the tests: resolved params in, raw rows out"""
    sig = _signal_u16(signal_u16)
    h = handle or get_handle(device)
    rows = np.zeros((sig.shape[0], _abi.NCOL), dtype=np.float64)
    h.icpc_run_host(params, sig.ctypes.data, sig.shape[0], sig.strides[0] // 2, rows.ctypes.data)
    return rows


# ----------------------------------------------------------------------------------------------
# dsp_icpc_compressed   (/root/reference/src/dsp_icpc.jl:293-499)
# ----------------------------------------------------------------------------------------------
# reference column order :463-499; value = (source, name) with source "pass" (input column), "pre" / "wdw" (row of the
# fused chain on the presummed / windowed waveform), "stat" ((window, field) of the statistics block)
_ST = {"mean": 0, "sigma": 1, "slope": 2, "offset": 3, "slope_sigma": 4}
_WIN = {"auxbl1": 0, "auxbl2": 1, "bl": 2, "auxpz1": 3, "auxpz2": 4}


def _aux(win):
    return [(f"{win}_mean", ("stat", (win, "mean"))), (f"{win}_sigma", ("stat", (win, "sigma"))),
            (f"{win}_slope_sigma", ("stat", (win, "slope_sigma")))]


COMPRESSED_COLUMNS = OrderedDict(
    [("blfc", ("pass", "baseline")), ("timestamp", ("pass", "timestamp")), ("eventID_fadc", ("pass", "eventnumber")),
     ("e_fc", ("pass", "daqenergy")), ("deadtime", ("pass", "deadtime")),
     ("n_sat_low", ("pre", "n_sat_low")), ("n_sat_high", ("pre", "n_sat_high")),
     ("n_sat_low_cons", ("pre", "n_sat_low_cons")), ("n_sat_high_cons", ("pre", "n_sat_high_cons")),
     ("t_sat_lo", ("pass", "t_sat_lo")), ("t_sat_hi", ("pass", "t_sat_hi")),
     ("blmean", ("pre", "blmean")), ("blsigma", ("pre", "blsigma")), ("blslope", ("pre", "blslope")),
     ("bloffset", ("pre", "bloffset")), ("bl_slope_sigma", ("stat", ("bl", "slope_sigma")))]
    + _aux("auxbl1") + _aux("auxbl2")
    + [("qc_label", ("pre", "qc_label")),
       ("e_max", ("wdw", "e_max")), ("e_min", ("wdw", "e_min")), ("e_max_pre", ("pre", "e_max")), ("e_min_pre", ("pre", "e_min")),
       ("tailmean", ("pre", "tailmean")), ("tailsigma", ("pre", "tailsigma")), ("tailslope", ("pre", "tailslope")),
       ("tailoffset", ("pre", "tailoffset")), ("tail_τ", ("pre", "tail_tau")), ("tail_mean", ("pre", "tail_mean")),
       ("tail_sigma", ("pre", "tail_sigma"))]
    + _aux("auxpz1") + _aux("auxpz2")
    + [("t0", ("wdw", "t0")), ("t10", ("wdw", "t10")), ("t50", ("wdw", "t50")), ("t80", ("wdw", "t80")),
       ("t90", ("wdw", "t90")), ("t99", ("wdw", "t99")), ("t50_pre", ("pre", "t50")),
       ("drift_time", ("wdw", "drift_time")), ("t50_current", ("pre", "t50_current")),
       ("e_10410", ("pre", "e_10410")), ("e_535", ("pre", "e_535")), ("e_313", ("pre", "e_313")),
       ("e_trap", ("pre", "e_trap")), ("e_cusp", ("pre", "e_cusp")), ("e_zac", ("pre", "e_zac")),
       ("e_trap_max", ("pre", "e_trap_max")), ("e_cusp_max", ("pre", "e_cusp_max")), ("e_zac_max", ("pre", "e_zac_max")),
       ("t_trap_max", ("pre", "t_trap_max")), ("t_cusp_max", ("pre", "t_cusp_max")), ("t_zac_max", ("pre", "t_zac_max")),
       ("qdrift", ("wdw", "qdrift")), ("lq", ("wdw", "lq")),
       ("a_sg", ("wdw", "a_sg")), ("a_60", ("wdw", "a_60")), ("a_100", ("wdw", "a_100")), ("a_raw", ("wdw", "a_raw")),
       ("inTrace_intersect", ("pre", "inTrace_intersect")), ("inTrace_n", ("pre", "inTrace_n")),
       ("e_10410_inv", ("pre", "e_10410_inv")), ("e_313_inv", ("pre", "e_313_inv")), ("t0_inv", ("wdw", "t0_inv"))])


def _signal_uint(sig) -> np.ndarray:
    """raw samples as uint16 when they fit, else uint32 (presummed traces: 16-bit ADC x presum rate)"""
    a = np.asarray(sig)
    if a.ndim != 2:
        raise ValueError("waveform signals must be a 2-D array [n_events, n_samples]")
    if not np.issubdtype(a.dtype, np.integer):
        raise TypeError("this implementation processes raw integer ADC samples; got dtype %s" % a.dtype)
    if a.dtype not in (np.uint16, np.uint32):
        if a.size and a.min() < 0:
            raise ValueError("negative samples")
        if a.size and a.max() > 0xFFFFFFFF:
            raise ValueError("samples outside the UInt32 range")
        a = a.astype(np.uint16 if (a.size == 0 or a.max() <= 65535) else np.uint32)
    if a.strides[1] != a.dtype.itemsize:
        a = np.ascontiguousarray(a)
    return a


def compressed_to_table(rows_pre: np.ndarray, rows_wdw: np.ndarray, stats: np.ndarray,
                        data: Optional[Mapping[str, Any]] = None) -> "OrderedDict[str, np.ndarray]":
    """assemble the reference's result table (:463-499) from the two row blocks and the statistics block"""
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, (src, key) in COMPRESSED_COLUMNS.items():
        if src == "pass":
            if data is not None and key in data:
                out[name] = np.asarray(data[key])
            continue
        if src == "stat":
            out[name] = np.ascontiguousarray(stats[:, _WIN[key[0]], _ST[key[1]]])
            continue
        col = (rows_pre if src == "pre" else rows_wdw)[:, _abi.COL[key]]
        out[name] = col.astype(np.int64) if key in _abi.INT_COLUMNS else np.ascontiguousarray(col)
    return out


def _dsp_icpc_compressed_encoded(data, wp, ww, config, τ, pars_filter, device, handle, policy, builders):
    ep, ew = wp.signal, ww.signal
    if len(ep) != len(ew):
        raise ValueError("waveform_presummed and waveform_windowed differ in length")
    rates = np.unique(np.asarray(data["presum_rate"]))
    if rates.size != 1:
        raise ValueError("presum_rate must be the same for all events")
    presum = int(rates[0])
    n_events = len(ep)
    P_pre, P_wdw, aux = resolve_compressed_params(config, τ, pars_filter, presum_rate=presum, n_pre=ep.n_samples,
                                                  t_first_pre=wp.t_first, step_pre=wp.step, n_wdw=ew.n_samples,
                                                  t_first_wdw=ww.t_first, step_wdw=ww.step, policy=policy, builders=builders)
    h = handle or get_handle(device)
    rows_pre = np.zeros((n_events, _abi.NCOL), dtype=np.float64)
    rows_wdw = np.zeros((n_events, _abi.NCOL), dtype=np.float64)
    stats = np.zeros((n_events, 5, 5), dtype=np.float64)
    h.icpc_compressed_run_encoded_host(P_pre, P_wdw, ep, ew, float(presum), aux, rows_pre.ctypes.data, rows_wdw.ctypes.data,
                                       stats.ctypes.data)
    return compressed_to_table(rows_pre, rows_wdw, stats, data)


def dsp_icpc_compressed(data: Mapping[str, Any], config: DSPConfig, τ: Q, pars_filter: Optional[Dict[str, Any]] = None, *,
                        f_evaluate_qc=None, device: int = 0, handle: Optional[Handle] = None,
                        policy: RddspPolicy = DEFAULT_POLICY, builders=None) -> "OrderedDict[str, np.ndarray]":
    """DSP routine for ICPC detectors on the compressed format: the reference's
    `dsp_icpc_compressed(data, config, τ, pars_filter)` (src/dsp_icpc.jl:293-499).

    `data` needs `waveform_presummed`, `waveform_windowed` (RDWaveforms of raw integer samples -- or of
    `codec.EncodedWaveforms`, which are decoded on the device like the reference's `decode_data` calls :313-314 -- each with
    its own time axis) and `presum_rate` (one value for all events, `only(unique(presum_rate))` :324); the pass-through columns
    `baseline, timestamp, eventnumber, daqenergy, t_sat_lo, t_sat_hi, deadtime` are copied when present."""
    if f_evaluate_qc is not None:
        raise NotImplementedError("f_evaluate_qc is not supported; qc_label is -1 as in the reference without a model")
    from .codec import EncodedWaveforms
    wp, ww = _as_waveforms(data["waveform_presummed"]), _as_waveforms(data["waveform_windowed"])
    encoded = isinstance(wp.signal, EncodedWaveforms) and isinstance(ww.signal, EncodedWaveforms)
    if encoded:
        # decode_data(data.waveform_presummed), decode_data(data.waveform_windowed)  src/dsp_icpc.jl:313-314 -- on the device
        return _dsp_icpc_compressed_encoded(data, wp, ww, config, τ, pars_filter, device, handle, policy, builders)
    if isinstance(wp.signal, EncodedWaveforms) or isinstance(ww.signal, EncodedWaveforms):
        from .codec import decode_data
        wp = RDWaveforms(decode_data(wp.signal, handle or get_handle(device)), wp.t_first, wp.step) if isinstance(wp.signal, EncodedWaveforms) else wp
        ww = RDWaveforms(decode_data(ww.signal, handle or get_handle(device)), ww.t_first, ww.step) if isinstance(ww.signal, EncodedWaveforms) else ww
    pre, wdw = _signal_uint(wp.signal), _signal_uint(ww.signal)
    if pre.shape[0] != wdw.shape[0]:
        raise ValueError("waveform_presummed and waveform_windowed differ in length")
    rates = np.unique(np.asarray(data["presum_rate"]))
    if rates.size != 1:                                           # only(unique(presum_rate))  :324
        raise ValueError("presum_rate must be the same for all events")
    presum = int(rates[0])
    n_events = pre.shape[0]
    P_pre, P_wdw, aux = resolve_compressed_params(config, τ, pars_filter, presum_rate=presum, n_pre=pre.shape[1],
                                                  t_first_pre=wp.t_first, step_pre=wp.step, n_wdw=wdw.shape[1],
                                                  t_first_wdw=ww.t_first, step_wdw=ww.step, policy=policy, builders=builders)
    h = handle or get_handle(device)
    rows_pre = np.zeros((n_events, _abi.NCOL), dtype=np.float64)
    rows_wdw = np.zeros((n_events, _abi.NCOL), dtype=np.float64)
    stats = np.zeros((n_events, 5, 5), dtype=np.float64)
    h.icpc_compressed_run_host(P_pre, P_wdw, pre.ctypes.data, pre.dtype.itemsize, pre.strides[0] // pre.dtype.itemsize,
                               wdw.ctypes.data, wdw.dtype.itemsize, wdw.strides[0] // wdw.dtype.itemsize, float(presum), aux,
                               n_events, rows_pre.ctypes.data, rows_wdw.ctypes.data, stats.ctypes.data)
    return compressed_to_table(rows_pre, rows_wdw, stats, data)
