"""`dsp_puls(data, config)` (/root/reference/src/dsp_puls.jl:29-66) and `dsp_decay_times(wvfs, config)`
(/root/reference/src/dsp_decaytime.jl:11-26): strict subsets of the `dsp_icpc` chain, served by the same fused kernel
with a reduced column-group mask.

* dsp_puls works on the baseline-subtracted waveform WITHOUT pole-zero correction (src/dsp_puls.jl:44-58): the kernel
  runs with pz_km1 = 0 (InvCRFilter becomes the identity), t50 at half the maximum with get_threshold's default
  mintot = 1000 ns (src/dsp_routines.jl:33), e_10410 = maximum of TrapezoidalChargeFilter(10 us, 4 us).
* dsp_decay_times = signalstats on bl_window, shift, tailstats on tail_window -> tau in us.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Mapping, Optional

import numpy as np

from . import _abi
from ._lib import Handle
from .config import DSPConfig, Q, RddspPolicy, DEFAULT_POLICY, _min_n, example_config_dict, ns, resolve_icpc_params, us
from .dsp_icpc import _as_waveforms, _signal_u16, get_handle

PULS_COLUMNS = ("blmean", "blsigma", "blslope", "bloffset", "t50", "e_max", "e_10410", "blfc", "timestamp", "eventID_fadc", "e_fc")
_PASS = {"blfc": "baseline", "timestamp": "timestamp", "eventID_fadc": "eventnumber", "e_fc": "daqenergy"}


def resolve_puls_params(config: DSPConfig, *, n_samples: int = 8192, t_first: Q = ns(0.0), step: Q = ns(16.0),
                        policy: RddspPolicy = DEFAULT_POLICY, builders=None) -> _abi.IcpcParams:
    """dsp_icpc parameters specialised to the pulser chain: no pole-zero correction, default get_threshold mintot"""
    P = resolve_icpc_params(config, us(500.0), None, n_samples=n_samples, t_first=t_first, step=step,
                            role="puls", policy=policy, builders=builders)
    P.pz_km1 = 0.0                               # no InvCRFilter in dsp_puls
    P.tx_min_n = _min_n(ns(1000.0), step)        # get_threshold(wvfs, thr) default mintot  (src/dsp_routines.jl:33)
    return P


def dsp_puls(data: Mapping[str, Any], config: DSPConfig, *, device: int = 0, handle: Optional[Handle] = None,
             policy: RddspPolicy = DEFAULT_POLICY) -> "OrderedDict[str, np.ndarray]":
    """DSP function for pulser processing: the reference's `dsp_puls(data, config)` (same columns, t50 in us)"""
    w = _as_waveforms(data["waveform"])
    sig = _signal_u16(w.signal)
    P = resolve_puls_params(config, n_samples=sig.shape[1], t_first=w.t_first, step=w.step, policy=policy)
    h = handle or get_handle(device)
    rows = np.zeros((sig.shape[0], _abi.NCOL), dtype=np.float64)
    h.icpc_run_host(P, sig.ctypes.data, sig.shape[0], sig.strides[0] // 2, rows.ctypes.data)
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name in PULS_COLUMNS:
        if name in _PASS:
            if _PASS[name] in data:
                out[name] = np.asarray(data[_PASS[name]])
        else:
            out[name] = np.ascontiguousarray(rows[:, _abi.COL[name]])
    return out


def dsp_puls_compressed(data: Mapping[str, Any], config: DSPConfig, *, device: int = 0, handle: Optional[Handle] = None,
                        policy: RddspPolicy = DEFAULT_POLICY) -> "OrderedDict[str, np.ndarray]":
    """`dsp_puls_compressed(data, config)` (src/dsp_puls.jl:98-134): `dsp_puls` on `decode_data(data.waveform_presummed)`.
    The codec (LegendDataTypes.decode_data) is outside the reference tree: `data["waveform_presummed"]` holds the DECODED
    integer samples (uint16, or uint32 sums of at most 4096 samples per waveform)."""
    w = _as_waveforms(data["waveform_presummed"])
    sig = np.asarray(w.signal)
    if sig.dtype == np.uint32 or (sig.dtype.kind in "iu" and sig.dtype.itemsize > 2 and sig.size and int(sig.max()) > 65535):
        sig = np.ascontiguousarray(sig, dtype=np.uint32)
        sample_bytes = 4
    else:
        sig = _signal_u16(sig)
        sample_bytes = 2
    P = resolve_puls_params(config, n_samples=sig.shape[1], t_first=w.t_first, step=w.step, policy=policy)
    h = handle or get_handle(device)
    rows = np.zeros((sig.shape[0], _abi.NCOL), dtype=np.float64)
    h.icpc_run_ext_host(P, sig.ctypes.data, sample_bytes, None, sig.shape[0], sig.strides[0] // sample_bytes, rows.ctypes.data)
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name in PULS_COLUMNS:
        if name in _PASS:
            if _PASS[name] in data:
                out[name] = np.asarray(data[_PASS[name]])
        else:
            out[name] = np.ascontiguousarray(rows[:, _abi.COL[name]])
    return out


def dsp_decay_times(wvfs, config_or_bl_window, tail_window=None, *, device: int = 0, handle: Optional[Handle] = None,
                    policy: RddspPolicy = DEFAULT_POLICY) -> np.ndarray:
    """decay time of the waveform tails in us: `dsp_decay_times(wvfs, config)` or
    `dsp_decay_times(wvfs, bl_window, tail_window)` with windows as (Q, Q) pairs; 0 where a tail sample is <= 0
    (src/tailstats.jl:27-33)"""
    if isinstance(config_or_bl_window, DSPConfig):
        cfg = config_or_bl_window
    else:
        d = example_config_dict()
        d["bl_window"] = {"min": config_or_bl_window[0], "max": config_or_bl_window[1]}
        d["tail_window"] = {"min": tail_window[0], "max": tail_window[1]}
        cfg = DSPConfig.from_dict(d)
    w = _as_waveforms(wvfs)
    sig = _signal_u16(w.signal)
    # role "decay": only the baseline and tail windows have to fit the trace (the reference's
    # dsp_decay_times(wvfs, bl_window, tail_window) takes nothing else)
    P = resolve_icpc_params(cfg, us(500.0), None, n_samples=sig.shape[1], t_first=w.t_first, step=w.step,
                            role="decay", policy=policy)
    h = handle or get_handle(device)
    rows = np.zeros((sig.shape[0], _abi.NCOL), dtype=np.float64)
    h.icpc_run_host(P, sig.ctypes.data, sig.shape[0], sig.strides[0] // 2, rows.ctypes.data)
    return rows[:, _abi.COL["tail_tau"]] * 1e-3     # ns -> us  (uconvert.(u"µs", decay_times.τ))
