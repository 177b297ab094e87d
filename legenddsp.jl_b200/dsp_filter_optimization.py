"""Trapezoidal filter-optimisation sweeps: host-side mirror of `dsp_trap_rt_optimization`
(/root/reference/src/dsp_filter_optimization.jl:102-133) and `dsp_trap_ft_optimization` (:241-274), plus the
batched (rt x ft) grid that callers of the reference build by looping over the ft sweep (BASELINE.json config 4).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import _abi
from ._lib import Handle
from .config import DSPConfig, Q, grid_values, resolve_sweep_params, trap_variants, us
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
    var = trap_variants(rts, [ft], w.step, mode="rt", pickoff=config.enc_pickoff_trap)
    out = _run(w, config, τ, var, device, handle)
    return np.ascontiguousarray(out.T).astype(np.float64)


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
