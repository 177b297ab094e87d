"""Synthetic ICPC waveforms: host generator (numpy buffer) and device generator."""
import ctypes as C

import numpy as np

from . import _abi
from ._lib import load_library

DEFAULT_SEED = 20260101


def synth_params(n_samples=8192, seed=DEFAULT_SEED, mode=0, noise_sigma=3.0, tau_samples=31250.0) -> _abi.SynthParams:
    return _abi.SynthParams(seed, n_samples, mode, noise_sigma, tau_samples)


def generate_host(n_events, first_event=0, **kw) -> np.ndarray:
    """events [first_event, first_event+n_events) as uint16[n_events, n_samples] (pure CPU, same stream as the GPU)"""
    sp = synth_params(**kw)
    out = np.zeros((int(n_events), sp.n_samples), dtype=np.uint16)
    rc = load_library().lgdsp_synth_generate_host(C.byref(sp), int(first_event), int(n_events), sp.n_samples,
                                                  C.c_void_p(out.ctypes.data))
    if rc != 0:
        raise ValueError(f"lgdsp_synth_generate_host failed: {rc}")
    return out


def generate_device(handle, d_ptr, n_events, first_event=0, ld=None, **kw):
    sp = synth_params(**kw)
    handle.synth_device(sp, first_event, n_events, ld or sp.n_samples, d_ptr)
    return sp


def compress(wf_u16: np.ndarray, presum_rate: int = 8, window=(2600, 1400)):
    """the DAQ's compressed format from full-rate traces: (presummed, windowed) where presummed[e, k] is the sum of
    `presum_rate` consecutive samples (uint32 when the sums exceed 16 bits; length n // presum_rate) and windowed is
    the full-rate slice [from, from + length) (what dsp_icpc_compressed reads as waveform_presummed /
    waveform_windowed, /root/reference/src/dsp_icpc.jl:313-314).  Input generation only."""
    wf = np.asarray(wf_u16)
    n_ev, n = wf.shape
    r = int(presum_rate)
    m = n // r
    pre = wf[:, :m * r].astype(np.uint32).reshape(n_ev, m, r).sum(axis=2, dtype=np.uint64)
    pre = pre.astype(np.uint16 if (pre.size == 0 or pre.max() <= 65535) else np.uint32)
    a, length = int(window[0]), int(window[1])
    wdw = np.ascontiguousarray(wf[:, a:a + length])
    return pre, wdw
