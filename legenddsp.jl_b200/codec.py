"""`decode_data` -- host-side mirror of LegendDataTypes.jl's waveform codecs for the dsp_* entry points
(call sites /root/reference/src/dsp_icpc.jl:313-314, src/dsp_puls.jl:103, src/dsp_sipm.jl:241).

An encoded waveform set is the reference's `VectorOfEncodedArrays`: one byte buffer plus element pointers
(`EncodedWaveforms.data`, `.offsets`).  Decoding runs on the GPU (csrc/lgdsp_codec.cu); the encoders are host
utilities for tests, benchmarks and round trips.  No CPU decode path exists in this package.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import Handle, load_library

RADWARE_SIGCOMPRESS = 1      # RadwareSigcompress(shift): 16-bit samples, shift = -32768 for UInt16 waveforms
ULEB128_ZIGZAG_DIFF = 2      # ULEB128 zig-zag difference codec: 16- / 32-bit samples


@dataclass
class EncodedWaveforms:
    """`VectorOfEncodedArrays` of equally long waveforms: event e is data[offsets[e]:offsets[e+1]]"""
    codec: int
    data: np.ndarray          # uint8
    offsets: np.ndarray       # int64[n_events + 1]
    n_samples: int
    sample_bytes: int = 2
    shift: int = -32768       # RadwareSigcompress only

    def __len__(self):
        return int(self.offsets.size - 1)

    @property
    def nbytes(self):
        return int(self.offsets[-1] - self.offsets[0])


def encode_waveforms(wf, codec=RADWARE_SIGCOMPRESS, shift=None) -> EncodedWaveforms:
    """host-side encoder: wf[n_events, n_samples] of uint16 (either codec) or uint32 (ULEB128_ZIGZAG_DIFF)"""
    a = np.ascontiguousarray(wf)
    if a.ndim != 2 or a.dtype not in (np.uint16, np.uint32):
        raise TypeError("waveforms must be a 2-D uint16 / uint32 array")
    sb = a.dtype.itemsize
    if codec == RADWARE_SIGCOMPRESS and sb != 2:
        raise TypeError("RadwareSigcompress holds 16-bit samples")
    if shift is None:
        shift = -32768 if codec == RADWARE_SIGCOMPRESS else 0
    L = load_library()
    ne, n = a.shape
    cap = int(L.lgdsp_codec_max_encoded_bytes(codec, n, sb)) * max(ne, 1) + 16
    buf = np.empty(cap, dtype=np.uint8)
    off = np.zeros(ne + 1, dtype=np.int64)
    rc = L.lgdsp_codec_encode_host(codec, C.c_void_p(a.ctypes.data), sb, ne, n, n, int(shift), C.c_void_p(buf.ctypes.data), cap,
                                   C.c_void_p(off.ctypes.data))
    if rc != 0:
        raise ValueError(f"encode failed (code {rc}): samples outside the codec's range?")
    return EncodedWaveforms(codec, buf[:int(off[-1])].copy(), off, n, sb, int(shift))


def decode_data(enc: EncodedWaveforms, handle: Handle = None) -> np.ndarray:
    """decode_data(encoded waveforms) -> wf[n_events, n_samples] (uint16 / uint32), decoded on the GPU"""
    from .dsp_icpc import get_handle
    h = handle or get_handle()
    ne = len(enc)
    out = np.empty((ne, enc.n_samples), dtype=np.uint16 if enc.sample_bytes == 2 else np.uint32)
    data = np.ascontiguousarray(enc.data, dtype=np.uint8)
    off = np.ascontiguousarray(enc.offsets, dtype=np.int64)
    h.decode_data_host(enc.codec, data.ctypes.data, off.ctypes.data, ne, enc.n_samples, enc.shift, out.ctypes.data, enc.sample_bytes,
                       enc.n_samples)
    return out
