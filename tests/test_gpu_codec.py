"""decode_data on the GPU (csrc/lgdsp_codec.cu) through the C ABI: bit-exact against the original samples and the oracle's
decoder, malformed streams, and dsp_icpc fed with encoded bytes == dsp_icpc fed with the decoded samples."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mixed(L, n_events, n=8192, seed=0):
    wf = L.synth.generate_host(n_events, first_event=1000 + seed)[:, :n].copy()
    rng = np.random.default_rng(seed)
    wf[0] = rng.integers(0, 65536, n)                    # incompressible
    wf[1] = 31000                                         # constant
    wf[2] = np.where(np.arange(n) % 2 == 0, 0, 65535)     # differences overflow 16 bits
    return wf


@pytest.mark.parametrize("n", [8192, 1400, 48, 130])
def test_radware_decode_matches_original(L, O, handle, n):
    wf = _mixed(L, 300, n, seed=n)
    enc = L.encode_waveforms(wf, L.RADWARE_SIGCOMPRESS)
    got = L.decode_data(enc, handle)
    assert got.dtype == np.uint16 and np.array_equal(got, wf)
    for e in (0, 1, 2, 17):
        assert np.array_equal(O.radware_decode(enc.data[enc.offsets[e]:enc.offsets[e + 1]]), wf[e])
    print("compression ratio", wf.nbytes / enc.nbytes)


@pytest.mark.parametrize("dtype,n", [("uint32", 1024), ("uint16", 2048), ("uint32", 4096)])
def test_uleb_decode_matches_original(L, O, handle, dtype, n):
    full = _mixed(L, 200, 8192, seed=5).astype(np.uint32)
    if dtype == "uint32":
        wf = full.reshape(200, n, 8192 // n).sum(axis=2, dtype=np.uint32)      # presummed waveforms
    else:
        wf = full[:, :n].astype(np.uint16)
    enc = L.encode_waveforms(wf, L.ULEB128_ZIGZAG_DIFF)
    got = L.decode_data(enc, handle)
    assert got.dtype == wf.dtype and np.array_equal(got, wf)
    assert np.array_equal(O.uleb128zzd_decode(enc.data[enc.offsets[3]:enc.offsets[4]], dtype=wf.dtype), wf[3])


def test_malformed_streams_are_errors(L, handle):
    wf = _mixed(L, 8, 1024)
    enc = L.encode_waveforms(wf, L.RADWARE_SIGCOMPRESS)
    bad = L.EncodedWaveforms(enc.codec, enc.data.copy(), enc.offsets.copy(), enc.n_samples, 2, enc.shift)
    o = int(bad.offsets[5])
    bad.data[o + 2:o + 4] = (0, 200)         # a section longer than the stream
    bad.data[o + 4:o + 6] = (0, 16)
    with pytest.raises(L.LgdspError) as ei:
        L.decode_data(bad, handle)
    assert "malformed" in str(ei.value) and "event 5" in str(ei.value)
    wrong_len = L.EncodedWaveforms(enc.codec, enc.data, enc.offsets, 1000, 2, enc.shift)     # stored length 1024 != 1000
    with pytest.raises(L.LgdspError):
        L.decode_data(wrong_len, handle)
    enc2 = L.encode_waveforms(wf.astype(np.uint32), L.ULEB128_ZIGZAG_DIFF)
    cut = L.EncodedWaveforms(enc2.codec, enc2.data, enc2.offsets.copy(), 1024, 4, 0)
    cut.offsets[-1] -= 1                      # last varint may lose its terminator / a value goes missing
    with pytest.raises(L.LgdspError):
        L.decode_data(cut, handle)
    assert len(L.decode_data(L.encode_waveforms(wf[:0], L.RADWARE_SIGCOMPRESS), handle)) == 0   # empty in, empty out


def test_dsp_icpc_on_encoded_waveforms(L, O, handle):
    """lgdsp_icpc_run_encoded == lgdsp_icpc_run on the decoded samples, bit for bit (several chunks, ragged tail)"""
    P = L.resolve_icpc_params(L.example_config(), L.us(500.0))
    wf = L.synth.generate_host(9000, first_event=42)
    want = L.dsp_icpc_rows(wf, P, handle=handle)
    enc = L.encode_waveforms(wf, L.RADWARE_SIGCOMPRESS)
    got = np.empty_like(want)
    handle.icpc_run_encoded_host(P, enc.codec, enc.data.ctypes.data, enc.offsets.ctypes.data, enc.shift, 2, None, len(enc),
                                 got.ctypes.data)
    same = (got == want) | (np.isnan(got) & np.isnan(want))
    assert same.all()
    print("bytes per event over the host link:", enc.nbytes / len(enc) + 8, "instead of", wf.shape[1] * 2)


def test_dsp_icpc_compressed_with_decode_data_on_the_device(L, O, handle):
    """`dsp_icpc_compressed` fed with encoded waveform sets (the reference's decode_data calls, src/dsp_icpc.jl:313-314)
    gives the same table as fed with the decoded samples"""
    n_events, presum, window = 3000, 8, (2600, 1400)
    wf = L.synth.generate_host(n_events, first_event=31)
    pre, wdw = L.synth.compress(wf, presum, window)
    step = L.ns(16.0)
    base = {"presum_rate": np.full(n_events, presum, dtype=np.uint16)}
    plain = dict(base, waveform_presummed=L.RDWaveforms(pre, L.ns(0.0), step * float(presum)),
                 waveform_windowed=L.RDWaveforms(wdw, step * float(window[0]), step))
    enc_pre = L.encode_waveforms(pre.astype(np.uint32), L.ULEB128_ZIGZAG_DIFF)
    enc_wdw = L.encode_waveforms(wdw.astype(np.uint16), L.RADWARE_SIGCOMPRESS)
    coded = dict(base, waveform_presummed=L.RDWaveforms(enc_pre, L.ns(0.0), step * float(presum)),
                 waveform_windowed=L.RDWaveforms(enc_wdw, step * float(window[0]), step))
    cfg = L.example_config()
    want = L.dsp_icpc_compressed(plain, cfg, L.us(500.0), handle=handle)
    got = L.dsp_icpc_compressed(coded, cfg, L.us(500.0), handle=handle)
    assert list(want) == list(got)
    for k in want:
        a, b = np.asarray(want[k], dtype=np.float64), np.asarray(got[k], dtype=np.float64)
        assert ((a == b) | (np.isnan(a) & np.isnan(b))).all(), k
    print("encoded bytes per event:", (enc_pre.nbytes + enc_wdw.nbytes) / n_events, "raw:", pre.shape[1] * 4 + wdw.shape[1] * 2)
