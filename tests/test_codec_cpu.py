"""decode_data codecs on the CPU: the oracle restatements (oracle/lgdsp_codec_oracle.c) against the hand-worked vectors of
tests/golden/make_codec_kat.py, encode -> decode round trips, and the product's independent host encoders byte for byte
against the oracle's (no GPU needed: the product encoders are host code)."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "codec_kat.json")))


def _populations(rng, n):
    yield "noise", np.clip(np.rint(12000 + rng.normal(0, 3.0, n)), 0, 65535).astype(np.uint16)
    yield "constant", np.full(n, 40000, dtype=np.uint16)
    yield "ramp", (np.arange(n) * 7 % 65536).astype(np.uint16)
    yield "full range", rng.integers(0, 65536, n).astype(np.uint16)
    yield "alternating extremes", np.where(np.arange(n) % 2 == 0, 0, 65535).astype(np.uint16)
    x = np.clip(np.rint(9000 + rng.normal(0, 2.0, n)), 0, 65535)
    x[n // 3:] += 30000 * np.exp(-np.arange(n - n // 3) / 3000.0)
    yield "pulse", np.clip(np.rint(x), 0, 65535).astype(np.uint16)
    yield "saturated", np.minimum(x * 3, 65520).astype(np.uint16)


@pytest.mark.parametrize("case", KAT["radware"], ids=lambda c: c["name"])
def test_radware_known_answers(O, L, case):
    sig = np.array(case["signal"], dtype=np.uint16)
    want = np.array(case["bytes"], dtype=np.uint8)
    got = O.radware_encode(sig, shift=case["shift"])
    assert got.tolist() == want.tolist()
    assert O.radware_decode(want, shift=case["shift"]).tolist() == sig.tolist()
    enc = L.encode_waveforms(sig[None, :], L.RADWARE_SIGCOMPRESS, shift=case["shift"])
    assert enc.data.tolist() == want.tolist() and enc.offsets.tolist() == [0, want.size]


@pytest.mark.parametrize("case", KAT["uleb128zzd"], ids=lambda c: c["name"])
def test_uleb_known_answers(O, L, case):
    sig = np.array(case["signal"], dtype=case["dtype"])
    want = np.array(case["bytes"], dtype=np.uint8)
    assert O.uleb128zzd_encode(sig).tolist() == want.tolist()
    assert O.uleb128zzd_decode(want, dtype=sig.dtype).tolist() == sig.tolist()
    enc = L.encode_waveforms(sig[None, :], L.ULEB128_ZIGZAG_DIFF)
    assert enc.data.tolist() == want.tolist()


@pytest.mark.parametrize("n", [1, 2, 47, 48, 49, 127, 128, 129, 1000, 8192])
def test_radware_round_trip_and_encoder_agreement(O, L, n):
    rng = np.random.default_rng(n)
    for name, x in _populations(rng, n):
        b = O.radware_encode(x)
        assert b.size % 4 == 0 and b.size <= L.load_library().lgdsp_codec_max_encoded_bytes(1, n, 2), name
        assert np.array_equal(O.radware_decode(b), x), name
        enc = L.encode_waveforms(x[None, :], L.RADWARE_SIGCOMPRESS)
        assert np.array_equal(enc.data, b), f"{name}: product encoder differs from the oracle's"
        # little-endian word order of the original C library: same words, swapped bytes
        le = O.radware_encode(x, big_endian=False)
        assert np.array_equal(le.reshape(-1, 2)[:, ::-1].ravel(), b), name


def test_radware_rejects_out_of_range_shift(O, L):
    x = np.array([0, 65535], dtype=np.uint16)
    with pytest.raises(ValueError):
        O.radware_encode(x, shift=0)          # 65535 does not fit int16 without the shift
    with pytest.raises(ValueError):
        L.encode_waveforms(x[None, :], L.RADWARE_SIGCOMPRESS, shift=0)


def test_radware_decoder_rejects_malformed_streams(O):
    good = O.radware_encode(np.arange(100, dtype=np.uint16) + 30000)
    for bad in (good[:6], good[:-8], np.concatenate([good[:4], np.array([0, 77], dtype=np.uint8), good[6:]])):
        with pytest.raises(ValueError):
            O.radware_decode(bad)


@pytest.mark.parametrize("dtype", ["uint16", "uint32"])
def test_uleb_round_trip_and_encoder_agreement(O, L, dtype):
    rng = np.random.default_rng(3)
    hi = 65536 if dtype == "uint16" else 2 ** 32
    for n in (1, 5, 1024, 4096):
        for x in (rng.integers(0, hi, n).astype(dtype), (np.arange(n) * 3).astype(dtype), np.zeros(n, dtype=dtype),
                  np.clip(np.rint(100000 * (dtype == "uint32") + 9000 + rng.normal(0, 20, n)), 0, hi - 1).astype(dtype)):
            b = O.uleb128zzd_encode(x)
            assert np.array_equal(O.uleb128zzd_decode(b, dtype=np.dtype(dtype)), x)
            enc = L.encode_waveforms(x[None, :], L.ULEB128_ZIGZAG_DIFF)
            assert np.array_equal(enc.data, b)


def test_batch_offsets(O, L):
    rng = np.random.default_rng(9)
    wf = np.stack([x for _, x in _populations(rng, 777)])
    enc = L.encode_waveforms(wf, L.RADWARE_SIGCOMPRESS)
    assert enc.offsets[0] == 0 and len(enc) == wf.shape[0]
    for e in range(wf.shape[0]):
        assert np.array_equal(O.radware_decode(enc.data[enc.offsets[e]:enc.offsets[e + 1]]), wf[e])
