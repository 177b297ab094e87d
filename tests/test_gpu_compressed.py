"""dsp_icpc_compressed (/root/reference/src/dsp_icpc.jl:293-499) on the GPU against the oracle's
restatement: presummed + windowed waveform per event, external baseline, 32-bit samples, auxiliary window statistics."""
import numpy as np
import pytest

from parity import TOL, EXACT

pytestmark = pytest.mark.gpu

# oracle / table column -> tolerance rule of the dsp_icpc column it corresponds to (parity.py)
_RULE = {"e_max_pre": "e_max", "e_min_pre": "e_min", "t50_pre": "t50", "tail_τ": "tail_tau"}
_STAT_TOL = {"mean": (1e-12, 1e-9), "sigma": (1e-7, 2e-5), "slope_sigma": (1e-6, 2e-5)}


def _check(table, ref, skip=()):
    bad = {}
    for name, got in table.items():
        key = "tail_tau" if name == "tail_τ" else name
        if key not in ref or name in skip:
            continue
        a, b = np.asarray(got, dtype=np.float64), ref[key]
        nan_a, nan_b = np.isnan(a), np.isnan(b)
        n_bad = int((nan_a != nan_b).sum())
        ok = ~(nan_a | nan_b)
        rule = _RULE.get(name, name)
        if rule == "tail_tau":
            with np.errstate(divide="ignore"):
                a = np.where(a == 0, 0.0, 1.0 / np.where(a == 0, 1.0, a))
                b = np.where(b == 0, 0.0, 1.0 / np.where(b == 0, 1.0, b))
        d = np.abs(a[ok] - b[ok])
        if rule in EXACT:
            n_bad += int((a[ok] != b[ok]).sum())
        else:
            if rule in TOL:
                rtol, atol = TOL[rule]
            else:                       # aux*/bl_slope_sigma statistics
                field = rule.split("_", 1)[1] if not rule.startswith("bl_") else "slope_sigma"
                rtol, atol = _STAT_TOL[field]
                if field != "mean":
                    # sigma = sqrt(E[y^2] - E[y]^2) with sequentially accumulated sums in the reference's formula: absolute
                    # error ~ sqrt(n_window * eps) * |mean| (flat, clipped windows of presummed samples ~ 5e5 give
                    # sigma ~ 0.04 instead of 0 there; the device sums are exact integers)
                    mean = ref[(rule[:-len(field)] + "mean") if not rule.startswith("bl_") else "blmean"]
                    atol = atol + 1e-6 * np.abs(mean[ok])
            n_bad += int((d > atol + rtol * np.abs(b[ok])).sum())
        if n_bad:
            bad[name] = (float(d.max()) if d.size else 0.0, n_bad)
    return bad


def _data(L, n_events, first_event, presum, window=(2600, 1400), **kw):
    wf = L.synth.generate_host(n_events, first_event=first_event, **kw)
    pre, wdw = L.synth.compress(wf, presum, window)
    step = L.ns(16.0)
    return {
        "waveform_presummed": L.RDWaveforms(pre, L.ns(0.0), step * float(presum)),
        "waveform_windowed": L.RDWaveforms(wdw, step * float(window[0]), step),
        "presum_rate": np.full(n_events, presum, dtype=np.uint16),
        "baseline": np.zeros(n_events, np.float32), "timestamp": np.arange(n_events, dtype=np.uint64),
        "eventnumber": np.arange(n_events, dtype=np.uint32), "daqenergy": np.zeros(n_events, np.uint16),
        "t_sat_lo": np.zeros(n_events, np.uint16), "t_sat_hi": np.zeros(n_events, np.uint16),
        "deadtime": np.zeros(n_events, np.uint16),
    }


def _oracle(L, O, data, cfg, presum):
    wp, ww = data["waveform_presummed"], data["waveform_windowed"]
    Pp, Pw, aux = L.resolve_compressed_params(cfg, L.us(500.0), None, presum_rate=presum, n_pre=wp.signal.shape[1],
                                              t_first_pre=wp.t_first, step_pre=wp.step, n_wdw=ww.signal.shape[1],
                                              t_first_wdw=ww.t_first, step_wdw=ww.step, builders=O.OracleBuilders())
    return O.dsp_icpc_compressed(Pp, Pw, wp.signal, ww.signal, presum, aux)


def test_reference_fixture(L, O, handle):
    """the reference's own test (test/test_dsp_icpc.jl:164-200): 3 identical noise-free events, both waveforms the full
    8192-sample trace, presum_rate 1, example config, pars_filter = PropDict(): table shape, timing order, finiteness"""
    wf = L.synth.generate_host(3, mode=1)
    data = {"waveform_presummed": L.RDWaveforms(wf), "waveform_windowed": L.RDWaveforms(wf),
            "presum_rate": np.ones(3, np.uint16), "baseline": np.zeros(3, np.float32), "timestamp": np.zeros(3, np.uint64),
            "eventnumber": np.arange(1, 4, dtype=np.uint32), "daqenergy": np.zeros(3, np.uint16),
            "t_sat_lo": np.zeros(3, np.uint16), "t_sat_hi": np.zeros(3, np.uint16), "deadtime": np.zeros(3, np.uint16)}
    cfg = L.example_config()
    res = L.dsp_icpc_compressed(data, cfg, L.us(500.0), {}, handle=handle, builders=O.OracleBuilders())
    assert all(len(v) == 3 for v in res.values())
    for col in ("blmean", "blsigma", "blslope", "bloffset", "tailmean", "tailsigma", "tailslope", "tailoffset", "t0", "t50",
                "t90", "drift_time", "e_10410", "e_313", "e_trap", "e_cusp", "e_zac", "qdrift", "lq", "a_sg", "n_sat_low",
                "n_sat_high", "inTrace_intersect", "inTrace_n", "e_10410_inv", "e_313_inv", "t0_inv"):
        assert col in res, col                                               # :177-186
    assert list(res.keys()) == list(L.COMPRESSED_COLUMNS.keys())
    assert (res["t0"] < res["t50"]).all() and (res["t50"] < res["t90"]).all() and (res["drift_time"] >= 0).all()  # :189-193
    for col in ("e_10410", "e_313", "e_trap"):
        assert np.isfinite(res[col]).all()                                   # :195-199
    ref = _oracle(L, O, data, cfg, 1)
    # noise-free: the in-trace threshold is 5 sigma of rounding noise, undefined (see test_gpu_icpc.py)
    bad = _check(res, ref, skip=("inTrace_intersect", "inTrace_n"))
    assert not bad, bad
    assert np.allclose(res["blmean"], 1000.0) and np.allclose(res["e_max"], 10000.0)
    assert np.array_equal(res["eventID_fadc"], data["eventnumber"])


@pytest.mark.parametrize("presum", [8, 4, 1])
def test_compressed_parity(L, O, handle, presum):
    """mixed population in the DAQ's compressed format: presummed trace (uint32 sums for presum > 1) + 1400-sample
    full-rate window, against the oracle's restatement of src/dsp_icpc.jl:293-499"""
    cfg = L.tiefree_config()
    n_events = 1024
    data = _data(L, n_events, 7000, presum)
    assert data["waveform_presummed"].signal.dtype == (np.uint32 if presum > 1 else np.uint16)
    res = L.dsp_icpc_compressed(data, cfg, L.us(500.0), None, handle=handle, builders=O.OracleBuilders())
    ref = _oracle(L, O, data, cfg, presum)
    bad = _check(res, ref, skip=("a_sg", "a_60", "a_100", "a_raw"))
    assert not bad, bad
    # currents: argmax ties inside the window are the only allowed difference (few events)
    for col in ("a_sg", "a_60", "a_100", "a_raw"):
        rtol, atol = TOL[col]
        n_bad = int((np.abs(res[col] - ref[col]) > atol + rtol * np.abs(ref[col])).sum())
        assert n_bad <= max(2, n_events // 100), (col, n_bad)
    # the population exercises saturation (at sat_high * presum_rate), empty events and pile-up
    assert (ref["n_sat_high"] > 0).any() and (ref["t0"] == 0).any() and (ref["inTrace_n"] > 1).any()
    assert (ref["t0"] > 0).sum() > n_events // 2


def test_ext_baseline_and_wide_samples(L, O, handle):
    """lgdsp_icpc_run_ext: (a) uint32 samples give the rows of the same values as uint16, (b) an external baseline equal
    to the waveform's own blmean reproduces the plain run bit for bit, (c) a different baseline moves e_max by exactly
    the difference"""
    from importlib import import_module
    cfgm = import_module("legenddsp.jl_b200.config")
    d = cfgm.example_config_dict()
    us = cfgm.us
    d["bl_window"] = {"min": us(0.0), "max": us(19.5)}
    d["tail_window"] = {"min": us(35.0), "max": us(55.0)}
    d["current_window"] = {"min": us(21.5), "max": us(31.0)}
    d["flt_length_cusp"] = d["flt_length_zac"] = us(8.0)
    d["flt_defaults"]["trap"] = d["flt_defaults"]["cusp"] = d["flt_defaults"]["zac"] = {"rt": us(1.0), "ft": us(0.5)}
    cfg = cfgm.DSPConfig.from_dict(d)
    n = 4096
    P = L.resolve_icpc_params(cfg, L.us(500.0), n_samples=n, builders=O.OracleBuilders())
    wf = L.synth.generate_host(256, first_event=11, n_samples=n)
    plain = L.dsp_icpc_rows(wf, P, handle=handle)

    def run_ext(sig, baseline):
        rows = np.zeros((sig.shape[0], L.NCOL))
        handle.icpc_run_ext_host(P, sig.ctypes.data, sig.dtype.itemsize, baseline.ctypes.data if baseline is not None else None,
                                 sig.shape[0], sig.shape[1], rows.ctypes.data)
        return rows

    wide = run_ext(wf.astype(np.uint32), None)
    assert np.array_equal(wide, plain, equal_nan=True)
    bl = np.ascontiguousarray(plain[:, L.COL["blmean"]])
    same = run_ext(wf, bl)
    assert np.array_equal(same, plain, equal_nan=True)
    moved = run_ext(wf, bl + 2.0)
    assert np.array_equal(moved[:, L.COL["e_max"]], (wf.max(axis=1) - (bl + 2.0)))
    assert np.array_equal(moved[:, L.COL["blmean"]], plain[:, L.COL["blmean"]])       # own statistics still reported
    # 32-bit samples need n_samples <= LGDSP_MAX_SAMPLES / 2
    P8 = L.resolve_icpc_params(L.example_config(), L.us(500.0), builders=O.OracleBuilders())
    big = np.zeros((2, 8192), np.uint32)
    with pytest.raises(L.LgdspError):
        handle.icpc_run_ext_host(P8, big.ctypes.data, 4, None, 2, 8192, np.zeros((2, L.NCOL)).ctypes.data)


def test_window_stats(L, O, handle):
    """lgdsp_window_stats_run against the oracle's signalstats (+ residual sigma) on raw and shifted traces"""
    wf = L.synth.generate_host(64, first_event=3)
    pre, _ = L.synth.compress(wf, 8)
    wins = [(0, 155), (156, 304), (547, 702), (10, 10 + 63), (0, 1023)]
    shift = np.linspace(70000.0, 110000.0, pre.shape[0]) + 0.37
    for sig, sh in ((pre, None), (pre, shift), (wf[:, :4096].copy(), None)):
        out = np.zeros((sig.shape[0], len(wins), 5))
        handle.window_stats_host(sig.ctypes.data, sig.dtype.itemsize, sig.shape[0], sig.shape[1], sig.shape[1], 5.0, 128.0,
                                 sh.ctypes.data if sh is not None else None, wins, out.ctypes.data)
        for e in range(0, sig.shape[0], 7):
            y = sig[e].astype(np.float64) - (sh[e] if sh is not None else 0.0)
            for w, (a, b) in enumerate(wins):
                ref = O.signalstats5(y, 5.0, 128.0, a, b)
                got = out[e, w]
                assert abs(got[0] - ref[0]) <= 1e-12 * abs(ref[0]) + 1e-9, (e, w, got, ref)
                assert abs(got[1] - ref[1]) <= 1e-7 * abs(ref[1]) + 2e-4, (e, w, got, ref)
                assert abs(got[2] - ref[2]) <= 1e-7 * abs(ref[2]) + 1e-12, (e, w, got, ref)
                assert abs(got[3] - ref[3]) <= 1e-9 * abs(ref[3]) + 1e-6, (e, w, got, ref)
                assert abs(got[4] - ref[4]) <= 1e-6 * abs(ref[4]) + 2e-4, (e, w, got, ref)
    with pytest.raises(L.LgdspError):
        handle.window_stats_host(pre.ctypes.data, 4, pre.shape[0], pre.shape[1], pre.shape[1], 0.0, 128.0, None, [(5, 2000)],
                                 np.zeros((pre.shape[0], 1, 5)).ctypes.data)


def test_compressed_errors_and_empty(L, O, handle):
    cfg = L.tiefree_config()
    data = _data(L, 4, 0, 8)
    bad = dict(data)
    bad["presum_rate"] = np.array([8, 8, 4, 8], np.uint16)
    with pytest.raises(ValueError):
        L.dsp_icpc_compressed(bad, cfg, L.us(500.0), None, handle=handle)
    # a window that does not contain current_window: the reference's index assertion
    short = _data(L, 4, 0, 8, window=(2000, 1400))
    with pytest.raises(AssertionError):
        L.dsp_icpc_compressed(short, cfg, L.us(500.0), None, handle=handle)
    # an empty table has no presum rate: `only(unique(presum_rate))` (:324) throws in the reference as well
    with pytest.raises(ValueError):
        L.dsp_icpc_compressed(_data(L, 0, 0, 8), cfg, L.us(500.0), None, handle=handle)
    # empty batches through the C ABI are fine
    Pp, Pw, aux = L.resolve_compressed_params(cfg, L.us(500.0), None, presum_rate=8, n_pre=1024, step_pre=L.ns(128.0),
                                              n_wdw=1400, t_first_wdw=L.ns(41600.0), step_wdw=L.ns(16.0))
    handle.icpc_compressed_run_host(Pp, Pw, 0, 4, 1024, 0, 2, 1400, 8.0, aux, 0, 0, 0, 0)


def test_wide_samples_beyond_32_bit_sums_give_nan_rows(L, O, handle):
    """32-bit samples whose sum leaves 32 bits (mean >= 2^20 over 4096 samples) cannot be processed with exact uint32 prefix
    sums: the row is NaN, the neighbouring events are untouched (both execution paths)"""
    cfg = L.example_config()
    cfg.bl_window = (L.us(0.0), L.us(20.0))
    cfg.tail_window = (L.us(52.0), L.us(62.0))
    cfg.current_window = (L.us(30.0), L.us(55.0))
    P = L.resolve_icpc_params(cfg, L.us(500.0), n_samples=4096, groups=0x07)
    wf = L.synth.generate_host(6, first_event=5)[:, :4096].astype(np.uint32)
    big = wf.copy()
    big[2] = big[2] * 64 + (1 << 20)          # sums to ~ 2^33
    rows, ref = [], None
    for path in ("split", "fused"):
        handle.set_icpc_path(path)
        out = np.zeros((6, L.NCOL))
        ok = np.zeros((6, L.NCOL))
        handle.icpc_run_ext_host(P, big.ctypes.data, 4, None, 6, 4096, out.ctypes.data)
        handle.icpc_run_ext_host(P, wf.ctypes.data, 4, None, 6, 4096, ok.ctypes.data)
        assert np.isnan(out[2]).all(), path
        keep = [0, 1, 3, 4, 5]
        assert np.array_equal(np.nan_to_num(out[keep]), np.nan_to_num(ok[keep])), path
    handle.set_icpc_path("split")
