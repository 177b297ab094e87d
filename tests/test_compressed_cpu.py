"""CPU checks of the dsp_icpc_compressed restatement (oracle) and of the host logic around it (no GPU)."""
import numpy as np
import pytest


def _fixture_params(L, O, presum=1, n_pre=8192, n_wdw=8192, t_first_wdw=0.0):
    return L.resolve_compressed_params(L.example_config(), L.us(500.0), None, presum_rate=presum, n_pre=n_pre,
                                       step_pre=L.ns(16.0 * presum), n_wdw=n_wdw, t_first_wdw=L.ns(t_first_wdw),
                                       step_wdw=L.ns(16.0), builders=O.OracleBuilders())


def test_oracle_compressed_on_reference_fixture(L, O):
    """closed-form expectations on the reference's noise-free generator and the properties the
    reference's own test asserts (test/test_dsp_icpc.jl:164-200)"""
    Pp, Pw, aux = _fixture_params(L, O)
    wf = L.synth.generate_host(3, mode=1)
    r = O.dsp_icpc_compressed(Pp, Pw, wf, wf, 1, aux)
    assert set(O.compressed_columns()) == {("tail_tau" if k == "tail_τ" else k) for k, (src, _) in L.COMPRESSED_COLUMNS.items()
                                           if src != "pass"}
    assert np.allclose(r["blmean"], 1000.0, atol=1e-9) and np.allclose(r["bloffset"], 1000.0, atol=1e-6)
    assert np.allclose(r["blsigma"], 0.0, atol=1e-4) and np.allclose(r["bl_slope_sigma"], 0.0, atol=1e-4)
    for k in ("auxbl1_mean", "auxbl2_mean"):
        assert np.allclose(r[k], 1000.0, atol=1e-9)
    assert np.allclose(r["e_max"], 10000.0, atol=0.5) and np.allclose(r["e_max_pre"], r["e_max"])
    assert np.allclose(r["tail_tau"], 500000.0, rtol=1e-4)
    assert (r["t0"] < r["t50"]).all() and (r["t50"] < r["t90"]).all() and (r["drift_time"] >= 0).all()
    assert np.isfinite(r["e_10410"]).all() and np.isfinite(r["e_313"]).all() and np.isfinite(r["e_trap"]).all()
    # the tail of the baseline-subtracted trace inside auxpz1 (70..90 us) is the decaying exponential: mean within its range
    assert ((r["auxpz1_mean"] > r["auxpz2_mean"]) & (r["auxpz2_mean"] > 8000.0)).all()
    assert (r["qc_label"] == -1).all()


def test_oracle_compressed_equals_dsp_icpc_when_uncompressed(L, O):
    """presum_rate 1 and the full trace as both waveforms: every column that dsp_icpc also has must be identical to the
    oracle's dsp_icpc (two restatements of the same steps), except those of the in-trace filter whose window length is
    sg_wl * presum_rate / 2 here (src/dsp_icpc.jl:439 vs :181)"""
    Pp, Pw, aux = _fixture_params(L, O)
    P = L.resolve_icpc_params(L.example_config(), L.us(500.0), builders=O.OracleBuilders())
    wf = L.synth.generate_host(96, first_event=50)
    r = O.dsp_icpc_compressed(Pp, Pw, wf, wf, 1, aux)
    rows, _ = O.dsp_icpc(P, wf)
    for name in L.COLUMNS:
        if name in ("inTrace_intersect", "inTrace_n", "t50_current"):
            continue
        assert np.array_equal(r[name], rows[:, L.COL[name]], equal_nan=True), name
    assert np.array_equal(r["t50_pre"], r["t50"]) and np.array_equal(r["e_max_pre"], r["e_max"])


def test_signalstats5_against_polyfit(O):
    rng = np.random.default_rng(5)
    t0, dt = 32.0, 16.0
    y = 1000.0 + 0.01 * (t0 + dt * np.arange(500)) + rng.normal(0, 3.0, 500)
    a, b = 40, 339
    s = O.signalstats5(y, t0, dt, a, b)
    x = t0 + dt * np.arange(a, b + 1)
    slope, offset = np.polyfit(x, y[a:b + 1], 1)
    res = y[a:b + 1] - (offset + slope * x)
    assert np.allclose(s[:4], [y[a:b + 1].mean(), y[a:b + 1].std(), slope, offset], rtol=1e-8)
    assert np.isclose(s[4], np.sqrt(np.mean(res ** 2)), rtol=1e-9)
    assert s[4] <= s[1]


def test_resolve_compressed_params(L, O):
    Pp, Pw, aux = _fixture_params(L, O, presum=8, n_pre=1024, n_wdw=1400, t_first_wdw=41600.0)
    kw = L.example_config().kwargs_pars
    assert Pp.sat_high == (2 ** 16 - 16) * 8 and Pw.sat_high == 2 ** 16 - 16                    # src/dsp_icpc.jl:334
    assert Pp.dt_ns == 128.0 and Pw.dt_ns == 16.0 and Pw.t_first_ns == 41600.0
    # InvCRFilter(tau) per time axis (:370-372)
    assert np.isclose(Pp.pz_km1, 128.0 / 500000.0) and np.isclose(Pw.pz_km1, 16.0 / 500000.0)
    # the in-trace filter: SavitzkyGolayFilter(sg_wl * presum / 2) = 400 ns on a 128 ns axis -> 3 taps (:439)
    assert Pp.sg[0].n_taps == 3
    assert (Pp.trap_10410.navg, Pp.trap_10410.ngap) == (78, 31) and Pp.cusp.n_taps == 297
    assert Pw.t0_trap.navg2 == 125 and Pw.t0_min_n == 94
    assert aux == [(0, 156), (156, 305), (547, 703), (703, 859)]
    assert Pp.groups & L._abi.GROUP_CUSPZAC and not (Pw.groups & L._abi.GROUP_CUSPZAC)
    # a windowed trace that does not cover current_window -> the reference's index assertion
    with pytest.raises(AssertionError):
        _fixture_params(L, O, presum=8, n_pre=1024, n_wdw=1400, t_first_wdw=32000.0)
    # presum rate so high that the signal estimator window has fewer samples than polynomial coefficients
    with pytest.raises(ValueError):
        _fixture_params(L, O, presum=16, n_pre=512, n_wdw=1400, t_first_wdw=41600.0)


def test_compress_helper(L):
    wf = L.synth.generate_host(5, first_event=1)
    pre, wdw = L.synth.compress(wf, 8, (2600, 1400))
    assert pre.shape == (5, 1024) and pre.dtype == np.uint32 and wdw.shape == (5, 1400) and wdw.dtype == np.uint16
    assert np.array_equal(pre[:, 3], wf[:, 24:32].astype(np.uint32).sum(axis=1))
    assert np.array_equal(wdw, wf[:, 2600:4000])
    pre1, _ = L.synth.compress(wf, 1)
    assert pre1.dtype == np.uint16 and np.array_equal(pre1, wf)
