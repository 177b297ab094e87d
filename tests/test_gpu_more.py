"""More GPU tests through the C ABI: sweeps, generator, ragged/edge inputs, error behaviour, size-independent
properties at large batch sizes."""
import ctypes as C

import numpy as np
import pytest

from parity import assert_parity, compare_rows
from test_gpu_icpc import assert_parity_with_ties

pytestmark = pytest.mark.gpu


def test_trap_sweeps_match_the_oracle(L, O, handle):
    """dsp_trap_rt_optimization / dsp_trap_ft_optimization (src/dsp_filter_optimization.jl:102-133, 241-274)"""
    cfg = L.example_config()
    tau = L.us(500.0)
    wf = L.synth.generate_host(300, first_event=50)
    W = L.RDWaveforms(wf)
    so = L.resolve_sweep_params(cfg, tau, builders=O.OracleBuilders())
    # rt sweep, fixed pick-off (Float64 [n_rt, n_events])
    got = L.dsp_trap_rt_optimization(W, cfg, tau, ft=L.us(2.0), handle=handle)
    var = L.trap_variants(L.grid_values(cfg.e_grid_rt_trap), [L.us(2.0)], L.ns(16.0), mode="rt", pickoff=cfg.enc_pickoff_trap)
    ref = O.trap_sweep(so, wf, var).T
    assert got.shape == (31, 300) and got.dtype == np.float64
    assert np.allclose(got, ref, rtol=1e-6, atol=1e-6, equal_nan=True)
    # ft sweep, pick-off t50 + rt + ft/2 (Float32 [n_ft, n_events])
    got = L.dsp_trap_ft_optimization(W, cfg, tau, L.us(5.0), handle=handle)
    var = L.trap_variants([L.us(5.0)], L.grid_values(cfg.e_grid_ft_trap), L.ns(16.0), mode="ft")
    ref = O.trap_sweep(so, wf, var).T
    assert got.shape == (16, 300) and got.dtype == np.float32
    assert np.allclose(got, ref, rtol=1e-6, atol=1e-6, equal_nan=True)
    # the batched 20x10 grid equals the loop over the ft sweep
    rts = [L.us(1.0 + 0.75 * i) for i in range(20)]
    fts = [L.us(1.0 + 0.3 * i) for i in range(10)]
    grid = L.dsp_trap_rtft_grid(W, cfg, tau, rts, fts, handle=handle)
    assert grid.shape == (20, 10, 300)
    var = L.trap_variants(rts, fts, L.ns(16.0), mode="ft")
    ref = O.trap_sweep(so, wf, var)
    assert np.allclose(grid.reshape(200, 300).T, ref, rtol=1e-6, atol=1e-6, equal_nan=True)
    # e_trap of the full chain is the (rt=5, ft=2.5) point of the sweep when both use the same t50 convention:
    # the chain thresholds at 0.5*max of the PRE-PZ waveform (src/dsp_icpc.jl:133), the sweep at 0.5*max of the PZ one
    # (src/dsp_filter_optimization.jl:260) -- so they differ slightly; only check the scale
    P = L.resolve_icpc_params(cfg, tau, builders=O.OracleBuilders())
    rows = L.dsp_icpc_rows(wf, P, handle=handle)
    one = L.dsp_trap_rtft_grid(W, cfg, tau, [L.us(5.0)], [L.us(2.5)], handle=handle)[0, 0]
    big = rows[:, L.COL["e_max"]] > 500
    assert np.allclose(one[big], rows[big, L.COL["e_trap"]], rtol=2e-3)


def test_generator_host_equals_device(L, handle):
    import torch
    n = 257
    d = torch.empty((n, 8192), dtype=torch.int16, device="cuda")
    L.synth.generate_device(handle, d.data_ptr(), n, first_event=12345)
    handle.synchronize()
    dev = d.cpu().numpy().view(np.uint16)
    host = L.synth.generate_host(n, first_event=12345)
    # identical counter-based stream; libm (host) vs CUDA math differ in the last ulp of exp/log/sincos, which can
    # flip the rounding of a sample in rare cases
    mism = int((dev != host).sum())
    assert mism <= n * 8192 * 1e-6, mism
    assert int(np.abs(dev.astype(np.int32) - host.astype(np.int32)).max()) <= 1
    # the population has every class: clipped, empty, pile-up
    assert (host.max(axis=1) == 65520).any() and (host.max(axis=1) < 16000).any()


@pytest.mark.parametrize("n_samples", [1024, 4096, 5592])
def test_ragged_sample_counts(L, O, handle, n_samples):
    """shorter traces (windowed waveforms): a config scaled to the trace length"""
    scale = n_samples / 8192.0
    d = L.config.example_config_dict() if hasattr(L, "config") else None
    from importlib import import_module
    cfgm = import_module("legenddsp.jl_b200.config")
    d = cfgm.example_config_dict()
    us = cfgm.us
    d["bl_window"] = {"min": us(0.0), "max": us(39.0 * scale)}
    d["tail_window"] = {"min": us(70.0 * scale), "max": us(110.0 * scale)}
    d["current_window"] = {"min": us(43.0 * scale), "max": us(62.0 * scale)}
    d["flt_length_cusp"] = d["flt_length_zac"] = us(16.0 * scale)
    d["flt_defaults"]["trap"] = d["flt_defaults"]["cusp"] = d["flt_defaults"]["zac"] = {"rt": us(2.0 * scale), "ft": us(1.0 * scale)}
    d["qdrift_int_length"] = d["lq_int_length"] = (us(1.0 * scale), us(2.0 * scale))
    cfg = cfgm.DSPConfig.from_dict(d)
    # the fixed 10/4 us trapezoid must still fit: only for traces >= ~25 us
    if n_samples * 16e-3 < 30:
        with pytest.raises((ValueError, AssertionError)):
            L.resolve_icpc_params(cfg, L.us(500.0), n_samples=n_samples, builders=O.OracleBuilders())
        return
    P = L.resolve_icpc_params(cfg, L.us(500.0), n_samples=n_samples, builders=O.OracleBuilders())
    wf = L.synth.generate_host(200, first_event=3, n_samples=n_samples)
    got = L.dsp_icpc_rows(wf, P, handle=handle)
    ref, _ = O.dsp_icpc(P, wf)
    assert_parity_with_ties(L, O, P, wf, got, ref)


def test_separate_cusp_and_zac_parameters_and_direct_mode(L, O, handle):
    """pars_filter with different rt/ft for cusp and zac (two structured passes) == direct FIR == oracle"""
    pf = {"cusp": {"rt": L.us(3.0), "ft": L.us(1.2)}, "zac": {"rt": L.us(7.5), "ft": L.us(3.1)},
          "trap": {"rt": L.us(8.0), "ft": L.us(3.0)}, "sg": {"wl": L.ns(180.0)}}
    cfg = L.example_config()
    wf = L.synth.generate_host(256, first_event=900)
    Ps = L.resolve_icpc_params(cfg, L.us(500.0), pf, builders=O.OracleBuilders())
    Pd = L.resolve_icpc_params(cfg, L.us(500.0), pf, builders=O.OracleBuilders(), cuspzac_direct=True)
    ref, _ = O.dsp_icpc(Ps, wf)
    gs = L.dsp_icpc_rows(wf, Ps, handle=handle)
    gd = L.dsp_icpc_rows(wf, Pd, handle=handle)
    assert_parity_with_ties(L, O, Ps, wf, gs, ref)
    assert_parity_with_ties(L, O, Pd, wf, gd, ref)


def test_lib_builders_equal_oracle_builders_end_to_end(L, O, handle):
    wf = L.synth.generate_host(64, first_event=77)
    Pl = L.resolve_icpc_params(L.example_config(), L.us(500.0))
    Po = L.resolve_icpc_params(L.example_config(), L.us(500.0), builders=O.OracleBuilders())
    a, b = L.dsp_icpc_rows(wf, Pl, handle=handle), L.dsp_icpc_rows(wf, Po, handle=handle)
    assert_parity(a, b, L.COLUMNS)


def test_public_api_table(L, O, handle):
    """dsp_icpc(data, config, tau, pars_filter): column names/order of src/dsp_icpc.jl:210-229, pass-through columns"""
    n = 10
    wf = L.synth.generate_host(n, first_event=1)
    data = {"waveform": L.RDWaveforms(wf, L.ns(0.0), L.ns(16.0)), "baseline": np.arange(n, dtype=np.float32),
            "timestamp": np.arange(n, dtype=np.uint64), "eventnumber": np.arange(n, dtype=np.uint32),
            "daqenergy": np.arange(n, dtype=np.uint16)}
    tab = L.dsp_icpc(data, L.example_config(), L.us(500.0), {}, handle=handle)
    assert tuple(tab.keys()) == L.TABLE_COLUMNS and len(tab) == 53
    assert tab["qc_label"].dtype == np.int64 and (tab["qc_label"] == -1).all()
    assert np.array_equal(tab["blfc"], data["baseline"]) and np.array_equal(tab["eventID_fadc"], data["eventnumber"])
    P = L.resolve_icpc_params(L.example_config(), L.us(500.0), builders=O.OracleBuilders())
    ref, _ = O.dsp_icpc(P, wf)
    assert np.allclose(tab["e_trap"], ref[:, L.COL["e_trap"]], rtol=1e-9)
    with pytest.raises(NotImplementedError):
        L.dsp_icpc(data, L.example_config(), L.us(500.0), {}, f_evaluate_qc=lambda x: x, handle=handle)


def test_empty_and_strided_inputs(L, O, handle):
    P = L.resolve_icpc_params(L.example_config(), L.us(500.0), builders=O.OracleBuilders())
    rows = L.dsp_icpc_rows(np.zeros((0, 8192), dtype=np.uint16), P, handle=handle)
    assert rows.shape == (0, 49)
    # leading dimension > n_samples (a view into a wider buffer)
    wide = np.zeros((33, 8200), dtype=np.uint16)
    wf = L.synth.generate_host(33, first_event=5)
    wide[:, :8192] = wf
    a = L.dsp_icpc_rows(wide[:, :8192], P, handle=handle)
    b = L.dsp_icpc_rows(wf, P, handle=handle)
    assert np.array_equal(np.nan_to_num(a, nan=-7), np.nan_to_num(b, nan=-7))


def test_error_behaviour(L, O, handle):
    P = L.resolve_icpc_params(L.example_config(), L.us(500.0), builders=O.OracleBuilders())
    wf = L.synth.generate_host(4)
    bad = L._abi.IcpcParams.from_buffer_copy(P)
    bad.tail_until = 9000                       # window outside the trace -> AssertionError in the reference
    with pytest.raises(L.LgdspError) as ei:
        L.dsp_icpc_rows(wf, bad, handle=handle)
    assert ei.value.code == L._abi.LGDSP_ERR_INVALID_ARG and "tail_window" in str(ei.value)
    bad = L._abi.IcpcParams.from_buffer_copy(P)
    bad.version = 1
    with pytest.raises(L.LgdspError):
        L.dsp_icpc_rows(wf, bad, handle=handle)
    bad = L._abi.IcpcParams.from_buffer_copy(P)
    bad.n_samples = 8190
    with pytest.raises(L.LgdspError) as ei:
        L.dsp_icpc_rows(wf, bad, handle=handle)
    assert ei.value.code == L._abi.LGDSP_ERR_UNSUPPORTED
    bad = L._abi.IcpcParams.from_buffer_copy(P)
    bad.zac.coeffs[100] *= 1.01                 # coefficient array inconsistent with (sigma, flat, tau, L, beta)
    with pytest.raises(L.LgdspError) as ei:
        L.dsp_icpc_rows(wf, bad, handle=handle)
    assert "coeffs" in str(ei.value)
    with pytest.raises(TypeError):
        L.dsp_icpc_rows(wf.astype(np.float64), P, handle=handle)
    # the handle still works after errors
    assert L.dsp_icpc_rows(wf, P, handle=handle).shape == (4, 49)


def test_full_size_properties(L, O, handle):
    """size-independent properties on a batch too large for the oracle (BASELINE-size batches are checked with the
    same properties by bench.py's pool): determinism, batch-split invariance, permutation equivariance, exact
    relations between columns, and oracle parity on a random subsample"""
    import torch
    n = 65536
    d = torch.empty((n, 8192), dtype=torch.int16, device="cuda")
    L.synth.generate_device(handle, d.data_ptr(), n, first_event=2_000_000)
    P = L.resolve_icpc_params(L.tiefree_config(), L.us(500.0), builders=O.OracleBuilders())
    out = torch.empty((n, 49), dtype=torch.float64, device="cuda")
    handle.icpc_run_device(P, d.data_ptr(), n, 8192, out.data_ptr())
    handle.synchronize()
    a = out.cpu().numpy()
    out2 = torch.empty((n, 49), dtype=torch.float64, device="cuda")
    handle.icpc_run_device(None, d.data_ptr(), n, 8192, out2.data_ptr())
    handle.synchronize()
    b = out2.cpu().numpy()
    assert np.array_equal(np.nan_to_num(a, nan=-7), np.nan_to_num(b, nan=-7))          # deterministic
    # split invariance: second half alone == rows of the full run
    h2 = torch.empty((n // 2, 49), dtype=torch.float64, device="cuda")
    handle.icpc_run_device(None, d[n // 2:].data_ptr(), n // 2, 8192, h2.data_ptr())
    handle.synchronize()
    assert np.array_equal(np.nan_to_num(h2.cpu().numpy(), nan=-7), np.nan_to_num(a[n // 2:], nan=-7))
    c = L.COL
    # exact relations: drift_time = (t90 - t0)*1000, qc_label, e_max - e_min = raw max - raw min, counts within bounds
    assert np.array_equal(a[:, c["drift_time"]], (a[:, c["t90"]] - a[:, c["t0"]]) * 1000.0)
    assert (a[:, c["qc_label"]] == -1).all()
    raw = d.cpu().numpy().view(np.uint16)
    assert np.array_equal(a[:, c["e_max"]] - a[:, c["e_min"]], (raw.max(axis=1).astype(float) - raw.min(axis=1)))
    assert np.array_equal(a[:, c["n_sat_high"]], (raw == 65520).sum(axis=1))
    assert np.array_equal(a[:, c["n_sat_low"]], (raw == 0).sum(axis=1))
    assert (a[:, c["n_sat_high_cons"]] <= a[:, c["n_sat_high"]]).all()
    assert (a[:, c["e_trap_max"]] >= a[:, c["e_trap"]] - 1e-3 * np.abs(a[:, c["e_trap"]]) - 5).all()
    assert np.array_equal(a[:, c["blmean"]], raw[:, :2439].astype(np.float64).sum(axis=1) * (1.0 / 2439))
    # oracle parity on a random subsample
    rng = np.random.default_rng(0)
    idx = np.sort(rng.choice(n, 1500, replace=False))
    ref, _ = O.dsp_icpc(P, raw[idx])
    assert_parity_with_ties(L, O, P, raw[idx], a[idx], ref)


@pytest.mark.parametrize("n_samples", [2048, 1400, 512])
def test_short_traces(L, O, handle, n_samples):
    """short traces (the lengths of the compressed format's windowed waveforms): the whole chain, or for lengths the fixed
    trapezoids do not fit the windowed-waveform pass of dsp_icpc_compressed, against the oracle on full-rate slices"""
    from importlib import import_module
    cfgm = import_module("legenddsp.jl_b200.config")
    us = cfgm.us
    scale = n_samples / 2048.0
    d = cfgm.example_config_dict()
    d["bl_window"] = {"min": us(0.0), "max": us(4.0 * scale)}
    d["tail_window"] = {"min": us(16.0 * scale), "max": us(30.0 * scale)}
    d["current_window"] = {"min": us(4.0 * scale), "max": us(12.0 * scale)}
    d["flt_length_cusp"] = d["flt_length_zac"] = us(6.0 * scale)
    d["flt_defaults"]["trap"] = d["flt_defaults"]["cusp"] = d["flt_defaults"]["zac"] = {"rt": us(1.0 * scale), "ft": us(0.5 * scale)}
    d["qdrift_int_length"] = d["lq_int_length"] = (us(1.0 * scale), us(2.0 * scale))
    cfg = cfgm.DSPConfig.from_dict(d)
    full = L.synth.generate_host(400, first_event=77)
    start = 3000 - int(400 * scale)
    wf = np.ascontiguousarray(full[:, start:start + n_samples])
    if n_samples < 1600:
        # the fixed 10/4 us trapezoid (1500 samples) does not fit: the reference's filters would fail the same way
        with pytest.raises((ValueError, AssertionError)):
            L.resolve_icpc_params(cfg, L.us(500.0), n_samples=n_samples, builders=O.OracleBuilders())
        # ... so this length runs the windowed-waveform pass of dsp_icpc_compressed (timing, Q-drift, currents)
        P = L.resolve_icpc_params(cfg, L.us(500.0), n_samples=n_samples, builders=O.OracleBuilders(), role="wdw")
    else:
        P = L.resolve_icpc_params(cfg, L.us(500.0), n_samples=n_samples, builders=O.OracleBuilders())
    got = L.dsp_icpc_rows(wf, P, handle=handle)
    ref, _ = O.dsp_icpc(P, wf)
    if n_samples < 1600:
        # columns of the groups this pass switches off are 0 on the device; the oracle evaluates the placeholders
        for name in ("e_10410", "e_535", "e_313", "e_10410_inv", "e_313_inv", "e_trap", "e_trap_max", "t_trap_max", "e_cusp",
                     "e_zac", "e_cusp_max", "e_zac_max", "t_cusp_max", "t_zac_max", "t50_current", "inTrace_intersect", "inTrace_n"):
            assert (got[:, L.COL[name]] == 0).all(), name
            ref[:, L.COL[name]] = 0.0
    assert_parity_with_ties(L, O, P, wf, got, ref)
    assert (ref[:, L.COL["t0"]] > 0).sum() > (200 if n_samples >= 1400 else 50)
