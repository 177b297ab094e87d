"""Randomised-configuration parity of the fused dsp_icpc kernel: the reference's example config
(test/test_dsp_icpc.jl:50-161) is only one point of the parameter space a LEGEND channel config spans.  Every seed draws
a different DSPConfig (windows, filter lengths, thresholds, interpolation orders), decay constant, optimised filter
parameters (`pars_filter`, src/utils.jl:72-82), sampling step and sample count; the C-ABI result must equal the
float64 oracle on the same seeded waveforms (structured CUSP/ZAC evaluation against the oracle's direct FIRs).
LGDSP_FUZZ_SEEDS=<n> (and _SWEEP / _SIPM / _COMPRESSED) widen the sweeps; 400 / 80 / 80 / 40 seeds were run green on the
B200 in round 1."""
import os
from importlib import import_module

import numpy as np
import pytest

from test_gpu_icpc import assert_parity_with_ties

pytestmark = pytest.mark.gpu


def _draw(L, seed):
    cfgm = import_module("legenddsp.jl_b200.config")
    rng = np.random.default_rng(1000 + seed)
    us, ns = L.us, L.ns
    step_ns = float(rng.choice([16.0, 16.0, 8.0, 12.5]))
    n = int(rng.choice([8192, 8192, 7000, 6144]))
    scale = step_ns / 16.0                     # the synthetic pulse starts near sample 3000: windows follow in samples

    def u(lo, hi, q=None):
        """uniform in [lo, hi] microseconds (of the 16 ns layout), optionally snapped to multiples of q ns"""
        v = rng.uniform(lo, hi) * 1000.0 * scale
        if q:
            v = round(v / q) * q
        return ns(v)

    d = cfgm.example_config_dict()
    d["bl_window"] = {"min": ns(0.0), "max": u(20.0, 41.0, step_ns)}
    tail_lo = rng.uniform(62.0, 80.0)
    t_end_us = (n - 1) * 16.0 / 1000.0
    d["tail_window"] = {"min": u(tail_lo, tail_lo, step_ns), "max": u(min(tail_lo + 10.0, t_end_us - 1.0), t_end_us - 0.5, step_ns)}
    d["current_window"] = {"min": u(42.0, 45.0, step_ns), "max": u(55.0, 64.0, step_ns)}
    flt_len = rng.uniform(24.0, 44.0)
    d["flt_length_cusp"] = u(flt_len, flt_len, 2 * step_ns)
    d["flt_length_zac"] = d["flt_length_cusp"] if seed % 2 == 0 else u(24.0, 44.0, 2 * step_ns)
    d["t0_threshold"] = float(rng.uniform(2.5, 8.0))
    d["inTraceCut_std_threshold"] = float(rng.uniform(3.5, 7.0))
    d["sg_flt_degree"] = int(rng.choice([2, 3]))
    q1 = rng.uniform(1.0, 3.0)
    d["qdrift_int_length"] = (u(q1, q1), u(q1 + 1.0, q1 + 4.0))
    l1 = rng.uniform(1.0, 3.0)
    d["lq_int_length"] = (u(l1, l1), u(l1 + 1.0, l1 + 4.0))
    kw = d["kwargs_pars"]
    kw["t0_flt_pars"] = [ns(step_ns * int(rng.integers(2, 6))), ns(step_ns * int(rng.integers(3, 12))),
                         ns(float(rng.uniform(1000.0, 3000.0)))]
    kw["t0_mintot"] = ns(float(rng.uniform(200.0, 3000.0)) * scale)
    kw["tx_mintot"] = ns(float(rng.uniform(16.0, 120.0)) * scale)
    kw["intrace_mintot"] = ns(float(rng.uniform(40.0, 300.0)) * scale)
    kw["int_interpolation_order"] = int(rng.integers(1, 4))
    kw["int_interpolation_length"] = ns(step_ns * int(rng.integers(4, 12)))
    kw["sig_interpolation_order"] = int(rng.integers(1, 4))
    kw["sig_interpolation_length"] = ns(step_ns * int(rng.integers(8, 60)))
    cfg = cfgm.DSPConfig.from_dict(d)
    pars_filter = {
        "trap": {"rt": u(1.0, 12.0), "ft": u(0.5, 4.0)},
        "cusp": {"rt": u(2.0, 12.0), "ft": u(0.5, 4.0)},
        "zac": {"rt": u(2.0, 12.0), "ft": u(0.5, 4.0)},
        "sg": {"wl": ns(float(rng.uniform(60.0, 400.0)) * scale)},
    }
    if seed % 2 == 0:                          # shared CUSP/ZAC parameters: the kernel's one-pass variant
        pars_filter["zac"] = dict(pars_filter["cusp"])
    tau = us(float(rng.uniform(150.0, 900.0)) * scale)
    return cfg, tau, pars_filter, n, ns(step_ns)


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("LGDSP_FUZZ_SEEDS", "10")))))
def test_random_configurations(L, O, handle, seed):
    cfg, tau, pars_filter, n, step = _draw(L, seed)
    P = L.resolve_icpc_params(cfg, tau, pars_filter, n_samples=n, step=step, builders=O.OracleBuilders())
    wf = np.ascontiguousarray(L.synth.generate_host(160, first_event=7000 * (seed + 1))[:, :n])
    got = L.dsp_icpc_rows(wf, P, handle=handle)
    ref, _ = O.dsp_icpc(P, wf)
    # (exact argmax ties of the current traces -- equal integer sample differences -- show up in 1-2 % of the events; every
    # mismatch must still be explained by a tie, sample by sample)
    res, n_ties = assert_parity_with_ties(L, O, P, wf, got, ref, max_ties=8)
    c = L.COL
    # the population is not degenerate (high t0 thresholds leave small pulses without a t0)
    assert np.isfinite(ref[:, c["e_trap"]]).sum() > 100 and (ref[:, c["t0"]] > 0).sum() > 4


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("LGDSP_FUZZ_SEEDS_SWEEP", "6")))))
def test_random_trap_sweeps(L, O, handle, seed):
    """dsp_trap_ft-style grids with random (rt, ft) lists under random configurations: the one-thread-per-variant path,
    the t50 chunk pruning and the pick-off windows near the trace ends (clamped DNI windows) against the oracle"""
    cfg, tau, _, n, step = _draw(L, 100 + seed)
    rng = np.random.default_rng(5000 + seed)
    sc = step.ns() / 16.0
    rts = [L.ns(float(v) * 1000.0 * sc) for v in np.sort(rng.uniform(0.3, 30.0, int(rng.integers(3, 24))))]
    fts = [L.ns(float(v) * 1000.0 * sc) for v in np.sort(rng.uniform(0.1, 6.0, int(rng.integers(2, 12))))]
    wf = np.ascontiguousarray(L.synth.generate_host(96, first_event=300 * (seed + 1))[:, :n])
    so = L.resolve_sweep_params(cfg, tau, n_samples=n, step=step, builders=O.OracleBuilders())
    # variants whose trapezoid leaves fewer outputs than the DNI window are rejected by the library: keep the others
    keep_r = [r for r in rts if 2 * round(r.ns() / step.ns()) + round(fts[-1].ns() / step.ns()) + so.sig_dni.n_w < n]
    var = L.trap_variants(keep_r, fts, step, mode="ft")
    got = L.dsp_trap_rtft_grid(L.RDWaveforms(wf, L.ns(0.0), step), cfg, tau, keep_r, fts, handle=handle)
    ref = O.trap_sweep(so, wf, var)
    assert got.shape == (len(keep_r), len(fts), 96)
    assert np.allclose(got.reshape(len(keep_r) * len(fts), 96).T, ref, rtol=1e-6, atol=1e-6, equal_nan=True)
    assert np.isfinite(ref).mean() > 0.9


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("LGDSP_FUZZ_SEEDS_SIPM", "6")))))
def test_random_sipm_configurations(L, O, handle, seed):
    """dsp_sipm (src/dsp_sipm.jl:47-158) under random filter / threshold / window parameters and trace lengths"""
    from test_gpu_sipm import _compare, sipm_population
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.choice([6250, 5000, 8192, 3000]))
    t_end_us = (n - 1) * 0.016
    lo = float(rng.uniform(2.0, 0.5 * t_end_us))
    cfg = L.example_sipm_config()
    cfg["t0_hpge_window"] = (L.us(lo), L.us(float(rng.uniform(lo + 1.0, t_end_us))))
    cfg["sg_flt_degree"] = int(rng.choice([2, 3]))
    s = float(rng.uniform(2.5, 6.0))
    cfg["filters"]["sg"].update({"n_σ_threshold": s, "n_σ_dc_threshold": s + float(rng.uniform(0.5, 4.0))},
                                min_threshold=-float(rng.uniform(2.0, 6.0)), max_threshold=float(rng.uniform(2.0, 6.0)),
                                min_dc_threshold=-float(rng.uniform(20.0, 60.0)), max_dc_threshold=float(rng.uniform(20.0, 60.0)),
                                min_tot_intersect=L.ns(float(rng.uniform(16.0, 120.0))),
                                max_tot_intersect=L.ns(float(rng.uniform(130.0, 600.0))))
    s = float(rng.uniform(2.5, 6.0))
    cfg["filters"]["trap"].update({"n_σ_threshold": s, "n_σ_dc_threshold": s + float(rng.uniform(0.5, 4.0))},
                                  rt=L.ns(16.0 * int(rng.integers(2, 16))), ft=L.ns(16.0 * int(rng.integers(1, 8))),
                                  pz_tau=L.us(float(rng.uniform(0.5, 20.0))),
                                  min_threshold=-float(rng.uniform(8.0, 25.0)), max_threshold=float(rng.uniform(8.0, 25.0)),
                                  min_dc_threshold=-float(rng.uniform(20.0, 50.0)), max_dc_threshold=float(rng.uniform(20.0, 50.0)),
                                  min_tot_intersect=L.ns(float(rng.uniform(16.0, 100.0))),
                                  max_tot_intersect=L.ns(float(rng.uniform(110.0, 500.0))))
    wl = L.ns(float(rng.uniform(60.0, 500.0)))
    P = L.resolve_sipm_params(cfg, {"sg": {"wl": wl}}, n_samples=n, builders=O.OracleBuilders(), max_triggers=64)
    wf = sipm_population(96, n=n, seed=40 + seed)
    rows, trig = L.sipm_rows(wf, P, handle=handle)
    ref_rows, ref_trig = O.dsp_sipm(P, wf)          # (sipm_rows grows P.max_triggers in place when a list was cut)
    _compare(L, rows, trig, ref_rows, ref_trig)


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("LGDSP_FUZZ_SEEDS_COMPRESSED", "4")))))
def test_random_compressed_configurations(L, O, handle, seed):
    """dsp_icpc_compressed (src/dsp_icpc.jl:293-499) with random presum rates, window placements and configurations"""
    from test_gpu_compressed import _check, _data
    from parity import TOL
    rng = np.random.default_rng(300 + seed)
    presum = int(rng.choice([2, 4, 8]))
    # the full-rate window must hold the current window (43 .. 62 us = samples 2688 .. 3875) plus the filter lengths
    w0 = int(rng.integers(2300, 2650))
    wlen = int((3950 - w0 + rng.integers(0, 300)) // 8 * 8)
    tau_c = L.us(float(rng.uniform(200.0, 800.0)))
    cfg = L.tiefree_config() if seed == 0 else L.example_config()
    n_events = 256
    data = _data(L, n_events, 11000 * (seed + 1), presum, window=(w0, wlen))
    res = L.dsp_icpc_compressed(data, cfg, tau_c, None, handle=handle, builders=O.OracleBuilders())
    wp, ww = data["waveform_presummed"], data["waveform_windowed"]
    Pp, Pw, aux = L.resolve_compressed_params(cfg, tau_c, None, presum_rate=presum, n_pre=wp.signal.shape[1],
                                              t_first_pre=wp.t_first, step_pre=wp.step, n_wdw=ww.signal.shape[1],
                                              t_first_wdw=ww.t_first, step_wdw=ww.step, builders=O.OracleBuilders())
    ref = O.dsp_icpc_compressed(Pp, Pw, wp.signal, ww.signal, presum, aux)
    bad = _check(res, ref, skip=("a_sg", "a_60", "a_100", "a_raw"))
    assert not bad, bad
    for col in ("a_sg", "a_60", "a_100", "a_raw"):
        rtol, atol = TOL[col]
        n_bad = int((np.abs(res[col] - ref[col]) > atol + rtol * np.abs(ref[col])).sum())
        assert n_bad <= max(2, n_events // 50), (col, n_bad)


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("LGDSP_FUZZ_SEEDS_MI", "8")))))
def test_random_multi_intersect(L, O, handle, seed):
    """MultiIntersect (src/multi_intersect.jl:36-104) with random threshold lists, time-over-threshold requirements (1 .. 70
    samples: runs inside one 32-sample block, across blocks, longer than a block), polynomial windows and trace lengths;
    pulses on noise, traces that start above the thresholds, plateaus at a threshold"""
    rng = np.random.default_rng(7000 + seed)
    n = int(rng.integers(700, 8193))
    n_ev = 48
    hw, d = int(rng.integers(1, 5)), int(rng.integers(1, 4))
    d = min(d, 2 * hw - 1)
    rate = int(rng.choice([1, 2, 4, 8]))
    min_n = int(rng.choice([1, 2, 3, 5, 9, 17, 31, 32, 33, 40, 70]))
    nthr = int(rng.integers(1, 100))
    ratios = np.sort(rng.uniform(0.02, 0.95, nthr))
    kk = np.arange(n)
    Y = np.empty((n_ev, n))
    for e in range(n_ev):
        s0 = int(rng.integers(100 + min_n, n // 2))
        rise = int(rng.integers(max(8, 2 * min_n), max(10, 2 * min_n) + 400))
        amp = rng.uniform(200, 5000)
        Y[e] = amp * np.clip((kk - s0) / rise, 0, 1) + rng.normal(0, 1.0, n)
        if e % 3 == 1:
            Y[e, :int(rng.integers(1, 90))] += rng.uniform(0.1, 0.9) * amp         # starts above some thresholds
        if e % 3 == 2:
            Y[e] = np.round(Y[e] / (amp / 16)) * (amp / 16)                          # staircase: exact ties with thresholds
    f = L.MultiIntersect(threshold_ratios=list(ratios), mintot=L.ns(16.0 * min_n), n=hw, d=d, sampling_rate=rate)
    n_checked = 0
    for e0 in range(0, n_ev, 16):
        batch = Y[e0:e0 + 16]
        refs, ok_rows = [], []
        for e in range(len(batch)):
            try:
                refs.append(O.multi_intersect(batch[e], 4.0, 16.0, f.threshold_ratios, min_n, hw, d, rate))
                ok_rows.append(e)
            except AssertionError:                       # boundary assertion of the reference (:85-88): not part of this test
                pass
        if not ok_rows:
            continue
        got = f(np.ascontiguousarray(batch[ok_rows]), t_first=L.ns(4.0), step=L.ns(16.0), handle=handle, builders=O.OracleBuilders())
        for g, r in zip(got, refs):
            assert np.array_equal(np.isnan(g), np.isnan(r))
            ok = ~np.isnan(r)
            assert np.allclose(g[ok], r[ok], rtol=0, atol=1e-6)
            n_checked += 1
    assert n_checked >= n_ev // 2
