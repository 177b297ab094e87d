"""Pins the CPU oracle against the known-answer tests the reference's own suite holds for the hot-path
primitives (tests/golden/kat_reference_tests.json, transcribed from /root/reference/test/*.jl), against the
closed-form expectations on the reference's fixture waveform and against brute-force
cross-checks.  CPU only."""
import json
import math
import os

import numpy as np
import pytest

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat_reference_tests.json")))


def test_extremestats_kat(O):
    k = KAT["extremestats"]
    y = np.array(k["signal"])
    for c in k["cases"]:
        r = O.extremestats(y, k["t0"], k["dt"], c["from"], c["until"])
        assert r["min"] == c["min"] and r["max"] == c["max"]
        assert r["tmin"] == c["tmin"] and r["tmax"] == c["tmax"]


def test_get_wvf_maximum_kat(O):
    for c in KAT["get_wvf_maximum"]["cases"]:
        y = np.zeros(c["n"])
        for i, v in c["set"].items():
            y[int(i)] = v
        m = O.get_wvf_maximum(y, c["from"], c["until"])
        assert c["ge"] <= m < c["lt"]
        assert m == c["exact"]


def test_derivative_kat(O):
    rng = np.random.default_rng(1)
    x = rng.random(100)
    for gain in (1.0, 0.37):
        y = O.derivative(x, gain)
        ref = gain * np.concatenate([[x[1] - x[0]], np.diff(x)])
        assert np.array_equal(y, ref)  # src/derivative.jl:47-55, test/test_derivative.jl:11-18


def test_intersect_ramp_kat(O):
    k = KAT["intersect_ramp"]
    y = np.array(k["signal"], dtype=float)
    for c in k["cases"]:
        r = O.intersect(y, k["t0"], k["dt"], c["thr"], c["min_n"])
        assert math.isclose(r["x"], c["x"], rel_tol=1e-12)
        assert r["multiplicity"] == 1


def test_intersect_state_machine_kat(O):
    k = KAT["intersect_state_machine"]
    for c in k["cases"]:
        y = np.zeros(k["n"])
        for i, v in c.get("set", {}).items():
            y[int(i)] = v
        for a, b, v in c.get("ranges", []):
            y[a:b + 1] = v
        r = O.intersect(y, 0.0, k["dt"], k["thr"], k["min_n"])
        assert r["multiplicity"] == c["multiplicity"], c["name"]
        if "x_gt" in c:
            assert r["x"] > c["x_gt"]
        if "x_lt" in c:
            assert r["x"] < c["x_lt"]


def test_intersect_edge_rules(O):
    # first sample already above threshold: that run never fires (src/multi_intersect.jl:55)
    y = np.array([5.0, 5, 5, 0, 0, 5, 5, 5, 0])
    r = O.intersect(y, 0.0, 1.0, 1.0, 2)
    assert r["multiplicity"] == 1 and r["pos"] == 5
    assert math.isclose(r["x"], 4.0 + 1.0 / 5.0)
    # run shorter than min_n does not count
    r = O.intersect(np.array([0.0, 5, 0, 5, 5, 5, 0]), 0.0, 1.0, 1.0, 3)
    assert r["multiplicity"] == 1 and r["pos"] == 3
    # nothing found -> NaN
    r = O.intersect(np.zeros(10), 0.0, 1.0, 1.0, 1)
    assert math.isnan(r["x"]) and r["multiplicity"] == 0 and r["pos"] == -1
    # empty input
    r = O.intersect(np.zeros(0), 0.0, 1.0, 1.0, 1)
    assert math.isnan(r["x"]) and r["multiplicity"] == 0


def test_saturation_semantics(O):
    # src/saturation.jl:28-65: counts and longest runs; a `low` sample ends a `high` run and vice versa
    y = np.array([0, 0, 7, 65520, 65520, 65520, 0, 65520, 3, 0, 0, 0], dtype=np.uint16)
    r = O.saturation(y, 0, 65520)
    assert r == dict(low=6, high=4, max_cons_low=3, max_cons_high=3)
    r = O.saturation(np.full(50, 9, dtype=np.uint16), 0, 65520)
    assert r == dict(low=0, high=0, max_cons_low=0, max_cons_high=0)
    r = O.saturation(np.zeros(50, dtype=np.uint16), 0, 65520)
    assert r == dict(low=50, high=0, max_cons_low=50, max_cons_high=0)


def test_tailstats_exact_exponential(O):
    # log of an exact exponential is linear: tau recovered, sigma = sqrt((n^2-1)/12)/tau_samples (App. C)
    n, tau_s, dt = 8192, 31250.0, 16.0
    i = np.arange(n)
    y = 1e4 * np.exp(-(i - 3124) / tau_s)
    r = O.tailstats(y, 0.0, dt, 4375, 6875)
    assert math.isclose(r["tau"], tau_s * dt, rel_tol=1e-9)
    assert math.isclose(r["mean"], math.log(1e4) - (5625 - 3124) / tau_s, rel_tol=1e-12)
    assert math.isclose(r["sigma"], math.sqrt((2501 ** 2 - 1) / 12) / tau_s, rel_tol=1e-7)
    # any sample <= 0 -> zeros (src/tailstats.jl:27-33)
    y[5000] = 0.0
    assert O.tailstats(y, 0.0, dt, 4375, 6875) == dict(mean=0.0, sigma=0.0, tau=0.0)


def test_signalstats_against_numpy(O):
    rng = np.random.default_rng(2)
    y = 1000 + rng.normal(0, 3, 4000)
    a, b = 17, 2455
    r = O.signalstats(y, 0.0, 16.0, a, b)
    x = 16.0 * np.arange(a, b + 1)
    assert math.isclose(r["mean"], y[a:b + 1].mean(), rel_tol=1e-13)
    assert math.isclose(r["sigma"], y[a:b + 1].std(), rel_tol=1e-7)
    slope, offset = np.polyfit(x, y[a:b + 1], 1)
    assert math.isclose(r["slope"], slope, rel_tol=1e-6, abs_tol=1e-12)
    assert math.isclose(r["offset"], offset, rel_tol=1e-9)


def test_trap_running_sum_vs_bruteforce(O):
    rng = np.random.default_rng(3)
    y = np.cumsum(rng.normal(0, 1, 2048)) + 1e4
    for a, g, a2 in ((2, 6, 125), (312, 156, 312), (188, 62, 188), (5, 0, 3)):
        f = O.trap(y, a, g, a2)
        b = O.trap(y, a, g, a2, bruteforce=True)
        assert f.shape == b.shape == (2048 - (a + g + a2) + 1,)
        assert np.allclose(f, b, rtol=0, atol=1e-8)


def test_invcr_closed_form(O):
    # y[i] = y[i-1] + x[i]/alpha - x[i-1]  <=>  y = x + cumsum(x)/RC   
    rng = np.random.default_rng(4)
    x = rng.normal(0, 1, 4096) + 50
    RC = 31250.0
    km1 = 1.0 / (RC / (RC + 1.0)) - 1.0
    y = O.invcr(x, km1)
    assert np.allclose(y, x + np.cumsum(x) / RC, rtol=1e-12)
    # an exponential with the matching decay constant becomes a (nearly) flat step
    i = np.arange(4096)
    p = np.where(i < 100, 0.0, 1e4 * np.exp(-(i - 100) / RC))
    q = O.invcr(p, km1)
    assert abs(q[4000] - q[200]) < 1.0


def test_fixture_closed_form_expectations(L, O, example_params):
    """SURVEY.md Appendix C on make_fake_waveform (test/test_dsp_icpc.jl:11-32), rounded to UInt16"""
    wf = L.synth.generate_host(3, mode=1)
    rows, idx, _ = O.dsp_icpc(example_params, wf, want_idx=True)
    assert np.array_equal(rows[0], rows[1]) and np.array_equal(rows[0], rows[2])
    r = dict(zip(L.COLUMNS, rows[0]))
    assert abs(r["blmean"] - 1000) < 1e-9 and r["blsigma"] < 1e-4 and abs(r["blslope"]) < 1e-12
    assert r["e_max"] == 10000.0 and abs(r["e_min"]) < 1e-9
    assert r["qc_label"] == -1
    assert all(r[k] == 0 for k in ("n_sat_low", "n_sat_high", "n_sat_low_cons", "n_sat_high_cons"))
    assert abs(r["tail_tau"] - 500000.0) < 50.0            # +-0.5 ADC rounding of the samples
    assert abs(r["tail_mean"] - 9.130308) < 1e-5 and abs(r["tail_sigma"] - 0.023103) < 1e-5
    for k in ("tailmean", "e_10410", "e_535", "e_trap", "e_cusp", "e_zac"):
        assert abs(r[k] - 10020.2) < 1.0, k
    assert abs(r["t50"] - 48.983) < 2e-3 and abs(r["t0"] - 48.033) < 2e-3 and abs(r["t90"] - 49.781) < 2e-3
    assert abs(r["drift_time"] - 1748) < 3
    # the reference's own assertions (test/test_dsp_icpc.jl:189-199)
    assert r["t0"] < r["t50"] < r["t90"] and r["drift_time"] >= 0
    assert all(np.isfinite(r[k]) for k in ("e_10410", "e_313", "e_trap"))
    assert abs(r["e_10410_inv"]) < 0.5 and abs(r["e_313_inv"]) < 0.5 and r["t0_inv"] == 0.0
    assert r["inTrace_n"] == 1
    assert idx[0][2] == 3062  # t50 crossing between samples 3062 and 3063 (1-based), App. C


def test_oracle_rows_golden(L, O, example_params):
    """regression pin: oracle output on the 3-event fixture + 8 mixed events (oracle-generated golden file)"""
    path = os.path.join(os.path.dirname(__file__), "golden", "oracle_rows.json")
    g = json.load(open(path))
    wf = np.concatenate([L.synth.generate_host(1, mode=1), L.synth.generate_host(8, first_event=g["first_event"])])
    rows, _ = O.dsp_icpc(example_params, wf)
    ref = np.array(g["rows"], dtype=float)
    assert rows.shape == ref.shape
    both_nan = np.isnan(rows) & np.isnan(ref)
    assert np.allclose(np.where(both_nan, 0, rows), np.where(both_nan, 0, ref), rtol=1e-9, atol=1e-9)


def test_timed_oracle_build_agrees_with_the_strict_one(L, O):
    """bench.py's CPU arms run the -O3 -march=native build of the oracle sources (contraction allowed, the
    reference's second ZAC pass paid for); it must agree with the strict checker build to 1e-9 on every column"""
    from parity import compare_rows
    P = L.resolve_icpc_params(L.example_config(), L.us(500.0), builders=O.OracleBuilders())
    wf = O.synth_generate(96, first_event=77)
    strict, _ = O.dsp_icpc(P, wf)
    O.use_fast_build(True)
    try:
        fast, _ = O.dsp_icpc(P, wf)
        lean = O.pz_trap(P, wf)
    finally:
        O.use_fast_build(False)
    tie_cols = {"a_sg", "a_60", "a_100", "a_raw"}
    bad = {k: v for k, v in compare_rows(fast, strict, L.COLUMNS).items() if v[1] > 0 and k not in tie_cols}
    assert not bad, bad
    for i, k in enumerate(("blmean", "t0", "t50", "e_trap", "e_10410")):
        ref = strict[:, L.COL[k]]
        assert np.allclose(lean[:, i], ref, rtol=1e-9, atol=1e-7), k


def test_oracle_event_generator_matches_the_product_host_generator(L, O):
    """the CPU arms' own statement of the synthetic stream (oracle/lgdsp_synth_oracle.c) == lgdsp_synth_generate_host"""
    for kw in (dict(first_event=0), dict(first_event=123456789), dict(mode=1), dict(n_samples=1024, first_event=5)):
        a = O.synth_generate(40, **kw)
        b = L.synth.generate_host(40, **kw)
        assert np.array_equal(a, b), kw
