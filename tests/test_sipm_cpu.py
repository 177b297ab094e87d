"""The SiPM chain's in-tree primitives of the oracle against the reference's own known-answer tests
(test/test_intersect_maximum.jl, test/test_thresholdstats.jl, test/test_stats.jl:57-110) -- parity PINNED for these --
and the oracle's dsp_sipm on the reference's fixture (test/test_dsp_sipm.jl)."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(HERE, "golden", "kat_reference_tests.json")) as f:
        return json.load(f)


def _signal(n, case):
    y = np.zeros(n)
    for k, v in case.get("set", {}).items():
        y[int(k)] = v
    for a, b, v in case.get("ranges", []):
        y[a:b + 1] = v
    return y


def check_intersect_maximum_case(r, c, x0=0.0):
    assert r["multiplicity"] == c["multiplicity"], c["name"]
    assert len(r["x"]) == len(r["max"]) == len(r["x_high"]) == len(r["x_tot"]) == c["multiplicity"]
    if "x_gt" in c:
        assert r["x"][0] > c["x_gt"], c["name"]
    if "x_lt" in c:
        assert r["x"][0] < c["x_lt"], c["name"]
    if "max_ge" in c:
        assert c["max_ge"] <= r["max"][0] < c["max_lt"], c["name"]
    if "max_eq" in c:
        assert r["max"][0] == c["max_eq"], c["name"]
    if c.get("x_high_gt_x"):
        assert r["x_high"][0] > r["x"][0]
        assert np.isclose(r["x_tot"][0], r["x_high"][0] - r["x"][0])
    if "x_high_eq" in c:
        assert r["x_high"][0] == c["x_high_eq"], c["name"]
    if "tot_gt" in c:
        assert (r["x_tot"] > c["tot_gt"]).all(), c["name"]
    if c.get("tot_increasing"):
        assert r["x_tot"][1] > r["x_tot"][0]


def test_intersect_maximum_kat(O, kat):
    k = kat["intersect_maximum"]
    for c in k["cases"]:
        y = _signal(k["n"], c)
        r = O.intersect_maximum(y, 0.0, k["dt"], k["thr"], k["min_n"], c["max_n"])
        check_intersect_maximum_case(r, c)
    r = O.intersect_maximum(np.zeros(0), 0.0, k["dt"], k["thr"], k["min_n"], 100)
    assert r["multiplicity"] == 0 and len(r["x"]) == 0


def test_intersect_maximum_against_python_transcription(O):
    """the oracle against a direct transcription of the state machine on noisy pulses (all fields, all triggers)"""
    rng = np.random.default_rng(11)
    for trial in range(20):
        n = int(rng.integers(50, 400))
        y = rng.normal(0, 1, n)
        for _ in range(int(rng.integers(0, 5))):
            a = int(rng.integers(0, n - 5))
            y[a:a + int(rng.integers(1, 30))] += rng.uniform(2, 6)
        thr, min_n, max_n = 2.0, int(rng.integers(1, 5)), int(rng.integers(1, 20))
        r = O.intersect_maximum(y, 8.0, 16.0, thr, min_n, max_n)
        ups, counter, cand = [], (min_n + 1 if y[0] > thr else 0), 1
        for i in range(n):
            high = y[i] >= thr
            if high and counter == 0:
                cand = i
            counter = counter + 1 if high else 0
            if counter == min_n and cand > 0:
                ups.append(cand)
        assert r["multiplicity"] == len(ups)
        for q, up in enumerate(ups):
            t = lambda i: 8.0 + 16.0 * i
            x = (thr - y[up - 1]) * (t(up) - t(up - 1)) / (y[up] - y[up - 1]) + t(up - 1)
            assert r["x"][q] == x
            lo, hi = max(up - 2, 0), min(up + max_n, n - 1)
            w = y[lo:hi + 1]
            im = int(np.argmax(w))
            if 0 < im < len(w) - 1:
                a = w[im + 1] - 4 * w[im] + 3 * w[im - 1]
                m = w[im - 1] - a * a / (8 * (w[im + 1] - 2 * w[im] + w[im - 1]))
            else:
                m = w[im]
            assert r["max"][q] == m
            down = next((j for j in range(up + min_n, n) if y[j] < thr), None)
            xh = t(n - 1) if down is None else (thr - y[down - 1]) * 16.0 / (y[down] - y[down - 1]) + t(down - 1)
            assert np.isclose(r["x_high"][q], xh, rtol=1e-15) and np.isclose(r["x_tot"][q], xh - x)


def test_thresholdstats_mad_kat(O, kat):
    for c in kat["thresholdstats_mad"]["cases"]:
        v = O.thresholdstats_mad(np.array(c["signal"]), float(c["min"]), float(c["max"]))
        if "expect" in c:
            assert abs(v - c["expect"]) <= c["atol"], c["name"]
        else:
            assert v < c["lt"], c["name"]
    # definition check: 1.4826 * median(|y - median(y)|) of the samples inside the bounds, Julia's median of an even count
    rng = np.random.default_rng(3)
    y = rng.normal(0, 2.0, 1001)
    for mn, mx in ((-np.inf, np.inf), (-1.0, 1.5), (0.0, 0.1)):
        f = y[(y >= mn) & (y <= mx)]
        med = np.sort(f)[len(f) // 2] if len(f) % 2 else np.sort(f)[len(f) // 2 - 1] / 2 + np.sort(f)[len(f) // 2] / 2
        d = np.sort(np.abs(f - med))
        mad = d[len(d) // 2] if len(d) % 2 else d[len(d) // 2 - 1] / 2 + d[len(d) // 2] / 2
        assert O.thresholdstats_mad(y, mn, mx) == 1.4826 * mad


def test_thresholdstats_kat(O, kat):
    k = kat["thresholdstats"]
    rng = np.random.default_rng(k["seed"])
    sigma = 10.0 * rng.random()
    y = sigma * rng.standard_normal(k["n"])
    assert np.isclose(O.thresholdstats(y), sigma, rtol=0.05)
    assert np.isclose(O.thresholdstats(y), np.std(y, ddof=1), rtol=k["rtol_all"])
    for _ in range(200):
        mn, mx = -sigma * rng.random(), sigma * rng.random()
        sel = y[(y >= mn) & (y <= mx)]
        if len(sel) > 50:
            assert np.isclose(O.thresholdstats(y, mn, mx), np.std(sel, ddof=1), rtol=k["rtol_bounds"])


def sipm_fixture(n=6250):
    """make_sipm_waveform  test/test_dsp_sipm.jl:10-27 (1-based i)"""
    i = np.arange(1, n + 1)
    ps, pw, amp, tau = round(50e3 / 16), 10, 5.0, 30.0
    sig = np.zeros(n)
    rise = (i >= ps) & (i < ps + pw)
    sig[rise] = amp * (1 - np.exp(-(i[rise] - ps) / 3.0))
    dec = i >= ps + pw
    sig[dec] = amp * np.exp(-(i[dec] - ps - pw) / tau)
    return sig


def test_oracle_dsp_sipm_on_reference_fixture(L, O):
    """test/test_dsp_sipm.jl:70-109: table shape, thresholds finite and >= 0, times inside the waveform range"""
    P = L.resolve_sipm_params(L.example_sipm_config(), {"sg": {"wl": L.ns(200.0)}}, n_samples=6250, sample_kind="f32",
                              builders=O.OracleBuilders())
    wf = np.tile(sipm_fixture().astype(np.float32), (10, 1))
    rows, trig = O.dsp_sipm(P, wf)
    c = L._abi.SIPM_COL
    assert rows.shape == (10, L._abi.SIPM_NCOL)
    for name in ("threshold", "threshold_trap"):
        assert np.isfinite(rows[:, c[name]]).all() and (rows[:, c[name]] >= 0).all()
    for name in ("t_max", "t_min"):
        assert ((rows[:, c[name]] >= 0.0) & (rows[:, c[name]] <= 100.0)).all()
    # noise-free fixture: MAD = 0 -> threshold 0 -> every sample is "high": no up-crossing after the first sample
    assert np.all(rows == rows[0])
    assert np.isclose(rows[0, c["e_max"]], sipm_fixture().astype(np.float32).max())
    assert abs(rows[0, c["t_max"]] - (round(50e3 / 16) + 9 - 1) * 16e-3) < 0.02
