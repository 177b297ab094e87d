"""SiPM trigger chain on the GPU: the primitives against the reference's own known-answer tests and
the oracle, dsp_sipm against the oracle's restatement of src/dsp_sipm.jl:47-158."""
import json
import os

import numpy as np
import pytest

from test_sipm_cpu import _signal, check_intersect_maximum_case, sipm_fixture

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(HERE, "golden", "kat_reference_tests.json")) as f:
        return json.load(f)


def test_intersect_maximum_kat_on_gpu(L, handle, kat):
    """test/test_intersect_maximum.jl:6-107 through lgdsp_intersect_maximum"""
    k = kat["intersect_maximum"]
    for c in k["cases"]:
        y = _signal(k["n"], c)
        f = L.IntersectMaximum(L.ns(k["min_n"] * k["dt"]), L.ns(c["max_n"] * k["dt"]))
        r = f(y, k["thr"], step=L.ns(k["dt"]), handle=handle)
        check_intersect_maximum_case(r, c)
    r = L.IntersectMaximum(L.ns(32.0), L.ns(1600.0))(np.zeros(0), 0.4, handle=handle)
    assert r["multiplicity"] == 0 and len(r["x"]) == 0 and len(r["x_high"]) == 0 and len(r["max"]) == 0      # :85-93


def test_thresholdstats_kat_on_gpu(L, handle, kat):
    """test/test_thresholdstats.jl:7-65 and test/test_stats.jl:57-110 through lgdsp_thresholdstats"""
    for c in kat["thresholdstats_mad"]["cases"]:
        v = L.thresholdstats_mad(np.array(c["signal"]), float(c["min"]), float(c["max"]), handle=handle)
        if "expect" in c:
            assert abs(v - c["expect"]) <= c["atol"], c["name"]
        else:
            assert v < c["lt"], c["name"]
    k = kat["thresholdstats"]
    rng = np.random.default_rng(k["seed"])
    sigma = 10.0 * rng.random()
    y = sigma * rng.standard_normal(k["n"])
    assert np.isclose(L.thresholdstats(y, handle=handle), sigma, rtol=0.05)
    assert np.isclose(L.thresholdstats(y, handle=handle), np.std(y, ddof=1), rtol=k["rtol_all"])
    for _ in range(50):
        mn, mx = -sigma * rng.random(), sigma * rng.random()
        sel = y[(y >= mn) & (y <= mx)]
        if len(sel) > 50:
            assert np.isclose(L.thresholdstats(y, mn, mx, handle=handle), np.std(sel, ddof=1), rtol=k["rtol_bounds"])


def test_primitives_against_oracle(L, O, handle):
    """random and degenerate traces: medians are exact order statistics, trigger lists identical"""
    rng = np.random.default_rng(2024)
    traces = [rng.normal(0, 1.0, 6238), rng.normal(0, 1e-3, 777), np.round(rng.normal(0, 2.0, 5000)),     # heavy ties
              np.zeros(1000), np.full(333, 2.5), np.concatenate([np.zeros(3000), np.full(3001, 1.0)]),
              rng.standard_cauchy(4096), np.arange(1000.0), np.array([1.0]), np.array([3.0, -1.0])]
    for y in traces:
        for mn, mx in ((-np.inf, np.inf), (-1.0, 1.0), (0.0, 0.5), (5.0, 6.0)):
            assert L.thresholdstats_mad(y, mn, mx, handle=handle) == O.thresholdstats_mad(y, mn, mx), (len(y), mn, mx)
    for trial in range(12):
        n = int(rng.integers(40, 9000))
        y = rng.normal(0, 1, n)
        for _ in range(int(rng.integers(0, 30))):
            a = int(rng.integers(0, n - 5))
            y[a:a + int(rng.integers(1, 60))] += rng.uniform(1.5, 6)
        thr, min_n, max_n = float(rng.uniform(1.0, 3.0)), int(rng.integers(1, 6)), int(rng.integers(1, 40))
        ref = O.intersect_maximum(y, 8.0, 16.0, thr, min_n, max_n, cap=4096)
        got = handle.intersect_maximum(y, 8.0, 16.0, thr, min_n, max_n, 4096)
        assert got["multiplicity"] == ref["multiplicity"]
        assert np.array_equal(got["x"], ref["x"]) and np.array_equal(got["x_high"], ref["x_high"])
        assert np.array_equal(got["x_tot"], ref["x_tot"])
        assert np.allclose(got["max"], ref["max"], rtol=1e-13, atol=1e-13)
    # capacity smaller than the multiplicity: the true count is reported, the first entries are kept
    y = np.tile(np.array([0.0, 0.0, 3.0, 3.0, 3.0, 0.0]), 100)
    ref = O.intersect_maximum(y, 0.0, 16.0, 1.0, 2, 5, cap=1024)
    got = handle.intersect_maximum(y, 0.0, 16.0, 1.0, 2, 5, 7)
    assert got["multiplicity"] == ref["multiplicity"] == 100 and np.array_equal(got["x"], ref["x"][:7])


def sipm_population(n_events, n=6250, seed=7):
    """raw UInt16 SiPM-like traces: baseline + white noise + 0..6 photo-electron pulses + occasional negative discharge"""
    rng = np.random.default_rng(seed)
    k = np.arange(n)
    wf = np.empty((n_events, n), dtype=np.uint16)
    for e in range(n_events):
        y = 2000.0 + rng.normal(0, 2.0, n)
        for _ in range(int(rng.integers(0, 7))):
            s0, amp = int(rng.integers(100, n - 400)), rng.uniform(15, 120)
            d = k - s0
            y += np.where(d >= 0, amp * (1 - np.exp(-np.maximum(d, 0) / 3.0)) * np.exp(-np.maximum(d, 0) / 30.0), 0.0)
        if e % 5 == 0:
            s0 = int(rng.integers(500, n - 600))
            y -= np.where((k >= s0) & (k < s0 + 40), 60.0, 0.0)
        wf[e] = np.clip(np.rint(y), 0, 65535).astype(np.uint16)
    return wf


def _compare(L, rows, trig, ref_rows, ref_trig, sg_only=False):
    c = L._abi.SIPM_COL
    counts = [c[k] for k in (("n_trig",) if sg_only else ("n_trig", "n_trig_DC", "n_trig_trap", "n_trig_DC_trap"))]
    assert np.array_equal(rows[:, counts], ref_rows[:, counts])
    exact = ("t_max", "t_min", "t_max_lar", "t_min_lar", "e_max", "e_min", "e_max_lar", "e_min_lar", "threshold")
    for name in exact:
        assert np.array_equal(rows[:, c[name]], ref_rows[:, c[name]]), name
    for name in ("blmean", "blsigma", "blslope", "bloffset", "wfmean", "wfsigma", "wfslope", "wfoffset", "threshold_DC",
                 "threshold_trap", "threshold_DC_trap"):
        a, b = rows[:, c[name]], ref_rows[:, c[name]]
        assert np.allclose(a, b, rtol=1e-9, atol=1e-9 * (1.0 + np.abs(ref_rows[:, c["wfmean"]]).max())), name
    # SG trigger list: the SG trace is bit-identical, so positions are too
    assert np.array_equal(trig[:, 0, :3, :], ref_trig[:, 0, :3, :])
    assert np.allclose(trig[:, 0, 3, :], ref_trig[:, 0, 3, :], rtol=1e-12, atol=1e-12)
    # lists on the integrated / trapezoid traces: summation order differs (parallel scans): times to 1e-6 ns
    for lst in (() if sg_only else (1, 2, 3)):
        assert np.allclose(trig[:, lst, :3, :], ref_trig[:, lst, :3, :], rtol=0, atol=1e-6), lst
        assert np.allclose(trig[:, lst, 3, :], ref_trig[:, lst, 3, :], rtol=1e-9, atol=1e-7), lst


def test_dsp_sipm_reference_fixture(L, O, handle):
    """the reference's own test (test/test_dsp_sipm.jl:70-109): 10 identical noise-free events of 6250 samples"""
    N = 10
    data = {"waveform": L.RDWaveforms(np.tile(sipm_fixture(), (N, 1))), "baseline": np.zeros(N, np.float32),
            "timestamp": np.zeros(N, np.uint64), "eventnumber": np.arange(1, N + 1, dtype=np.uint32),
            "daqenergy": np.zeros(N, np.uint16)}
    res = L.dsp_sipm(data, L.example_sipm_config(), {"sg": {"wl": L.ns(200.0)}}, handle=handle, builders=O.OracleBuilders())
    expected = ["blfc", "timestamp", "eventID_fadc", "e_fc", "t_max", "t_min", "t_max_lar", "t_min_lar", "e_max", "e_min",
                "e_max_lar", "e_min_lar", "blmean", "blsigma", "blslope", "bloffset", "wfmean", "wfsigma", "wfslope", "wfoffset",
                "threshold", "threshold_DC", "trig_pos", "trig_max", "trig_pos_DC", "trig_max_DC", "threshold_trap",
                "threshold_DC_trap", "trig_pos_trap", "trig_pos_high_trap", "trig_pos_tot_trap", "trig_max_trap",
                "trig_pos_DC_trap", "trig_pos_high_DC_trap", "trig_pos_tot_DC_trap", "trig_max_DC_trap"]
    assert list(res.keys()) == expected                                          # :80-92 (+ the reference's column order)
    assert all(len(v) == N for v in res.values())                                # :77
    assert (res["timestamp"] == 0).all() and np.array_equal(res["eventID_fadc"], np.arange(1, N + 1))   # :95-96
    for k in ("threshold", "threshold_trap"):
        assert np.isfinite(res[k]).all() and (res[k] >= 0).all()                 # :99-102
    for k in ("t_max", "t_min"):
        assert ((res[k] >= 0.0) & (res[k] <= 100.0)).all()                       # :105-106
    P = L.resolve_sipm_params(L.example_sipm_config(), {"sg": {"wl": L.ns(200.0)}}, n_samples=6250, sample_kind="f32",
                              builders=O.OracleBuilders())
    wf = np.tile(sipm_fixture().astype(np.float32), (N, 1))
    rows, trig = L.sipm_rows(wf, P, handle=handle)
    ref_rows, ref_trig = O.dsp_sipm(P, wf)
    # noise-free input: every MAD threshold is exactly 0, so the triggers on the integrated / trapezoid traces compare
    # rounding noise (+-1e-17 around an exact 0) with 0 -- undefined in the reference as well; the SG list is well defined
    _compare(L, rows, trig, ref_rows, ref_trig, sg_only=True)


@pytest.mark.parametrize("n_samples", [6250, 8192, 2000])
def test_dsp_sipm_parity(L, O, handle, n_samples):
    wf = sipm_population(192, n=n_samples, seed=n_samples)
    cfg = L.example_sipm_config()
    if n_samples == 2000:
        cfg["t0_hpge_window"] = (L.us(4.0), L.us(9.0))
    # raw ADC units: bounds of the threshold estimate scaled to the noise of the population
    cfg["filters"]["sg"].update(min_threshold=-3.0, max_threshold=3.0, min_dc_threshold=-40.0, max_dc_threshold=40.0)
    cfg["filters"]["trap"].update(min_threshold=-15.0, max_threshold=15.0, min_dc_threshold=-30.0, max_dc_threshold=30.0)
    P = L.resolve_sipm_params(cfg, {"sg": {"wl": L.ns(200.0)}}, n_samples=n_samples, builders=O.OracleBuilders(), max_triggers=64)
    rows, trig = L.sipm_rows(wf, P, handle=handle)
    ref_rows, ref_trig = O.dsp_sipm(P, wf)
    _compare(L, rows, trig, ref_rows, ref_trig)
    c = L._abi.SIPM_COL
    assert (ref_rows[:, c["n_trig"]] > 0).sum() > 100 and (ref_rows[:, c["n_trig_trap"]] > 0).sum() > 100
    assert (ref_rows[:, c["n_trig_DC"]] > 0).any()
    again, trig2 = L.sipm_rows(wf, P, handle=handle)
    assert np.array_equal(again, rows) and np.array_equal(trig2, trig)            # deterministic
    # table form: VectorOfVectors with the reference's element pointers
    tbl = L.sipm_to_table(rows, trig)
    e = int(np.argmax(ref_rows[:, c["n_trig"]]))
    assert len(tbl["trig_pos"][e]) == int(ref_rows[e, c["n_trig"]])
    assert np.array_equal(tbl["trig_pos"][e], ref_trig[e, 0, 0, :len(tbl["trig_pos"][e])])


def test_dsp_sipm_capacity_and_errors(L, O, handle):
    wf = sipm_population(16, seed=3)
    cfg = L.example_sipm_config()
    cfg["filters"]["sg"].update(min_threshold=-3.0, max_threshold=3.0, n_σ_threshold=1.0)     # low threshold: many triggers
    P = L.resolve_sipm_params(cfg, {"sg": {"wl": L.ns(200.0)}}, n_samples=6250, builders=O.OracleBuilders(), max_triggers=4)
    rows, trig = L.sipm_rows(wf, P, handle=handle)          # grows the capacity until every list fits
    assert P.max_triggers > 4 and rows[:, L._abi.SIPM_COL["n_trig"]].max() <= P.max_triggers
    ref_rows, ref_trig = O.dsp_sipm(P, wf)
    _compare(L, rows, trig, ref_rows, ref_trig)
    with pytest.raises(AssertionError):
        c2 = L.example_sipm_config()
        c2["t0_hpge_window"] = (L.us(200.0), L.us(300.0))
        L.resolve_sipm_params(c2, {"sg": {"wl": L.ns(200.0)}}, n_samples=6250)
    bad = L.resolve_sipm_params(cfg, {"sg": {"wl": L.ns(200.0)}}, n_samples=6250)
    bad.max_triggers = 0
    with pytest.raises(L.LgdspError):
        L.sipm_rows(wf, bad, handle=handle)
    empty = L.dsp_sipm({"waveform": L.RDWaveforms(np.zeros((0, 6250), np.uint16))}, cfg, {"sg": {"wl": L.ns(200.0)}}, handle=handle)
    assert len(empty["t_max"]) == 0 and len(empty["trig_pos"]) == 0


def test_multi_intersect(L, O, handle, kat):
    """MultiIntersect: the reference's tests (test/test_multiintersect.jl:8-26: linear ramp, thresholds 10..90 % -> 10..90 s;
    one ratio == Intersect) and the oracle on noisy pulses with polynomial up-sampling"""
    k = kat["intersect_ramp"]
    ramp = np.array(k["signal"], dtype=np.float64)
    f = L.MultiIntersect(threshold_ratios=[0.1 * q for q in range(1, 10)], mintot=L.ns(1.0))
    x = f(ramp, t_first=L.ns(k["t0"]), step=L.ns(k["dt"]), handle=handle, builders=O.OracleBuilders())
    assert np.allclose(x, [c["x"] for c in k["cases"]])                                   # :18-24
    x1 = L.MultiIntersect(threshold_ratios=[0.5], mintot=L.ns(1.0))(ramp, t_first=L.ns(1.0), step=L.ns(1.0), handle=handle)
    assert np.isclose(x1[0], O.intersect(ramp, 1.0, 1.0, 50.0, 1)["x"])                    # :9-14
    # batch of noisy PZ-like pulses, default ratios 0.01:0.01:0.9, several (n, d, rate)
    rng = np.random.default_rng(8)
    n_ev, n = 96, 2000
    kk = np.arange(n)
    Y = np.empty((n_ev, n))
    for e in range(n_ev):
        s0, rise, amp = rng.integers(300, 900), rng.integers(10, 120), rng.uniform(50, 5000)
        Y[e] = amp * np.clip((kk - s0) / rise, 0, 1) + rng.normal(0, 1.0, n)
    for (hw, d, rate, mintot) in ((1, 1, 1, 64.0), (2, 2, 4, 32.0), (4, 3, 8, 16.0)):
        f = L.MultiIntersect(mintot=L.ns(mintot), n=hw, d=d, sampling_rate=rate)
        got = f(Y, t_first=L.ns(8.0), step=L.ns(16.0), handle=handle, builders=O.OracleBuilders())
        for e in range(n_ev):
            ref = O.multi_intersect(Y[e], 8.0, 16.0, f.threshold_ratios, max(1, round(mintot / 16.0)), hw, d, rate)
            assert np.array_equal(np.isnan(got[e]), np.isnan(ref)), (hw, e)
            ok = ~np.isnan(ref)
            assert np.allclose(got[e][ok], ref[ok], rtol=0, atol=1e-6), (hw, e, np.abs(got[e][ok] - ref[ok]).max())
        assert np.isfinite(got).mean() > 0.9
    # long time-over-threshold requirements (runs that span several 32-sample blocks), traces that START above the
    # thresholds (:56: such a run never fires) and a length that is not a multiple of the 32-sample blocks
    n = 1777
    kk = np.arange(n)
    Y = np.empty((n_ev, n))
    for e in range(n_ev):
        s0, rise, amp = rng.integers(500, 900), rng.integers(60, 200), rng.uniform(200, 5000)
        Y[e] = amp * np.clip((kk - s0) / rise, 0, 1) + rng.normal(0, 1.0, n)
        if e % 2:
            Y[e, :rng.integers(1, 150)] += 0.6 * amp
    for (hw, d, rate, mintot) in ((2, 2, 4, 640.0), (1, 1, 2, 192.0), (3, 3, 4, 528.0)):
        f = L.MultiIntersect(mintot=L.ns(mintot), n=hw, d=d, sampling_rate=rate)
        got = f(Y, t_first=L.ns(0.0), step=L.ns(16.0), handle=handle, builders=O.OracleBuilders())
        for e in range(n_ev):
            ref = O.multi_intersect(Y[e], 0.0, 16.0, f.threshold_ratios, max(1, round(mintot / 16.0)), hw, d, rate)
            assert np.array_equal(np.isnan(got[e]), np.isnan(ref)), (hw, e)
            ok = ~np.isnan(ref)
            assert np.allclose(got[e][ok], ref[ok], rtol=0, atol=1e-6), (hw, e, np.abs(got[e][ok] - ref[ok]).max())
        assert np.isfinite(got).mean() > 0.9
    # boundary assertion (:85-88): a trace whose last threshold is never reached keeps the default position 2 -> with n = 2
    # the left boundary check fails
    flat = np.zeros((1, 500)); flat[0, 0] = 1.0
    with pytest.raises(AssertionError):
        L.MultiIntersect(threshold_ratios=[0.5, 2.0], mintot=L.ns(16.0), n=2, d=1)(flat, handle=handle)
    with pytest.raises(AssertionError):
        O.multi_intersect(flat[0], 0.0, 16.0, [0.5, 2.0], 1, 2, 1, 1)


def test_trigger_lists_compacted_on_the_device(L, O, handle):
    """VectorOfVectors form (flat data + element pointers) built on the device == the host packing of the padded lists"""
    import torch
    n_ev, cap = 3000, 32
    wf = sipm_population(48, seed=5)
    wf = np.tile(wf, (n_ev // 48 + 1, 1))[:n_ev]
    cfg = L.example_sipm_config()
    cfg["filters"]["sg"].update(min_threshold=-3.0, max_threshold=3.0, min_dc_threshold=-40.0, max_dc_threshold=40.0)
    cfg["filters"]["trap"].update(min_threshold=-15.0, max_threshold=15.0, min_dc_threshold=-30.0, max_dc_threshold=30.0)
    P = L.resolve_sipm_params(cfg, {"sg": {"wl": L.ns(200.0)}}, n_samples=6250, max_triggers=cap)
    d_wf = torch.from_numpy(wf.view(np.int16)).cuda()
    d_rows = torch.zeros((n_ev, L._abi.SIPM_NCOL), dtype=torch.float64, device="cuda")
    d_trig = torch.zeros((n_ev, 4, 4, cap), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()   # the handle launches on its own non-blocking stream: torch's fills / copies must be done first
    handle.sipm_run_device(P, d_wf.data_ptr(), n_ev, 6250, d_rows.data_ptr(), d_trig.data_ptr())
    handle.synchronize()
    rows, trig = d_rows.cpu().numpy(), d_trig.cpu().numpy()
    tbl = L.sipm_to_table(rows, trig)
    for lst, name in ((0, "trig_pos"), (2, "trig_pos_trap")):
        d_ptr = torch.zeros(n_ev + 1, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        total = handle.sipm_list_pointers_device(d_rows.data_ptr(), n_ev, lst, cap, d_ptr.data_ptr())
        ref = tbl[name]
        assert total == len(ref.data) and np.array_equal(d_ptr.cpu().numpy(), ref.elem_ptr)
        d_flat = torch.zeros((4, max(total, 1)), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        handle.sipm_list_gather_device(d_trig.data_ptr(), n_ev, lst, cap, d_ptr.data_ptr(), d_flat.data_ptr(), max(total, 1))
        handle.synchronize()
        flat = d_flat.cpu().numpy()
        assert np.array_equal(flat[0, :total], ref.data)
        mx_name = "trig_max" if lst == 0 else "trig_max_trap"
        assert np.array_equal(flat[3, :total], tbl[mx_name].data)
    assert total > n_ev // 2


def test_sipm_rows_do_not_depend_on_batch_order(L, O, handle):
    """persistent CTAs process several events each: the result of an event must not depend on what the CTA did before
    (shuffled batch == shuffled rows, bit for bit), for dsp_sipm and for the two passes of dsp_icpc_compressed"""
    rng = np.random.default_rng(17)
    wf = sipm_population(96, seed=23)
    wf = np.tile(wf, (13, 1))                       # 1248 events > resident CTAs
    perm = rng.permutation(len(wf))
    cfg = L.example_sipm_config()
    cfg["filters"]["sg"].update(min_threshold=-3.0, max_threshold=3.0, min_dc_threshold=-40.0, max_dc_threshold=40.0)
    cfg["filters"]["trap"].update(min_threshold=-15.0, max_threshold=15.0, min_dc_threshold=-30.0, max_dc_threshold=30.0)
    P = L.resolve_sipm_params(cfg, {"sg": {"wl": L.ns(200.0)}}, n_samples=6250, max_triggers=48)
    r1, t1 = L.sipm_rows(wf, P, handle=handle)
    r2, t2 = L.sipm_rows(np.ascontiguousarray(wf[perm]), P, handle=handle)
    assert np.array_equal(r1[perm].view(np.int64), r2.view(np.int64)) and np.array_equal(t1[perm].view(np.int64), t2.view(np.int64))
    # identical input traces (the tiling) give identical rows wherever they sit
    assert np.array_equal(r1[:96].view(np.int64), r1[96:192].view(np.int64))
    # compressed dsp_icpc: same property through the host entry
    full = L.synth.generate_host(700, first_event=31000)
    pre, wdw = L.synth.compress(full, 8, (2600, 1400))
    cfgc = L.tiefree_config()
    Pp, Pw, aux = L.resolve_compressed_params(cfgc, L.us(500.0), None, presum_rate=8, n_pre=1024, step_pre=L.ns(128.0), n_wdw=1400,
                                              t_first_wdw=L.ns(16.0 * 2600), step_wdw=L.ns(16.0))
    perm = rng.permutation(len(full))

    def run(p_, w_):
        n = len(p_)
        a, b, s = np.zeros((n, L.NCOL)), np.zeros((n, L.NCOL)), np.zeros((n, 5, 5))
        handle.icpc_compressed_run_host(Pp, Pw, p_.ctypes.data, 4, 1024, w_.ctypes.data, 2, 1400, 8.0, aux, n, a.ctypes.data,
                                        b.ctypes.data, s.ctypes.data)
        return a, b, s
    a1, b1, s1 = run(pre, wdw)
    a2, b2, s2 = run(np.ascontiguousarray(pre[perm]), np.ascontiguousarray(wdw[perm]))
    for x, y in ((a1, a2), (b1, b2), (s1, s2)):
        assert np.array_equal(x[perm].view(np.int64), y.view(np.int64))
