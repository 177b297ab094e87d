"""Tiny invocations of every kernel for compute-sanitizer (memcheck / racecheck) runs on the GPU box:
  compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import importlib
import numpy as np
L = importlib.import_module("legenddsp.jl_b200")
h = L.Handle(0)
which = sys.argv[1:] or ["icpc", "compressed", "sipm", "mi", "sweep"]
if "icpc" in which:
    P = L.resolve_icpc_params(L.tiefree_config(), L.us(500.0))
    wf = L.synth.generate_host(6, first_event=0)
    rows = L.dsp_icpc_rows(wf, P, handle=h)
    print("icpc", np.isfinite(rows).mean())
if "compressed" in which:
    from test_gpu_compressed import _data
    d = _data(L, 6, 7000, 8)
    r = L.dsp_icpc_compressed(d, L.tiefree_config(), L.us(500.0), None, handle=h)
    print("compressed", float(np.nanmean(r["e_trap"])))
if "sipm" in which:
    from test_gpu_sipm import sipm_population
    cfg = L.example_sipm_config()
    cfg["filters"]["sg"].update(min_threshold=-3.0, max_threshold=3.0, min_dc_threshold=-40.0, max_dc_threshold=40.0)
    cfg["filters"]["trap"].update(min_threshold=-15.0, max_threshold=15.0, min_dc_threshold=-30.0, max_dc_threshold=30.0)
    Ps = L.resolve_sipm_params(cfg, {"sg": {"wl": L.ns(200.0)}}, n_samples=6250, max_triggers=32)
    rows, trig = L.sipm_rows(sipm_population(4), Ps, handle=h)
    print("sipm", rows[:, L._abi.SIPM_COL["n_trig"]])
    print("mad", L.thresholdstats_mad(np.random.default_rng(1).normal(0, 1, 3000), -1.0, 1.0, handle=h))
    print("im", L.IntersectMaximum(L.ns(32.0), L.ns(160.0))(np.random.default_rng(2).normal(0, 1, 3000), 1.5, handle=h)["multiplicity"])
if "mi" in which:
    y = np.clip((np.arange(2000) - 500) / 50.0, 0, 1) * 100 + np.random.default_rng(3).normal(0, 0.2, (5, 2000))
    print("mi", L.MultiIntersect(n=2, d=2, sampling_rate=4)(y, handle=h)[0, :3])
if "sweep" in which:
    wf = L.synth.generate_host(4, first_event=3)
    g = L.dsp_trap_ft_optimization(L.RDWaveforms(wf), L.tiefree_config(), L.us(500.0), L.us(5.0), handle=h)
    print("sweep", g.shape)
    # the 20x10 grid on the one-warp-per-waveform kernel: generator events plus steps at the trace ends (second window)
    ends = np.full((3, 8192), 10000, np.uint16)
    ends[0, 40:] = 50000; ends[1, 8192 - 20:] = 50000; ends[2, :] = 0
    wf2 = np.concatenate([L.synth.generate_host(40, first_event=11), L.synth.generate_host(8, mode=1), ends])
    rts = [L.us(1.0 + 0.75 * i) for i in range(20)]
    fts = [L.us(1.0 + 0.3 * i) for i in range(10)]
    os.environ["LGDSP_SWEEP_PATH"] = "warp"
    g2 = L.dsp_trap_rtft_grid(L.RDWaveforms(wf2), L.tiefree_config(), L.us(500.0), rts, fts, handle=h)
    os.environ.pop("LGDSP_SWEEP_PATH")
    print("sweep warp", g2.shape, float(np.nanmean(g2)))
h.close()
