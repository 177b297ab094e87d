"""Writes tests/golden/kat_reference_tests.json: the known-answer vectors that the reference's own unit tests
hold for the hot-path primitives, transcribed as data (the reference is Julia and cannot run here, so the
inputs are regenerated from the formulas in the cited test files and the expected values are copied from the
@test lines).  Run:  python tests/golden/make_kat.py
"""
import json
import math
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def sind(d):
    """Julia's sind: exact at multiples of 30/45/90 degrees"""
    d = d % 360
    sign = 1.0
    if d >= 180:
        d -= 180
        sign = -1.0
    if d > 90:
        d = 180 - d
    exact = {0: 0.0, 30: 0.5, 45: math.sqrt(0.5), 90: 1.0}
    if d in exact:
        return sign * exact[d]
    return sign * (math.sin(math.radians(d)) if d <= 45 else math.cos(math.radians(90 - d)))


kat = {
    "_source": "legend-exp/LegendDSP.jl v0.3.0 test suite, transcribed; file:line per entry",
    "extremestats": {
        "ref": "test/test_stats.jl:12-30 (waveform (0:1:360) ns, sind.(0:360))",
        "signal": [sind(k) for k in range(361)],
        "t0": 0.0, "dt": 1.0,
        "cases": [
            {"from": 0, "until": 360, "min": -1.0, "max": 1.0, "tmin": 270.0, "tmax": 90.0},
            {"from": 0, "until": 180, "min": 0.0, "max": 1.0, "tmin": 0.0, "tmax": 90.0},
            {"from": 135, "until": 225, "min": -math.sqrt(0.5), "max": math.sqrt(0.5), "tmin": 225.0, "tmax": 135.0},
        ],
    },
    "get_wvf_maximum": {
        "ref": "test/test_interpolation.jl:6-45 (dt = 16 ns, 100 samples; 0-based index windows)",
        "cases": [
            {"name": "max at window start", "n": 100, "set": {"0": 1.0, "1": 0.8, "2": 0.5, "3": 0.2},
             "from": 0, "until": 4, "ge": 1.0, "lt": 1.1, "exact": 1.0},
            {"name": "max at window end", "n": 100, "set": {"96": 0.2, "97": 0.5, "98": 0.8, "99": 1.0},
             "from": 95, "until": 99, "ge": 1.0, "lt": 1.1, "exact": 1.0},
            {"name": "max in the middle (parabola)", "n": 100, "set": {"49": 0.5, "50": 1.0, "51": 0.5},
             "from": 47, "until": 53, "ge": 1.0, "lt": 1.2, "exact": 1.0},
        ],
    },
    "derivative": {
        "ref": "test/test_derivative.jl:11-18: y == gain * vcat(x[2]-x[1], diff(x))",
    },
    "intersect_ramp": {
        "ref": "test/test_multiintersect.jl:8-26: wvf (1:100) s, values 1:100; Intersect(mintot=1s)(wvf, 50).x == 50 s; "
               "ratios 0.1..0.9 of the maximum 100 -> 10..90 s",
        "t0": 1.0, "dt": 1.0, "signal": list(range(1, 101)),
        "cases": [{"thr": 10.0 * k, "min_n": 1, "x": 10.0 * k} for k in range(1, 10)],
    },
    "intersect_state_machine": {
        "ref": "test/test_intersect_maximum.jl:6-107 (dt = 16 ns, 6200 samples, threshold 0.4, mintot = 2 samples): "
               "multiplicities and crossing brackets of the shared up-crossing state machine "
               "(src/intersect_maximum.jl:41-56)",
        "n": 6200, "dt": 16.0, "thr": 0.4, "min_n": 2,
        "cases": [
            {"name": "near start", "set": {"1": 0.5, "2": 0.6, "3": 0.2}, "multiplicity": 1, "x_gt": 0.0, "x_lt": 48.0},
            {"name": "near end", "set": {"6197": 0.5, "6198": 0.6, "6199": 0.2}, "multiplicity": 1,
             "x_gt": 6196 * 16.0, "x_lt": 6199 * 16.0},
            {"name": "rising to the last sample", "set": {"6195": 0.3, "6196": 0.5, "6197": 0.6, "6198": 0.8, "6199": 1.0},
             "multiplicity": 1},
            {"name": "crossing at the last two samples", "set": {"6197": 0.3, "6198": 0.5, "6199": 0.6},
             "multiplicity": 1, "x_gt": 6196 * 16.0},
            {"name": "two pulses", "ranges": [[99, 104, 0.8], [199, 214, 0.9]], "multiplicity": 2},
        ],
    },
    "intersect_maximum": {
        "ref": "test/test_intersect_maximum.jl:6-107: IntersectMaximum(mintot = 2 dt, maxtot = 100 dt) (case 'max at last "
               "sample': maxtot = 5 dt) on 6200 samples, dt = 16 ns, threshold 0.4; expected values/brackets from the @test lines",
        "n": 6200, "dt": 16.0, "thr": 0.4, "min_n": 2,
        "cases": [
            {"name": "near start (:13-32)", "max_n": 100, "set": {"0": 0.0, "1": 0.5, "2": 0.6, "3": 0.2},
             "multiplicity": 1, "x_gt": 0.0, "x_lt": 48.0, "max_ge": 0.6, "max_lt": 0.7, "x_high_gt_x": True, "tot_gt": 0.0},
            {"name": "near end (:35-52)", "max_n": 100, "set": {"6196": 0.0, "6197": 0.5, "6198": 0.6, "6199": 0.2},
             "multiplicity": 1, "x_gt": 6196 * 16.0, "x_lt": 6199 * 16.0, "max_ge": 0.6, "max_lt": 0.7, "x_high_gt_x": True,
             "tot_gt": 0.0},
            {"name": "max at last sample of the window (:56-68)", "max_n": 5,
             "set": {"6195": 0.3, "6196": 0.5, "6197": 0.6, "6198": 0.8, "6199": 1.0}, "multiplicity": 1, "max_eq": 1.0},
            {"name": "crossing at the last samples (:71-82)", "max_n": 100, "set": {"6197": 0.3, "6198": 0.5, "6199": 0.6},
             "multiplicity": 1, "x_gt": 6196 * 16.0, "x_high_eq": 6199 * 16.0},
            {"name": "two pulses (:96-106)", "max_n": 100, "ranges": [[99, 104, 0.8], [199, 214, 0.9]], "multiplicity": 2,
             "tot_gt": 0.0, "tot_increasing": True},
        ],
        "empty": {"ref": ":85-93", "multiplicity": 0},
    },
    "thresholdstats_mad": {
        "ref": "test/test_thresholdstats.jl:7-65",
        "cases": [
            {"name": "constant signal (:11-17)", "signal": [5.0] * 100, "min": -10.0, "max": 10.0, "expect": 0.0, "atol": 1e-10},
            {"name": "symmetric signal (:20-27)", "signal": [-1.0] * 50 + [1.0] * 50, "min": -5.0, "max": 5.0, "expect": 1.4826,
             "atol": 1e-10},
            {"name": "outlier robustness (:30-39)", "signal": [0.0] * 499 + [1000.0] * 11 + [0.0] * 490, "min": "-inf",
             "max": "inf", "lt": 1.0},
            {"name": "empty filter (:42-48)", "signal": [5.0] * 100, "min": 10.0, "max": 20.0, "expect": 0.0, "atol": 1e-10},
        ],
    },
    "thresholdstats": {
        "ref": "test/test_stats.jl:57-110: thresholdstats(wf, min, max) == std(wf[min .<= wf .<= max]) within rtol 0.005 "
               "(0.001 without bounds) on sigma * randn(10000); the random input is regenerated with a fixed seed",
        "n": 10000, "seed": 12345, "rtol_bounds": 0.005, "rtol_all": 0.001,
    },
    "thresholds_fixture": {
        "ref": "test/test_dsp_icpc.jl:189-199 properties on make_fake_waveform: t0 < t50 < t90, drift_time >= 0, "
               "e_10410/e_313/e_trap finite",
    },
}

with open(os.path.join(HERE, "kat_reference_tests.json"), "w") as f:
    json.dump(kat, f, indent=1)
print("written", os.path.join(HERE, "kat_reference_tests.json"))
