"""Randomised-configuration parity of the fused dsp_icpc kernel: the reference's example config
(test/test_dsp_icpc.jl:50-161) is only one point of the parameter space a LEGEND channel config spans.  Every seed draws
a different DSPConfig (windows, filter lengths, thresholds, interpolation orders), decay constant, optimised filter
parameters (`pars_filter`, src/utils.jl:72-82), sampling step and sample count; the C-ABI result must equal the
float64 oracle on the same seeded waveforms (structured CUSP/ZAC evaluation against the oracle's direct FIRs).
LGDSP_FUZZ_SEEDS=<n> widens the sweep (80 seeds were run green on the B200 in round 1)."""
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
    res, n_ties = assert_parity_with_ties(L, O, P, wf, got, ref)
    c = L.COL
    # the population is not degenerate (high t0 thresholds leave small pulses without a t0)
    assert np.isfinite(ref[:, c["e_trap"]]).sum() > 100 and (ref[:, c["t0"]] > 0).sum() > 40
