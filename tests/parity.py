"""Column-wise parity rules between the CUDA path and the CPU oracle (same inputs).

north_star tolerances: bit-exact for integer/index outputs, 1e-5 relative on energies, +-1 sample (16 ns) on
interpolated times.  The assertions below are much tighter because both sides compute in float64; what differs
is only the summation order (closed-form prefix sums vs the reference's recursions), i.e. ~1e-10 absolute.
"""
import numpy as np

EXACT = ("qc_label", "inTrace_n", "n_sat_low", "n_sat_high", "n_sat_low_cons", "n_sat_high_cons",
         "e_max", "e_min", "blmean", "t_trap_max", "t_cusp_max", "t_zac_max")
# (rtol, atol) ; times in us unless noted
TOL = {
    "blsigma": (1e-9, 1e-9), "blslope": (1e-9, 1e-15), "bloffset": (1e-12, 1e-9),
    "tailmean": (1e-10, 1e-8), "tailslope": (1e-6, 1e-11), "tailoffset": (1e-9, 1e-6),
    # sigma = sqrt(E[y^2]-E[y]^2) cancels ~1e8 against ~10: absolute error ~ sqrt(eps)*|mean| in BOTH implementations
    "tailsigma": (1e-6, 2e-4),
    "t0": (0, 1e-7), "t10": (0, 1e-7), "t50": (0, 1e-7), "t80": (0, 1e-7), "t90": (0, 1e-7), "t99": (0, 1e-7),
    "t50_current": (0, 1e-7), "t0_inv": (0, 1e-7),
    "drift_time": (0, 1e-4),            # ns
    # tail_tau = -1/slope is compared through its reciprocal (slope in 1/ns): clipped or empty events have a constant
    # tail, slope = 0 up to rounding, and tau is +-huge in the reference as well
    "tail_tau": (1e-7, 1e-13), "tail_mean": (1e-12, 1e-12),
    # sqrt(E[l^2]-E[l]^2) with l = log(y) ~ 9: cancellation error ~1e-6*|mean| (same in the reference)
    "tail_sigma": (1e-7, 1e-5),
    "e_10410": (1e-9, 1e-7), "e_535": (1e-9, 1e-7), "e_313": (1e-9, 1e-7),
    "e_10410_inv": (1e-9, 1e-7), "e_313_inv": (1e-9, 1e-7),
    "e_trap": (1e-9, 1e-7), "e_cusp": (1e-8, 1e-6), "e_zac": (1e-8, 1e-6),
    "e_trap_max": (1e-9, 1e-7), "e_cusp_max": (1e-8, 1e-6), "e_zac_max": (1e-8, 1e-6),
    "qdrift": (1e-9, 1e-3), "lq": (1e-9, 1e-3),   # second differences of integrals ~1e8
    "a_sg": (1e-9, 1e-7), "a_60": (1e-9, 1e-7), "a_100": (1e-9, 1e-7), "a_raw": (1e-9, 1e-7),
    "inTrace_intersect": (0, 1e-4),     # ns, NaN == NaN
}


def compare_rows(got, ref, columns):
    """returns {column: (max_abs_err, n_bad)} using the rules above; NaNs must coincide"""
    res = {}
    for j, name in enumerate(columns):
        a, b = got[:, j], ref[:, j]
        nan_a, nan_b = np.isnan(a), np.isnan(b)
        bad_nan = int((nan_a != nan_b).sum())
        ok = ~(nan_a | nan_b)
        if name == "tail_tau":
            with np.errstate(divide="ignore"):
                a = np.where(a == 0, 0.0, 1.0 / np.where(a == 0, 1.0, a))
                b = np.where(b == 0, 0.0, 1.0 / np.where(b == 0, 1.0, b))
        d = np.abs(a[ok] - b[ok])
        if name in EXACT:
            bad = int((a[ok] != b[ok]).sum())
        else:
            rtol, atol = TOL[name]
            bad = int((d > atol + rtol * np.abs(b[ok])).sum())
        res[name] = (float(d.max()) if d.size else 0.0, bad + bad_nan)
    return res


def assert_parity(got, ref, columns, allow=0, skip=()):
    res = compare_rows(got, ref, columns)
    bad = {k: v for k, v in res.items() if v[1] > allow and k not in skip}
    assert not bad, f"parity violations (column: (max abs err, count)): {bad}"
    return res
