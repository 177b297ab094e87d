"""The split dsp_icpc pipeline (prefix / extract / CUSP-ZAC kernels coupled through the prefix-sum ring) against the
single-kernel path of round 1, through the C ABI: both evaluate the same expressions in the same order, so the rows
must be IDENTICAL (bit for bit, NaNs in the same places), for every batch / stream setting and column-group mask.
Exception: the three tailstats columns -- the split path sums the tail logarithms in a different order and with shorter
log1p polynomials, so they are compared with the oracle tolerances of tests/parity.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rows(L, h, wf, P, path, batch=0, streams=0):
    h.set_icpc_path(path, batch, streams)
    try:
        return L.dsp_icpc_rows(wf, P, handle=h)
    finally:
        h.set_icpc_path("split")


_TAIL = ("tail_mean", "tail_sigma", "tail_tau")


def _identical(a, b, columns):
    from parity import compare_rows
    bad = {k: v for k, v in compare_rows(a, b, columns).items() if k in _TAIL and v[1] > 0}
    for j, name in enumerate(columns):
        if name in _TAIL:
            continue
        x, y = a[:, j], b[:, j]
        same = (x == y) | (np.isnan(x) & np.isnan(y))
        if not same.all():
            k = int(np.nonzero(~same)[0][0])
            bad[name] = (int((~same).sum()), k, float(x[k]), float(y[k]))
    return bad


@pytest.mark.parametrize("direct", [False, True])
def test_split_equals_fused_mixed_population(L, O, handle, direct):
    P = L.resolve_icpc_params(L.example_config(), L.us(500.0), cuspzac_direct=direct)
    n = 3000 if not direct else 300
    wf = L.synth.generate_host(n, first_event=5000)
    fused = _rows(L, handle, wf, P, "fused")
    for batch, streams in ((0, 0), (257, 3), (64, 1), (5000, 2)):
        split = _rows(L, handle, wf, P, "split", batch, streams)
        bad = _identical(split, fused, L.COLUMNS)
        assert not bad, f"batch {batch} streams {streams}: {bad}"


@pytest.mark.parametrize("groups", [0x01, 0x07, 0x0F, 0x27, 0x47, 0x1F, 0x6F])
def test_split_equals_fused_groups(L, O, handle, groups):
    P = L.resolve_icpc_params(L.tiefree_config(), L.us(400.0), groups=groups)
    wf = L.synth.generate_host(700, first_event=123456)
    fused = _rows(L, handle, wf, P, "fused")
    split = _rows(L, handle, wf, P, "split", 300, 2)
    bad = _identical(split, fused, L.COLUMNS)
    assert not bad, bad


def test_split_separate_cusp_zac_and_short_traces(L, O, handle):
    """unequal CUSP / ZAC lengths (two structured passes) and a 6144-sample trace"""
    cfg = L.example_config()
    pars = {"cusp": {"rt": L.us(3.0), "ft": L.us(1.0)}, "zac": {"rt": L.us(4.0), "ft": L.us(2.0)}}
    cfg.tail_window = (L.us(60.0), L.us(90.0))
    P = L.resolve_icpc_params(cfg, L.us(300.0), pars, n_samples=6144)
    wf = L.synth.generate_host(500, first_event=777)[:, :6144].copy()
    fused = _rows(L, handle, wf, P, "fused")
    split = _rows(L, handle, wf, P, "split", 128, 2)
    bad = _identical(split, fused, L.COLUMNS)
    assert not bad, bad


def test_lean_pz_trap_columns(L, O, handle):
    """LGDSP_GROUP_PZTRAP_LEAN (BASELINE configs[1]): the five columns it keeps are the full chain's values, bit for bit"""
    cfg = L.tiefree_config()
    wf = L.synth.generate_host(1500, first_event=99)
    full = L.dsp_icpc_rows(wf, L.resolve_icpc_params(cfg, L.us(500.0)), handle=handle)
    lean = L.dsp_icpc_rows(wf, L.resolve_icpc_params(cfg, L.us(500.0), groups=L._abi.GROUP_PZTRAP_LEAN), handle=handle)
    for name in ("blmean", "t0", "t50", "e_trap", "e_10410", "e_max", "e_min", "n_sat_high"):
        j = L.COL[name]
        assert np.array_equal(lean[:, j], full[:, j]), name
