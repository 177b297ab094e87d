"""Host-side resolution of the reference's example config into sample units."""
import pytest


def test_appendix_a_constants(L, O):
    P = L.resolve_icpc_params(L.example_config(), L.us(500.0), builders=O.OracleBuilders())
    s = L.params_summary(P)
    assert s["n_samples"] == 8192 and s["dt_ns"] == 16.0
    assert s["sat"] == (0, 65520)                       # 2^16 - 16 (sic), src/dsp_icpc.jl:94
    assert s["bl"] == (0, 2438)                         # 1-based 1..2439 (39 us / 16 ns = 2437.5 -> even)
    assert s["tail"] == (4375, 6875)                    # 1-based 4376..6876
    assert s["t0_trap"] == (2, 6, 125) and s["t0_min_n"] == 94
    assert s["tx_min_n"] == 2 and s["intrace_min_n"] == 6
    assert s["trap_10410"] == (625, 250, 625)
    assert s["trap_535"] == (312, 188, 312)             # 312.5 -> 312, 187.5 -> 188 (ties to even)
    assert s["trap_313"] == (188, 62, 188)              # 62.5 -> 62
    assert s["trap_e"] == (312, 156, 312)
    assert s["sig_dni"] == (3, 44) and s["int_dni"] == (3, 6)
    assert s["cusp"][:2] == (2375, 156) and s["cusp"][2] == 312.5
    assert abs(s["RC"] - 31250.0) < 1e-6
    assert P.trap_pickoff_ns == 6250.0 and P.cusp_pickoff_ns == 19000.0
    assert P.qdrift_first_ns == 2500.0 and P.qdrift_last_ns == 5000.0


def test_julia_rounding_emulation(L):
    from importlib import import_module
    cfgm = import_module("legenddsp.jl_b200.config")
    assert cfgm.julia_round(2.5) == 2 and cfgm.julia_round(3.5) == 4 and cfgm.julia_round(312.5) == 312
    assert cfgm._ratio(L.us(5.0), L.ns(16.0)) == 312.5          # (5.0/16.0) * 1000, exact
    assert cfgm._ratio(L.ns(1500.0), L.ns(16.0)) == 93.75


def test_get_fltpars_override(L):
    cfg = L.example_config()
    assert L.get_fltpars(None, "trap", cfg) == (L.us(5.0), L.us(2.5))
    assert L.get_fltpars({"trap": {"rt": L.us(8.0)}}, "trap", cfg) == (L.us(8.0), L.us(2.5))   # src/utils.jl:80
    assert L.get_fltpars({"sg": {"wl": L.ns(180.0)}}, "sg", cfg) == L.ns(180.0)
    P = L.resolve_icpc_params(cfg, L.us(500.0), {"trap": {"rt": L.us(8.0), "ft": L.us(3.0)}})
    assert P.trap_e.as_tuple() == (500, 188, 500) and P.trap_pickoff_ns == 9500.0


def test_window_assertions(L):
    cfg = L.example_config()
    cfg.tail_window = (L.us(70.0), L.us(140.0))          # beyond the 131 us trace -> @assert, src/tailstats.jl:23-25
    with pytest.raises(AssertionError):
        L.resolve_icpc_params(cfg, L.us(500.0))


def test_grid_values(L):
    cfg = L.example_config()
    assert len(L.grid_values(cfg.e_grid_rt_trap)) == 31 and len(L.grid_values(cfg.e_grid_ft_trap)) == 16
    assert len(L.grid_values(cfg.a_grid_wl_sg)) == 11
