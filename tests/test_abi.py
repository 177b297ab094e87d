"""The C ABI: the ctypes mirror matches include/lgdsp_b200.h byte for byte, the library loads without a GPU
and exports every declared symbol, and the compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "lgdsp_b200.h")


def _c_layout():
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "lgdsp_b200.h"
#define S(t) printf("sizeof " #t " %zu\n", sizeof(t))
#define O(t, f) printf("offsetof " #t "." #f " %zu\n", offsetof(t, f))
int main(void) {
  S(lgdsp_trap); S(lgdsp_dni); S(lgdsp_sg); S(lgdsp_cuspzac); S(lgdsp_icpc_params); S(lgdsp_trap_variant);
  S(lgdsp_sweep_params); S(lgdsp_synth_params); S(lgdsp_sweep_variant);
  O(lgdsp_icpc_params, groups); O(lgdsp_icpc_params, sat_high); O(lgdsp_icpc_params, pz_km1);
  O(lgdsp_icpc_params, t0inv_trap); O(lgdsp_icpc_params, t0_threshold); O(lgdsp_icpc_params, tx_frac);
  O(lgdsp_icpc_params, int_dni); O(lgdsp_icpc_params, sig_dni); O(lgdsp_icpc_params, trap_e);
  O(lgdsp_icpc_params, zac_pickoff_ns); O(lgdsp_icpc_params, sg); O(lgdsp_icpc_params, cur_from);
  O(lgdsp_icpc_params, intrace_nsigma); O(lgdsp_icpc_params, intrace_bl_until); O(lgdsp_icpc_params, cuspzac_direct);
  O(lgdsp_icpc_params, cusp); O(lgdsp_icpc_params, zac);
  O(lgdsp_cuspzac, coeffs); O(lgdsp_sweep_params, sig_dni); O(lgdsp_trap_variant, pickoff_mode);
  O(lgdsp_sweep_params, out_f64); O(lgdsp_sweep_variant, trap); O(lgdsp_sweep_variant, n_taps);
  O(lgdsp_sweep_variant, win_until); O(lgdsp_sweep_variant, coeffs);
  printf("ncol %d\n", (int)LGDSP_NCOL);
  printf("version %u\n", (unsigned)LGDSP_PARAMS_VERSION);
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "l.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "l")
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.run([cc, "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    return dict((" ".join(l.split()[:-1]), int(l.split()[-1])) for l in out.strip().splitlines())


def test_ctypes_mirror_matches_header(L):
    c = _c_layout()
    A = L._abi
    pairs = {"lgdsp_trap": A.Trap, "lgdsp_dni": A.Dni, "lgdsp_sg": A.Sg, "lgdsp_cuspzac": A.CuspZac,
             "lgdsp_icpc_params": A.IcpcParams, "lgdsp_trap_variant": A.TrapVariant,
             "lgdsp_sweep_params": A.SweepParams, "lgdsp_synth_params": A.SynthParams,
             "lgdsp_sweep_variant": A.SweepVariant}
    for name, cls in pairs.items():
        assert C.sizeof(cls) == c["sizeof " + name], name
    for key, off in c.items():
        if key.startswith("offsetof"):
            t, f = key.split()[1].split(".")
            assert getattr(pairs[t], f).offset == off, key
    assert c["ncol"] == A.NCOL == 49
    assert c["version"] == A.LGDSP_PARAMS_VERSION


def test_column_enum_order_matches_header(L):
    names = re.findall(r"LGDSP_COL_(\w+)", open(HDR).read().split("enum lgdsp_col")[1].split("LGDSP_NCOL")[0])
    assert tuple(names) == L.COLUMNS


def test_library_loads_and_exports_every_declared_symbol(L):
    lib = L.load_library()
    declared = set(re.findall(r"\b(lgdsp_[a-z_0-9]+)\s*\(", open(HDR).read()))
    declared -= {"lgdsp_b200"}
    assert declared == set(L.EXPORTED_SYMBOLS), declared ^ set(L.EXPORTED_SYMBOLS)
    for s in declared:
        assert getattr(lib, s) is not None
    assert b"sm_100a" in lib.lgdsp_version()


def test_no_cpu_fallback(L):
    """without a CUDA device the product path must fail loudly"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(L.LgdspError) as ei:
        L.Handle(0)
    assert ei.value.code == L._abi.LGDSP_ERR_NO_DEVICE
    import numpy as np
    with pytest.raises(L.LgdspError):
        L.dsp_icpc({"waveform": np.zeros((2, 8192), dtype=np.uint16)}, L.example_config(), L.us(500.0))


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under legenddsp.jl_b200/ may reference it"""
    pkg = os.path.join(ROOT, "legenddsp.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".jl")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liblgdsp_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_julia_wrapper_stays_in_sync_with_the_header(L):
    """julia/LegendDSPB200.jl cannot be run here (no Julia toolchain).  What can be checked: (1) julia/lgdsp_offsets.jl -- the
    struct layouts the wrapper writes its parameter blocks at -- is exactly what the C compiler measures on the header today
    and equals the ctypes mirror; (2) every offset / constant the wrapper uses exists in that file; (3) every C symbol it
    ccalls is exported; (4) its column list is the header's column enum."""
    import ctypes as C
    import importlib.util
    import re
    spec = importlib.util.spec_from_file_location("gen_julia_offsets", os.path.join(ROOT, "tools", "gen_julia_offsets.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    committed = open(os.path.join(ROOT, "julia", "lgdsp_offsets.jl")).read()
    assert committed == gen.generate(), "julia/lgdsp_offsets.jl is stale: run python tools/gen_julia_offsets.py"
    abi = L._abi
    mirror = {"LGDSP_TRAP": abi.Trap, "LGDSP_DNI": abi.Dni, "LGDSP_SG": abi.Sg, "LGDSP_CUSPZAC": abi.CuspZac,
              "LGDSP_ICPC_PARAMS": abi.IcpcParams, "LGDSP_SWEEP_PARAMS": abi.SweepParams, "LGDSP_TRAP_VARIANT": abi.TrapVariant,
              "LGDSP_SWEEP_VARIANT": abi.SweepVariant}
    offs = {}
    for name, body in re.findall(r"const OFF_(\w+) = \((.*?)\)\n", committed):
        offs[name] = dict((k, int(v)) for k, v in re.findall(r"(\w+) = (\d+)", body))
        for f, o in offs[name].items():
            assert getattr(mirror[name], f).offset == o, (name, f)
    for name, size in re.findall(r"const SIZEOF_(\w+) = (\d+)", committed):
        assert C.sizeof(mirror[name]) == int(size), name
    consts = dict((k, int(v)) for k, v in re.findall(r"const (LGDSP_\w+) = (\d+)", committed))
    assert consts["LGDSP_NCOL"] == abi.NCOL and consts["LGDSP_PARAMS_VERSION"] == abi.LGDSP_PARAMS_VERSION
    assert consts["LGDSP_GROUP_ALL"] == abi.GROUP_ALL and consts["LGDSP_MAX_FIR"] == abi.LGDSP_MAX_FIR

    src = open(os.path.join(ROOT, "julia", "LegendDSPB200.jl"), encoding="utf-8").read()
    for name, field in re.findall(r"\bOFF_(LGDSP_\w+)\.(\w+)", src):
        assert field in offs[name], f"OFF_{name}.{field} is not a field of the header's struct"
    aliases = set(v for _, v in re.findall(r"\b(\w+) = OFF_(LGDSP_\w+)\b", src))        # e.g. `O = OFF_LGDSP_ICPC_PARAMS`
    assert aliases, "the wrapper is expected to alias the offset tables"
    aliased_fields = set().union(*(offs[v] for v in aliases))
    for field in set(re.findall(r"\bO\.(\w+)", src)):
        assert field in aliased_fields, f"O.{field} is not a field of any aliased struct"
    for ident in set(re.findall(r"\b(LGDSP_[A-Z0-9_]+|SIZEOF_LGDSP_[A-Z_]+)\b", src)):
        assert ident in consts or ident in ("SIZEOF_" + k for k in mirror) or ident.startswith("LGDSP_B200"), ident
    lib = L.load_library()
    for sym in set(re.findall(r"ccall\(\(:(\w+), LIB\)", src)) | set(re.findall(r":(lgdsp_(?:cusp|zac)_coeffs)", src)):
        assert hasattr(lib, sym), f"{sym} is not exported by liblgdsp_b200.so"
    # the column list of the wrapper's output table is the header's column enum
    cols = re.search(r"const COLS = \((.*?)\)\n", src, re.S).group(1)
    jl_cols = [c.replace("tail_τ", "tail_tau") for c in re.findall(r":([\wτ]+)", cols)]
    assert jl_cols == list(abi.COLUMNS)
