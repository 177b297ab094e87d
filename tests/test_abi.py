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


def test_julia_wrapper_struct_layout_matches_the_ctypes_mirror(L):
    """julia/LegendDSPB200.jl cannot be run here (no Julia toolchain): at least its `struct` declarations must list the
    fields of include/lgdsp_b200.h in the same order with the same element types / array lengths as the ctypes mirror
    (which test_ctypes_mirror_matches_header ties to the header), and its constants must equal the header's."""
    import ctypes as C
    import re
    src = open(os.path.join(ROOT, "julia", "LegendDSPB200.jl"), encoding="utf-8").read()
    abi = L._abi
    consts = dict(re.findall(r"const (MAX_DNI|MAX_SG|MAX_FIR|NCOL)\b", src) and
                  zip(("MAX_DNI", "MAX_SG", "MAX_FIR", "NCOL"),
                      map(int, re.search(r"const MAX_DNI, MAX_SG, MAX_FIR, NCOL = (\d+), (\d+), (\d+), (\d+)", src).groups())))
    assert consts == {"MAX_DNI": abi.LGDSP_MAX_DNI, "MAX_SG": abi.LGDSP_MAX_SG, "MAX_FIR": abi.LGDSP_MAX_FIR, "NCOL": abi.NCOL}
    assert int(re.search(r"PARAMS_VERSION = UInt32\((\d+)\)", src).group(1)) == abi.LGDSP_PARAMS_VERSION

    jl_scalar = {"Int32": C.c_int32, "UInt32": C.c_uint32, "Int64": C.c_int64, "Float64": C.c_double}
    env = {"MAX_DNI": abi.LGDSP_MAX_DNI, "MAX_SG": abi.LGDSP_MAX_SG, "MAX_FIR": abi.LGDSP_MAX_FIR}
    structs = {"Trap": abi.Trap, "Dni": abi.Dni, "Sg": abi.Sg, "CuspZac": abi.CuspZac, "IcpcParams": abi.IcpcParams}

    def jl_fields(name):
        body = re.search(r"struct %s\b(.*?)\bend" % name, src, re.S).group(1)
        body = re.sub(r"#.*", "", body)
        return re.findall(r"(\w+)::([\w{}*, ]+?)(?:;|\n|$)", body)

    def ctype_of(jl_type):
        jl_type = jl_type.strip()
        m = re.fullmatch(r"NTuple\{(.+),\s*(\w+)\}", jl_type)
        if m:
            length = eval(m.group(1), {}, env)
            return ctype_of(m.group(2)) * length
        return jl_scalar.get(jl_type) or structs[jl_type]

    for name, cstruct in structs.items():
        jf = jl_fields(name)
        cf = list(cstruct._fields_)
        assert [f for f, _ in jf] == [f for f, _ in cf] or [f for f, _ in jf] == [f.replace("n_window", "n_w") for f, _ in cf], \
            (name, [f for f, _ in jf], [f for f, _ in cf])
        for (jn, jt), (cn, ct) in zip(jf, cf):
            assert C.sizeof(ctype_of(jt)) == C.sizeof(ct), (name, jn, jt, ct)
    # the column list of the wrapper's output table is the header's column enum
    cols = re.search(r"const COLS = \((.*?)\)\n", src, re.S).group(1)
    jl_cols = [c.replace("tail_τ", "tail_tau") for c in re.findall(r":([\wτ]+)", cols)]
    assert jl_cols == list(abi.COLUMNS)
