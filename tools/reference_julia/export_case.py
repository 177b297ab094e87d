"""Export one parity case for the REAL reference (run outside this image, where Julia is available).

    python tools/reference_julia/export_case.py out_dir [n_events]

writes
    out_dir/wf_u16.bin      n_events x 8192 UInt16 waveforms of the seeded synthetic stream (row = event)
    out_dir/case.json       sizes, tau, sampling step and the DSP config as {"val", "unit"} quantities
                            (the values of the reference's example config, test/test_dsp_icpc.jl:50-161)
    out_dir/oracle_rows.bin n_events x 49 float64: what oracle/ (the CPU restatement) gives, columns of _abi.COLUMNS

`julia tools/reference_julia/dump_reference.jl out_dir` then writes out_dir/reference_rows.bin, and
`python tools/reference_julia/compare_dump.py out_dir` compares the two with the tolerances of tests/parity.py.
The waveform generator needs no GPU (host Philox stream of the product library)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def _plain(v):
    """Q -> {"val", "unit"} (the unit is kept: Unitful's mixed-unit arithmetic decides rounding ties); containers recursively"""
    if hasattr(v, "val") and hasattr(v, "unit"):
        return {"val": v.val, "unit": "us" if v.unit in ("us", "µs", "μs") else v.unit}
    if isinstance(v, dict):
        return {k: _plain(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_plain(x) for x in v]
    return v


def main():
    import importlib
    L = importlib.import_module("legenddsp.jl_b200")
    cfgm = importlib.import_module("legenddsp.jl_b200.config")
    from oracle import oracle as O
    out = sys.argv[1]
    n_events = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    os.makedirs(out, exist_ok=True)
    wf = L.synth.generate_host(n_events, first_event=0)
    wf.tofile(os.path.join(out, "wf_u16.bin"))
    d = cfgm.example_config_dict()
    P = L.resolve_icpc_params(cfgm.DSPConfig.from_dict(d), L.us(500.0), builders=O.OracleBuilders())
    rows, _ = O.dsp_icpc(P, wf)
    np.ascontiguousarray(rows, dtype=np.float64).tofile(os.path.join(out, "oracle_rows.bin"))
    case = {"n_events": n_events, "n_samples": int(wf.shape[1]), "step_ns": 16.0, "tau_ns": 500000.0,
            "columns": list(L.COLUMNS), "units": dict(L._abi.UNITS),             "config": _plain(d)}
    with open(os.path.join(out, "case.json"), "w") as f:
        json.dump(case, f, indent=1)
    print(f"wrote {n_events} events to {out}")


if __name__ == "__main__":
    main()
