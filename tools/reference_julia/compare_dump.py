"""Compare the reference's rows (dump_reference.jl) with the oracle rows exported beside them.

    python tools/reference_julia/compare_dump.py case_dir

Same tolerances as the GPU parity tests (tests/parity.py).  A mismatch that follows one RddspPolicy switch
(legenddsp.jl_b200/config.py: SG even-length rule, SG / trapezoid time axes, CUSP/ZAC normalisation, DNI window placement)
shows up as a whole column off by a constant: flip the switch, re-export, compare again."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from parity import compare_rows
    d = sys.argv[1]
    case = json.load(open(os.path.join(d, "case.json")))
    n, cols = case["n_events"], case["columns"]
    ref = np.fromfile(os.path.join(d, "reference_rows.bin"), dtype=np.float64).reshape(n, len(cols))
    orc = np.fromfile(os.path.join(d, "oracle_rows.bin"), dtype=np.float64).reshape(n, len(cols))
    res = compare_rows(orc, ref, tuple(cols))
    bad = 0
    for name in cols:
        err, cnt = res[name]
        flag = "" if cnt == 0 else "   <-- MISMATCH"
        bad += cnt > 0
        print(f"{name:20s} max |err| {err:12.4e}   rows out of tolerance {cnt:6d}{flag}")
    print("PARITY PINNED" if bad == 0 else f"{bad} columns differ")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
