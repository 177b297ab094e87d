"""Writes tests/golden/oracle_rows.json: rows of the CPU ORACLE (not of the Julia reference, which cannot run
here) on 1 fixture event + 8 mixed synthetic events, example config.  Regression pin for the oracle itself.
Run from the repo root:  python tests/golden/make_oracle_rows.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import legenddsp.jl_b200 as L  # noqa: E402
from oracle import oracle as O  # noqa: E402

FIRST = 1000
P = L.resolve_icpc_params(L.example_config(), L.us(500.0), builders=O.OracleBuilders())
wf = np.concatenate([L.synth.generate_host(1, mode=1), L.synth.generate_host(8, first_event=FIRST)])
rows, _ = O.dsp_icpc(P, wf)
out = {"generator": "tests/golden/make_oracle_rows.py (oracle/lgdsp_oracle.c), NOT the Julia reference",
       "first_event": FIRST, "columns": list(L.COLUMNS),
       "rows": [[None if np.isnan(v) else float(v) for v in r] for r in rows]}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_rows.json")
json.dump(out, open(path, "w"))
print("written", path)
