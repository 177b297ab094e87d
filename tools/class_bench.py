"""Per-kernel device time of the split pipeline by EVENT CLASS of the synthetic population (amplitude bands, empty, over-range,
second pulse): where the data-dependent pruning pays and where it does not.
usage (GPU box): python tools/class_bench.py [events_per_class=8192]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import legenddsp.jl_b200 as L

m = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
h = L.Handle(0, stream=stream.cuda_stream)
P = L.resolve_icpc_params(L.tiefree_config(), L.us(500.0))
h.icpc_set_params(P)
wf = L.synth.generate_host(24 * m, first_event=0)
base = np.median(wf[:, :2000].astype(np.float64), axis=1)
amp = wf.max(axis=1).astype(np.float64) - base
sat = wf.max(axis=1) >= 65520
late = wf[:, 7000:].astype(np.float64).mean(axis=1) - base
classes = {
    "all (population mix)": np.ones(len(wf), bool),
    "empty (amp = 0)": amp < 25,
    "amp 50-150": (amp >= 35) & (amp < 150) & ~sat,
    "amp 150-1000": (amp >= 150) & (amp < 1000) & ~sat,
    "amp 1000-10000": (amp >= 1000) & (amp < 10000) & ~sat,
    "amp > 10000": (amp >= 10000) & ~sat,
    "over-range (clipped)": sat,
}
out = torch.empty((m, 49), dtype=torch.float64, device=dev)
for name, sel in classes.items():
    idx = np.nonzero(sel)[0]
    if len(idx) == 0:
        continue
    idx = np.resize(idx, m)   # repeat the class members up to m events
    d = torch.from_numpy(wf[idx].view(np.int16)).to(dev)
    best = None
    for _ in range(4):
        ms = h.icpc_profile_device(d.data_ptr(), m, 8192, out.data_ptr())
        best = ms if best is None else [min(a, b) for a, b in zip(best, ms)]
    print(json.dumps({"class": name, "members": int(sel.sum()), "share": float(sel.mean()), "events": m,
                      "us_per_1000_events_prefix_extract_select_finish": [round(1e6 * x / m, 2) for x in best],
                      "Mwf_s_serial": m / sum(best) / 1e3}), flush=True)
h.close()
