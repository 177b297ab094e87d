"""Device-resident throughput of the dsp_icpc paths for a list of (path, batch, streams) settings (one B200).
usage (GPU box): python tools/split_bench.py [n_events] [groups_hex] [settings: fused split:888:2 split:592:3 ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import legenddsp.jl_b200 as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
groups = int(sys.argv[2], 16) if len(sys.argv) > 2 else L._abi.GROUP_ALL
settings = sys.argv[3:] or ["fused", "split:888:2"]
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
h = L.Handle(0, stream=stream.cuda_stream)
P = L.resolve_icpc_params(L.tiefree_config(), L.us(500.0), groups=groups)
h.icpc_set_params(P)
npool = 3
pool = torch.empty((npool, n, 8192), dtype=torch.int16, device=dev)
for k in range(npool):
    L.synth.generate_device(h, pool[k].data_ptr(), n, first_event=k * n)
out = torch.empty((n, 49), dtype=torch.float64, device=dev)
h.synchronize()
res = []
for s in settings:
    f = s.split(":")
    path = f[0]
    if path == "prof":   # prof:<events>: serial per-kernel device times of the split pipeline on one batch
        m = int(f[1]) if len(f) > 1 else 16384
        best = None
        for k in range(4):
            ms = h.icpc_profile_device(pool[0].data_ptr(), m, 8192, out.data_ptr())
            best = ms if best is None else [min(a, b) for a, b in zip(best, ms)]
        print(json.dumps({"setting": s, "events": m, "ms_prefix_extract_select_finish": best, "sum_ms": sum(best),
                          "Mwf_s_serial": m / sum(best) / 1e3}), flush=True)
        continue
    batch = int(f[1]) if len(f) > 1 else 0
    streams = int(f[2]) if len(f) > 2 else 0
    h.set_icpc_path(path, batch, streams)
    for k in range(2):
        h.icpc_run_device(None, pool[k % npool].data_ptr(), n, 8192, out.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 6
    e0.record(stream)
    for k in range(steps):
        h.icpc_run_device(None, pool[k % npool].data_ptr(), n, 8192, out.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    r = {"setting": s, "ms_per_step": ms, "Mwf_s": n / ms / 1e3, "checksum": float(out[torch.isfinite(out)].sum().item())}
    print(json.dumps(r), flush=True)
    res.append(r)
h.close()
