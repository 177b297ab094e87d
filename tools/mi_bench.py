"""Throughput of MultiIntersect (one warp per trace of doubles) on the GPU box: python tools/mi_bench.py [n_traces] [n_samples]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
import numpy as np
import torch
L = importlib.import_module("legenddsp.jl_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
h = L.Handle(0, stream=stream.cuda_stream)
k = torch.arange(ns, device="cuda", dtype=torch.float64)[None, :]
g = torch.Generator(device="cuda"); g.manual_seed(1)
s0 = 2000 + 2000 * torch.rand((n, 1), generator=g, device="cuda", dtype=torch.float64)
rise = 10 + 110 * torch.rand((n, 1), generator=g, device="cuda", dtype=torch.float64)
y = 1000.0 * ((k - s0) / rise).clamp(0, 1) + torch.randn((n, ns), generator=g, device="cuda", dtype=torch.float64)
f = L.MultiIntersect(mintot=L.ns(64.0), n=2, d=2, sampling_rate=4)
P = f.params(ns, L.ns(0.0), L.ns(16.0))
x = torch.empty((n, P.n_thresholds), dtype=torch.float64, device="cuda")
fl = torch.empty(n, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
for _ in range(3):
    h.multi_intersect_device(P, y.data_ptr(), n, ns, x.data_ptr(), fl.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
reps = 10
for _ in range(reps):
    h.multi_intersect_device(P, y.data_ptr(), n, ns, x.data_ptr(), fl.data_ptr())
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"MultiIntersect: {n} traces x {ns} f64 samples, {P.n_thresholds} thresholds: {ms:.3f} ms per launch, {n / ms / 1e3:.2f} M traces/s, "
      f"{n * ns * 8 / ms / 1e6:.0f} GB/s of trace bytes; flagged {int(fl.sum())}, finite {float(torch.isfinite(x).double().mean()):.3f}")
import json
print(json.dumps({"kernel": "multi_intersect_kernel", "traces": n, "n_samples": ns, "thresholds": int(P.n_thresholds), "ms_per_launch": ms,
                  "Mtraces_s": n / ms / 1e3, "trace_GB_s": n * ns * 8 / ms / 1e6, "flagged": int(fl.sum())}))
