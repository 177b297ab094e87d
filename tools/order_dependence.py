"""Does the row of an event depend on its position in the batch / on the event processed before it by the same CTA?
usage (GPU box): python tools/order_dependence.py [n_events]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
import numpy as np
import torch
L = importlib.import_module("legenddsp.jl_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
P = L.resolve_icpc_params(L.tiefree_config(), L.us(500.0))
h = L.Handle(0)
h.icpc_set_params(P)
wf = torch.empty((n, 8192), dtype=torch.int16, device="cuda")
L.synth.generate_device(h, wf.data_ptr(), n, first_event=50_000_000)
rows = torch.empty((n, L.NCOL), dtype=torch.float64, device="cuda")
h.icpc_run_device(None, wf.data_ptr(), n, 8192, rows.data_ptr()); h.synchronize()
wf_r = torch.empty_like(wf)
for a0 in range(0, n, 65536):    # reversed copy in blocks (torch.flip mis-indexes tensors of more than 2^31 elements)
    b0 = min(n, a0 + 65536)
    wf_r[n - b0:n - a0] = torch.flip(wf[a0:b0], dims=(0,))
torch.cuda.synchronize()   # the library launches on the handle's own (non-blocking) stream: torch's copies must be done
rows_r = torch.empty_like(rows)
h.icpc_run_device(None, wf_r.data_ptr(), n, 8192, rows_r.data_ptr()); h.synchronize()
a = rows.cpu().numpy(); b = torch.flip(rows_r, dims=(0,)).cpu().numpy()
bits = a.view(np.int64) != b.view(np.int64)
print("events with a differing row:", int(bits.any(axis=1).sum()), "of", n)
for j, name in enumerate(L.COLUMNS):
    k = int(bits[:, j].sum())
    if k:
        d = np.abs(a[:, j] - b[:, j])
        e = int(np.argmax(bits[:, j]))
        print(f"  {name:18s} {k:7d} differing, max abs diff {np.nanmax(d):.3e}, nan-mismatch {int((np.isnan(a[:, j]) != np.isnan(b[:, j])).sum())}, first event {e}: {a[e, j]!r} vs {b[e, j]!r}")
