"""End-to-end throughput of the host entry points (H2D + kernels + D2H inside the timed region) for pinned, pageable and
radware-encoded host input.  usage (GPU box): python tools/e2e_bench.py [n_events] [steps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import legenddsp.jl_b200 as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
h = L.Handle(0)
P = L.resolve_icpc_params(L.tiefree_config(), L.us(500.0))
h.icpc_set_params(P)
wf = L.synth.generate_host(n, first_event=0)
rows = np.empty((n, 49))
pin_in = torch.from_numpy(wf.view(np.int16)).pin_memory()
pin_out = torch.empty((n, 49), dtype=torch.float64).pin_memory()
enc = L.encode_waveforms(wf, L.RADWARE_SIGCOMPRESS)
pin_enc = torch.from_numpy(enc.data).pin_memory()
pin_off = torch.from_numpy(enc.offsets).pin_memory()


def run(name, fn, nbytes):
    fn()
    h.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    h.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print(json.dumps({"input": name, "Mwf_s": n / dt / 1e6, "ms": dt * 1e3, "h2d_GBs": nbytes / dt / 1e9}), flush=True)


run("pinned uint16", lambda: h.icpc_run_host(None, pin_in.data_ptr(), n, 8192, pin_out.data_ptr()), wf.nbytes)
ref = pin_out.numpy().copy()
run("pageable uint16 (numpy)", lambda: h.icpc_run_host(None, wf.ctypes.data, n, 8192, rows.ctypes.data), wf.nbytes)
assert np.array_equal(np.nan_to_num(rows), np.nan_to_num(ref))
run("radware-encoded, pinned", lambda: h.icpc_run_encoded_host(None, enc.codec, pin_enc.data_ptr(), pin_off.data_ptr(), enc.shift, 2, None, n,
                                                             pin_out.data_ptr()), enc.nbytes)
assert np.array_equal(np.nan_to_num(pin_out.numpy()), np.nan_to_num(ref))
run("radware-encoded, pageable", lambda: h.icpc_run_encoded_host(None, enc.codec, enc.data.ctypes.data, enc.offsets.ctypes.data, enc.shift, 2,
                                                               None, n, rows.ctypes.data), enc.nbytes)
assert np.array_equal(np.nan_to_num(rows), np.nan_to_num(ref))
print("bytes/event encoded:", enc.nbytes / n, "threads:", os.environ.get("LGDSP_COPY_THREADS", "default"))
h.close()
