"""Device-resident throughput of the decode_data kernels (one B200): encoded bytes already in HBM -> UInt16/UInt32 waveforms in HBM.
usage (GPU box): python tools/codec_bench.py [n_events]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import legenddsp.jl_b200 as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
h = L.Handle(0, stream=stream.cuda_stream)
wf = L.synth.generate_host(n, first_event=0)
for codec, name, arr in ((L.RADWARE_SIGCOMPRESS, "radware", wf), (L.ULEB128_ZIGZAG_DIFF, "uleb128zzd_u16", wf),
                         (L.ULEB128_ZIGZAG_DIFF, "uleb128zzd_u32", wf[:, ::2].astype(np.uint32) * 8)):
    enc = L.encode_waveforms(arr, codec)
    ns = arr.shape[1]
    d_enc = torch.from_numpy(np.ascontiguousarray(enc.data)).to(dev)
    d_off = torch.from_numpy(np.ascontiguousarray(enc.offsets)).to(dev)
    out = torch.empty((n, ns), dtype=torch.int16 if arr.dtype == np.uint16 else torch.int32, device=dev)
    run = lambda: h.decode_data_device(codec, d_enc.data_ptr(), d_off.data_ptr(), n, ns, enc.shift, out.data_ptr(), enc.sample_bytes, ns)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record(stream)
    for _ in range(steps):
        run()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    got = out.cpu().numpy().view(arr.dtype)
    print(json.dumps({"codec": name, "events": n, "n_samples": ns, "bytes_per_event": float(enc.offsets[-1]) / n, "ms": ms,
                      "Mev_s": n / ms / 1e3, "GB_s_out": n * ns * arr.dtype.itemsize / ms / 1e6, "exact": bool(np.array_equal(got, arr))}), flush=True)
h.close()
