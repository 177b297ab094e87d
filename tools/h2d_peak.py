"""Host-link ceilings of the box, per rank (the ceilings of bench.py's e2e figures: 16 KB per waveform).

  python tools/h2d_peak.py                                     one process
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_peak.py     N ranks at once

Every rank copies from its own host buffers to its own GPU at the same time (barrier in front of every measurement); rank 0
prints one JSON line with the per-rank GB/s of
  h2d_pinned        cudaMemcpyAsync from page-locked memory (torch pin_memory), 64 MB copies
  h2d_pinned_wc     the same from write-combined page-locked memory (cudaHostAllocWriteCombined)
  d2h_pinned        device -> page-locked memory
  host_memcpy       pageable -> page-locked staging copy with the library's thread count (the ceiling of e2e_pageable)
"""
import ctypes as C
import json
import os
import threading
import time

import numpy as np
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

N = 1 << 30
CH = 64 << 20
dev = torch.empty(N, dtype=torch.uint8, device="cuda")
pin = torch.empty(N, dtype=torch.uint8).pin_memory()
rt = C.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else C.CDLL("libcudart.so")
wc_ptr = C.c_void_p()
rc = rt.cudaHostAlloc(C.byref(wc_ptr), C.c_size_t(N), C.c_uint(4))   # cudaHostAllocWriteCombined


def sync_all():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def timed(fn, reps=4):
    fn()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return N * reps / (time.perf_counter() - t0) / 1e9


def h2d_torch():
    for o in range(0, N, CH):
        dev[o:o + CH].copy_(pin[o:o + CH], non_blocking=True)


def d2h_torch():
    for o in range(0, N, CH):
        pin[o:o + CH].copy_(dev[o:o + CH], non_blocking=True)


def h2d_wc():
    for o in range(0, N, CH):
        rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr() + o), C.c_void_p(wc_ptr.value + o), C.c_size_t(CH), C.c_int(1), C.c_void_p(0))


threads = int(os.environ.get("LGDSP_COPY_THREADS", "8"))
page = np.ones(N, dtype=np.uint8)
stage = pin.numpy()


def host_copy():
    def work(t):
        a, b = N * t // threads, N * (t + 1) // threads
        C.memmove(stage.ctypes.data + a, page.ctypes.data + a, b - a)
    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [t.start() for t in th]
    [t.join() for t in th]


res = {"h2d_pinned": timed(h2d_torch), "d2h_pinned": timed(d2h_torch), "host_memcpy": timed(host_copy, reps=2)}
if rc == 0:
    res["h2d_pinned_wc"] = timed(h2d_wc)
vals = torch.tensor([res.get(k, 0.0) for k in ("h2d_pinned", "h2d_pinned_wc", "d2h_pinned", "host_memcpy")], dtype=torch.float64, device="cuda")
if world > 1:
    allv = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(allv, vals)
else:
    allv = [vals]
if rank == 0:
    table = torch.stack(allv).cpu().numpy()
    out = {"ranks": world, "copy_threads": threads, "host_cores": len(os.sched_getaffinity(0)), "unit": "GB/s per rank"}
    for i, k in enumerate(("h2d_pinned", "h2d_pinned_wc", "d2h_pinned", "host_memcpy")):
        out[k] = [round(float(x), 2) for x in table[:, i]]
        out[k + "_sum"] = round(float(table[:, i].sum()), 1)
    out["e2e_ceiling_Mwf_s_at_16KB"] = round(out["h2d_pinned_sum"] * 1e9 / 16384 / 1e6, 2)
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
