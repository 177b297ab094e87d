"""Event sharding + host-side gather with world_size 2 on the gloo backend (CPU)."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_event_slices(L):
    sh = L.sharding
    assert sh.all_slices(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert sh.all_slices(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    for n in (0, 1, 7, 100_000_000):
        for w in (1, 2, 4, 8):
            sl = sh.all_slices(n, w)
            assert sl[0][0] == 0 and sl[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            sizes = [b - a for a, b in sl]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_events, q):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import torch.distributed as dist
    import legenddsp.jl_b200 as L
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = L.sharding.event_slice(n_events, rank, world)
    # each rank generates and processes ITS slice of the event stream (the CPU oracle stands in for the GPU here:
    # this test covers the sharding/gather logic, which is identical for the CUDA path)
    wf = L.synth.generate_host(b - a, first_event=a)
    P = L.resolve_icpc_params(L.example_config(), L.us(500.0), builders=O.OracleBuilders(),
                              groups=L._abi.GROUP_PZTRAP)
    P.cusp.n_taps = P.zac.n_taps = 64   # keep the CPU work tiny
    rows, _ = O.dsp_icpc(P, wf, n_threads=1)
    full = L.sharding.gather_rows(torch.from_numpy(rows), n_events)
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process(L, O):
    import numpy as np
    import torch.multiprocessing as mp
    n_events = 9
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_events, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    wf = L.synth.generate_host(n_events)
    P = L.resolve_icpc_params(L.example_config(), L.us(500.0), builders=O.OracleBuilders(), groups=L._abi.GROUP_PZTRAP)
    P.cusp.n_taps = P.zac.n_taps = 64
    ref, _ = O.dsp_icpc(P, wf, n_threads=1)
    assert got.shape == ref.shape
    assert np.array_equal(np.nan_to_num(got, nan=-1.0), np.nan_to_num(ref, nan=-1.0))
