"""Event sharding for multi-GPU runs: one process per GPU, contiguous event slices, no collective on the data
path (every output row depends only on its own waveform).  The only communication is the
host-side gather of the output tables."""
from typing import List, Tuple


def event_slice(n_events: int, rank: int, world_size: int) -> Tuple[int, int]:
    """contiguous slice [start, stop) of rank `rank`: sizes differ by at most one, earlier ranks take the remainder"""
    if world_size < 1 or not (0 <= rank < world_size) or n_events < 0:
        raise ValueError("bad sharding arguments")
    base, rem = divmod(n_events, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_slices(n_events: int, world_size: int) -> List[Tuple[int, int]]:
    return [event_slice(n_events, r, world_size) for r in range(world_size)]


def gather_rows(local_rows, n_events: int, group=None):
    """host-side gather of per-rank output rows to rank 0 with torch.distributed (gloo or nccl process group
    already initialised).  local_rows: torch tensor [n_local, ncol] (CPU for gloo, CUDA for nccl).
    Returns the concatenated [n_events, ncol] tensor on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [b - a for a, b in all_slices(n_events, world)]
    ncol = local_rows.shape[1]
    assert local_rows.shape[0] == sizes[rank]
    pad = max(sizes)
    buf = torch.zeros((pad, ncol), dtype=local_rows.dtype, device=local_rows.device)
    buf[: sizes[rank]] = local_rows
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    if dist.get_backend(group) == "nccl":
        allb = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(allb, buf, group=group)
        outs = allb if rank == 0 else None
    else:
        dist.gather(buf, outs, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([o[:s] for o, s in zip(outs, sizes)], dim=0)
