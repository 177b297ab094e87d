"""Raw pinned-host -> device copy bandwidth of the box (the ceiling of the `e2e` figure of bench.py: 16 KB per waveform).
usage: python tools/h2d_peak.py"""
import torch

n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk_mb in (1024, 64, 32):
    c = chunk_mb << 20
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 5
    for _ in range(reps):
        for o in range(0, n, c):
            d[o:o + c].copy_(h[o:o + c], non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"H2D pinned, {chunk_mb:5d} MB copies: {n / ms / 1e6:7.2f} GB/s  -> ceiling {n / ms / 1e6 * 1e9 / 16384 / 1e6:.3f} M waveforms/s at 16 KB per waveform")
