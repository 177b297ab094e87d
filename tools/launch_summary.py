"""Per-kernel summary of an ncu launch list (gpu__time_duration.sum per launch, CSV) of the default bench command, next to the
live per-kernel times of tools/split_bench.py prof:N and the DRAM bytes of the --set full capture.
usage: python tools/launch_summary.py launches.csv split_prof.json split_ncu_full.txt > profiles/rNN_ncu_launches_split.txt"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr = rows[0]
ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
seq = [(r[ik].split("(")[0].replace("void ", "").replace("lgdsp::", ""), float(r[iv].replace(",", "")), r[ig]) for r in rows[1:] if len(r) > iv]
c, t, g = collections.Counter(), collections.Counter(), {}
for k, v, gr in seq:
    c[k] += 1
    t[k] += v
    g.setdefault(k, gr)
pipe = [k for k in c if k.startswith("icpc_")]
tot = sum(t[k] for k in pipe)
print("ncu launch list of `python bench.py --steps 3 --warmup 3` (default workload, split path), first %d launches" % len(seq))
print("command: ncu --clock-control none --metrics gpu__time_duration.sum -c 1500 --csv --log-file <csv> python bench.py --steps 3 --warmup 3")
print("raw list: the .csv beside this file (per-launch times are cold-cache and serialised by ncu: the SHARE is what compares with the live run)\n")
print("synth_kernel launches fill the resident pool (before the timed region), then the pipeline: one step of 131 072 events = 32 sub-batches")
print("of 4 096 events x 4 kernels (prefix -> extract || CUSP/ZAC select -> CUSP/ZAC finish), sub-batches round-robin over 4 stream pairs.\n")
print(f"{'kernel':42s} {'launches':>8s} {'grid':>14s} {'avg us':>9s} {'total ms':>9s} {'share':>7s}")
for k in [x for x in c if not x.startswith("icpc_")] + pipe:
    share = f"{100 * t[k] / tot:6.1f}%" if k in pipe else ""
    print(f"{k:42s} {c[k]:8d} {g[k]:>14s} {t[k] / c[k] / 1e3:9.1f} {t[k] / 1e6:9.2f} {share}")
nb = c[pipe[0]]
print(f"\npipeline kernels: {tot / 1e6:.2f} ms for {nb} sub-batches = {nb * 4096} events -> {tot / nb / 1e3:.1f} us per 4 096 events serialised"
      f" = {nb * 4096 / (tot / 1e9) / 1e6:.2f} M wf/s under ncu")
live = [json.loads(l) for l in open(sys.argv[2]) if l.startswith("{")]
names = ("prefix", "extract", "select", "finish")
for d in live:
    ms = d["ms_prefix_extract_select_finish"]
    s = sum(ms)
    print(f"live (CUDA events, warm, one batch of {d['events']} events run serially, tools/split_bench.py): "
          + " / ".join(f"{n} {x:.3f}" for n, x in zip(names, ms)) + f" ms = shares " + " / ".join(f"{100 * x / s:.1f}" for x in ms) + " %")
txt = open(sys.argv[3]).read()
rd = [float(x) for x in re.findall(r"dram__bytes_read.sum\s+([0-9.]+) Mbyte", txt)]
wr = [float(x) for x in re.findall(r"dram__bytes_write.sum\s+([0-9.]+) Mbyte", txt)]
if len(rd) == 4 and len(wr) == 4:
    # order of the capture: prefix, extract, select, finish
    total = sum(rd) + sum(wr)
    print("\nDRAM traffic (ncu --set full, one 4 096-event sub-batch per kernel, caches flushed between kernels; read + written MB): "
          + ", ".join(f"{n} {a:.1f} + {b:.1f}" for n, a, b in zip(names, rd, wr))
          + f" = {total:.1f} MB per 4 096 events = {total * 1e6 / 4096 / 1e3:.1f} KB/event (algorithmic 16.8 KB/event; the float64 prefix sums,"
          " 65.6 KB/event, are written once and read by three consumers through L2/HBM)")
