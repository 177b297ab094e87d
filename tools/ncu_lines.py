"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: python tools/ncu_lines.py dump.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.defaultdict(lambda: [0, 0, 0, ""])   # (file,line) -> inst, samples, thread inst, source
cur_file = None; H = None
stall_cols = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": H = r; ci = H.index("Instructions Executed"); cs = H.index("# Samples"); ct = H.index("Thread Instructions Executed"); continue
    if H is None or len(r) < len(H) - 5: continue
    try:
        inst = int(r[ci]); smp = int(r[cs]); ti = int(r[ct])
    except ValueError:
        continue
    if not r[0].strip(): continue   # SASS rows under a source line (already aggregated in the line row)
    key = (cur_file, r[0])
    a = agg[key]; a[0] += inst; a[1] += smp; a[2] += ti
    if r[1] and not a[3]: a[3] = r[1].strip()[:100]
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print(f"total warp-inst {tot}, samples {tots}")
if len(sys.argv) > 3:
    # phase summary by line ranges of lgdsp_icpc.cu: a-b:name,...
    for spec in sys.argv[3].split(","):
        rng, name = spec.split(":"); a_, b_ = map(int, rng.split("-"))
        ii = sum(v[0] for (f, l), v in agg.items() if f == "lgdsp_icpc.cu" and a_ <= int(l) <= b_)
        ss = sum(v[1] for (f, l), v in agg.items() if f == "lgdsp_icpc.cu" and a_ <= int(l) <= b_)
        print(f"  {name:28s} {100*ii/tot:5.1f}% inst {100*ss/tots:5.1f}% samples")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*a[0]/tot:5.1f}% inst {100*a[1]/tots:5.1f}% smp  {f}:{l:>4}  {a[3]}")
