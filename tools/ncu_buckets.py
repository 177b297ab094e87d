"""Instruction / stall-sample share per source-line bucket from `ncu --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_buckets.py dump.csv [bucket_lines] [file]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 25
F = sys.argv[3] if len(sys.argv) > 3 else 'lgdsp_icpc.cu'
agg = collections.defaultdict(lambda: [0, 0])
cur = None; H = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": H = r; ci = H.index("Instructions Executed"); cs = H.index("# Samples"); continue
    if H is None or len(r) < len(H) - 5: continue
    try: inst = int(r[ci]); smp = int(r[cs])
    except ValueError: continue
    if not r[0].strip(): continue
    a = agg[(cur, int(r[0]))]; a[0] += inst; a[1] += smp
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print("total warp-inst", tot, "samples", tots)
b = collections.defaultdict(lambda: [0, 0])
for (f, l), a in agg.items():
    key = (f, (l // B) * B) if f == F else (f, 0)
    b[key][0] += a[0]; b[key][1] += a[1]
for k, v in sorted(b.items()):
    if v[0] / tot > 0.004 or v[1] / tots > 0.004:
        print(f"{k[0]:28s} {k[1]:5d}  {100*v[0]/tot:5.1f}% inst {100*v[1]/tots:5.1f}% smp")
