"""Opcode histogram of every kernel of liblgdsp_b200.so (cuobjdump -sass; run here, no GPU needed) plus the lines that show
the bulk-copy / mbarrier machinery (UBLKCP = cp.async.bulk, SYNCS = mbarrier) and the FP64 / reduction instructions.
usage: python tools/sass_opcodes.py [outdir=profiles] [prefix=r02]   ->  <outdir>/<prefix>_sass_opcodes_<kernel>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "legenddsp.jl_b200", "liblgdsp_b200.so")
outdir = os.path.join(ROOT, sys.argv[1] if len(sys.argv) > 1 else "profiles")
prefix = sys.argv[2] if len(sys.argv) > 2 else "r02"
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
demangle = lambda s: subprocess.run(["cu++filt", s], capture_output=True, text=True).stdout.strip() or s
funcs = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        funcs[cur].append(line)
ins_re = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
SHOW = ("UBLKCP", "SYNCS", "UTMA", "REDUX", "BAR.", "DFMA", "DADD", "DMUL", "DSETP", "MUFU")
index = []
for name, lines in funcs.items():
    pretty = demangle(name)
    # (template arguments carry casts such as "(unsigned int)127": drop them before cutting at the argument list, so that every
    #  instantiation gets its own file)
    short = re.sub(r"\((?:unsigned )?(?:int|long|short|char|bool)\)", "", re.sub(r"^void ", "", pretty)).split("(")[0].replace("lgdsp::", "")
    short = re.sub(r"[^A-Za-z0-9]+", "_", short.replace("(anonymous namespace)::", "")).strip("_").replace("unnamed_", "")
    hist = collections.Counter()
    base = collections.Counter()
    shown = collections.defaultdict(list)
    for l in lines:
        m = ins_re.search(l)
        if not m:
            continue
        op = m.group(2)
        hist[op] += 1
        base[op.split(".")[0]] += 1
        for s in SHOW:
            if op.startswith(s) and len(shown[s]) < 6:
                shown[s].append(l.split("/*", 2)[0].rstrip() if False else re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip())
    n = sum(hist.values())
    path = os.path.join(outdir, f"{prefix}_sass_opcodes_{short}.txt")
    with open(path, "w") as f:
        f.write(f"{pretty}\n{n} SASS instructions ({n * 16} bytes of code), sm_100a, from cuobjdump -sass of liblgdsp_b200.so\n\n")
        f.write("opcode families (count, share):\n")
        for op, c in base.most_common():
            f.write(f"  {op:12s} {c:6d}  {100.0 * c / n:5.1f} %\n")
        f.write("\nfull opcodes (count):\n")
        for op, c in hist.most_common():
            f.write(f"  {op:36s} {c:6d}\n")
        f.write("\nfirst occurrences of the marker instructions (address, instruction):\n")
        for s in SHOW:
            tot = sum(c for op, c in hist.items() if op.startswith(s))
            f.write(f"  -- {s}: {tot} instructions\n")
            for l in shown[s]:
                f.write(f"     {l}\n")
    index.append((short, n, base.get("UBLKCP", 0), base.get("SYNCS", 0), base.get("DFMA", 0) + base.get("DADD", 0) + base.get("DMUL", 0),
                  base.get("IMAD", 0), base.get("BAR", 0)))
with open(os.path.join(outdir, f"{prefix}_sass_opcodes_INDEX.txt"), "w") as f:
    f.write("kernel, SASS instructions, UBLKCP, SYNCS, DFMA+DADD+DMUL, IMAD, BAR\n")
    for r in index:
        f.write(", ".join(str(x) for x in r) + "\n")
print("\n".join(", ".join(str(x) for x in r) for r in index))
