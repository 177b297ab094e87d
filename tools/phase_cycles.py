"""Per-phase barrier-to-barrier cycle shares of the fused dsp_icpc kernel (debug counters, see lgdsp_debug_phase_cycles).
usage (on a GPU box): python tools/phase_cycles.py [n_events] [groups_hex]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
import torch
L = importlib.import_module("legenddsp.jl_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
groups = int(sys.argv[2], 16) if len(sys.argv) > 2 else L._abi.GROUP_ALL
P = L.resolve_icpc_params(L.tiefree_config(), L.us(500.0), groups=groups)
h = L.Handle(0)
h.icpc_set_params(P)
wf = torch.empty((n, 8192), dtype=torch.int16, device="cuda")
out = torch.empty((n, 49), dtype=torch.float64, device="cuda")
L.synth.generate_device(h, wf.data_ptr(), n, first_event=0)
h.icpc_run_device(None, wf.data_ptr(), n, 8192, out.data_ptr()); h.synchronize()
h.phase_cycles(True)
h.icpc_run_device(None, wf.data_ptr(), n, 8192, out.data_ptr()); h.synchronize()
ms = h.last_kernel_ms()
c = h.phase_cycles(False)
tot = sum(c)
names = ["tma_wait", "P1", "P2", "P3", "P4a", "P4b", "P5", "-"]
print(f"kernel {ms:.3f} ms for {n} events; per-event cycles per CTA (barrier to barrier):")
for nm, v in zip(names, c):
    print(f"  {nm:9s} {v / n:9.0f} cyc/event  {100 * v / tot:5.1f}%")
print(f"  total     {tot / n:9.0f}")
sec = h.section_cycles() if hasattr(h, "section_cycles") else None
if sec and any(any(r) for r in sec):
    names = {31: "tma wait", 0: "P1 loop", 1: "P1 red+scan", 2: "B1 wait", 3: "fold/sat/blstats", 4: "P2 TT loop", 5: "t10..t99 masks",
             6: "tail log", 7: "B2 wait", 8: "resolve t10..", 9: "pz tail stats", 10: "full traps", 11: "coarse traps",
             12: "sg0 chunk pass", 13: "sg1/2+deriv", 14: "sg reductions", 15: "cz_scan", 16: "B3 wait", 17: "t50/pk/stash",
             18: "trap items", 19: "sg masks", 20: "cz_init+coarse", 21: "B4 wait", 22: "cz cand", 23: "cz_run", 24: "final partials",
             25: "cz_scan to its barrier", 26: "B6 wait", 27: "scalar jobs", 28: "cz jobs/B8 wait", 29: "queue items", 30: "B5 wait"}
    print("section cycles per event: max over warps | mean over warps | per warp")
    for i in [31] + list(range(31)):
        r = [v / n for v in sec[i]]
        print(f"  {i:2d} {names.get(i, ''):18s} {max(r):8.0f} {sum(r) / 8:8.0f} | " + " ".join(f"{v:6.0f}" for v in r))
