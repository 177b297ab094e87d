"""One launch set of the reference's rise-time sweep (31 rise times, fixed pick-off) for ncu captures:
  ncu --set full -k regex:sweep_warp --launch-skip 2 --launch-count 1 python tools/rt_sweep_once.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import legenddsp.jl_b200 as L
cfg, tau = L.example_config(), L.us(500.0)
h = L.Handle(0)
n = 16384
d = torch.empty((n, 8192), dtype=torch.int16, device="cuda")
L.synth.generate_device(h, d.data_ptr(), n, first_event=1)
S = L.resolve_sweep_params(cfg, tau, out_f64=True)
var = L.trap_sweep_variants(L.grid_values(cfg.e_grid_rt_trap), [L.us(2.0)], L.ns(16.0), mode="rt", pickoff=cfg.enc_pickoff_trap)
o = torch.zeros((n, len(var.array)), dtype=torch.float64, device="cuda")
for k in range(4):
    h.gsweep_run_device(S, d.data_ptr(), n, 8192, var.array, o.data_ptr())
h.synchronize()
print("ok", float(o.sum()))
