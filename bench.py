#!/usr/bin/env python
"""bench.py -- throughput of the dsp_icpc hot path (waveforms/s, 8192-sample UInt16 ICPC waveforms).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload W] [--batch B]

One "step" = one pass of the fused kernel over one batch of synthetic waveforms that is already resident in
HBM (`value`), and the same through the reference-facing C-ABI call with pinned HOST buffers (`e2e`).
Workloads: dsp_icpc (default; full 49-column chain incl. CUSP/ZAC = BASELINE.json configs[2]/[4]),
pz_trap (configs[1]), trap_sweep (configs[3], 200 trapezoid variants).
Multi-GPU: one process per GPU under torchrun, events sharded, no collective on the data path ("weak" scaling:
per-GPU batch fixed).  `--impl reference` times the CPU port of the reference algorithm (oracle/) on rank 0.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_IN = 8192 * 2
NCOL = 49
# DRAM traffic per waveform of icpc_kernel from the committed ncu --set full capture (profiles/r01_icpc_kernel_ncu_full.txt:
# dram__bytes_read.sum + dram__bytes_write.sum = 269.97 MB + 11.56 MB for 16 384 events); roofline.traffic scales it to
# the events of one launch.  Algorithmic bytes are 16 776 B/waveform: no re-reads.
NCU_DRAM_BYTES_PER_WF = {"dsp_icpc": (269.97e6 + 11.56e6) / 16384.0}


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "of measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "of fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """host cores this process may use (torchrun exports OMP_NUM_THREADS=1, so ask the scheduler, not OpenMP)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


PRESUM, WDW_FROM, WDW_N = 8, 2600, 1400   # the compressed workload: 1024 presummed uint32 + 1400 windowed uint16 samples


def compressed_setup(L, cfg, tau, builders=None):
    return L.resolve_compressed_params(cfg, tau, None, presum_rate=PRESUM, n_pre=8192 // PRESUM, step_pre=L.ns(16.0 * PRESUM),
                                       n_wdw=WDW_N, t_first_wdw=L.ns(16.0 * WDW_FROM), step_wdw=L.ns(16.0), builders=builders)


SIPM_N, SIPM_CAP = 6250, 32


def sipm_setup(L, builders=None):
    cfg = L.example_sipm_config()
    cfg["filters"]["sg"].update(min_threshold=-3.0, max_threshold=3.0, min_dc_threshold=-40.0, max_dc_threshold=40.0)
    cfg["filters"]["trap"].update(min_threshold=-15.0, max_threshold=15.0, min_dc_threshold=-30.0, max_dc_threshold=30.0)
    return L.resolve_sipm_params(cfg, {"sg": {"wl": L.ns(200.0)}}, n_samples=SIPM_N, builders=builders, max_triggers=SIPM_CAP)


def sipm_events(torch, n_events, seed, device):
    """raw UInt16 SiPM-like traces (baseline 2000 ADC, white noise sigma 2, 0..6 photo-electron pulses): int16 bit patterns"""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    k = torch.arange(SIPM_N, device=device, dtype=torch.float32)[None, :]
    y = 2000.0 + 2.0 * torch.randn((n_events, SIPM_N), generator=g, device=device)
    for q in range(6):
        on = (torch.rand((n_events, 1), generator=g, device=device) < 0.5).float()
        s0 = 100.0 + (SIPM_N - 500.0) * torch.rand((n_events, 1), generator=g, device=device)
        amp = 15.0 + 105.0 * torch.rand((n_events, 1), generator=g, device=device)
        d = (k - s0).clamp(min=0.0)
        y += on * amp * (k >= s0).float() * (1.0 - torch.exp(-d / 3.0)) * torch.exp(-d / 30.0)
    return y.round().clamp(0, 65535).to(torch.int32).to(torch.int16)


def cpu_port_throughput(L, O, P, n_events, workload, variants=None, sparams=None, reps=1, threads=0):
    """the CPU port of the reference algorithm (oracle/) on all host cores: waveforms/s"""
    if workload == "sipm":
        import torch
        wf = sipm_events(torch, n_events, 99, "cpu").numpy().view("uint16")
        t0 = time.perf_counter()
        O.dsp_sipm(P, wf, n_threads=threads)
        return n_events / (time.perf_counter() - t0)
    wf = L.synth.generate_host(n_events, first_event=10_000_000)
    if workload == "compressed":
        pre, wdw = L.synth.compress(wf, PRESUM, (WDW_FROM, WDW_N))
        Pp, Pw, aux = P
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        if workload == "compressed":
            O.dsp_icpc_compressed(Pp, Pw, pre, wdw, PRESUM, aux, n_threads=threads)
        elif workload == "trap_sweep":
            O.trap_sweep(sparams, wf, variants, n_threads=threads)
        else:
            O.dsp_icpc(P, wf, n_threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_events / best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dsp_icpc", choices=["dsp_icpc", "pz_trap", "trap_sweep", "compressed", "sipm"])
    ap.add_argument("--batch", type=int, default=131072, help="waveforms per step per GPU (131072 = 2.1 GB >> L2)")
    ap.add_argument("--pool", type=int, default=4, help="distinct resident batches cycled by the steps")
    ap.add_argument("--direct", action="store_true", help="CUSP/ZAC as direct 2375-tap FIRs (validation mode)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=0, help="events of the cpu_baseline sample (0: auto)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--groups", default=None, help="override the column-group mask (hex), for per-stage timing experiments")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import importlib
    L = importlib.import_module("legenddsp.jl_b200")

    cfg = L.tiefree_config()
    tau = L.us(500.0)
    groups = {"dsp_icpc": L._abi.GROUP_ALL, "pz_trap": L._abi.GROUP_PZTRAP, "trap_sweep": 0, "compressed": 0, "sipm": 0}[args.workload]
    wl_name = {"dsp_icpc": "full dsp_icpc, 49 columns incl. CUSP+ZAC (BASELINE configs[2]/[4])",
               "pz_trap": "pole-zero + trapezoid energies/t0 only (BASELINE configs[1])",
               "trap_sweep": "20x10 trapezoid (rt, ft) sweep, 200 variants (BASELINE configs[3])",
               "sipm": "dsp_sipm trigger chain on 6250-sample UInt16 SiPM traces (SURVEY 8f rank 3)",
               "compressed": "dsp_icpc_compressed: 1024 presummed uint32 (x8) + 1400 windowed uint16 samples per event "
                             "(SURVEY 8f rank 1)"}[args.workload]
    if args.groups is not None:
        groups = int(args.groups, 16)
        wl_name += f" [experimental group mask {groups:#x}]"
    out_bytes = {"dsp_icpc": NCOL * 8, "pz_trap": 5 * 8, "trap_sweep": 200 * 4, "compressed": 65 * 8,
                 "sipm": (24 + 16 * SIPM_CAP) * 8}[args.workload]
    bytes_in = {"compressed": (8192 // PRESUM) * 4 + WDW_N * 2, "sipm": SIPM_N * 2}.get(args.workload, BYTES_IN)
    bytes_per_wf = bytes_in + out_bytes
    variants = sparams = None
    if args.workload == "trap_sweep":
        rts = [L.us(1.0 + 0.75 * i) for i in range(20)]
        fts = [L.us(1.0 + 0.3 * i) for i in range(10)]
        variants = L.trap_variants(rts, fts, L.ns(16.0), mode="ft")

    # ------------------------------------------------------------------------------------------
    # reference arm: the CPU port of the reference algorithm (the reference itself is Julia and cannot be
    # installed here -- DESIGN.md section 3), all host threads, bounded sample per step
    # ------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle import oracle as O
        P = L.resolve_icpc_params(cfg, tau, groups=groups or L._abi.GROUP_ALL, builders=O.OracleBuilders())
        if args.workload == "trap_sweep":
            sparams = L.resolve_sweep_params(cfg, tau, builders=O.OracleBuilders())
        threads = host_threads()
        n_s = args.cpu_sample or (32 * threads if args.workload != "trap_sweep" else 16 * threads)
        wf = L.synth.generate_host(n_s, first_event=10_000_000)
        if args.workload == "compressed":
            Pp, Pw, aux = compressed_setup(L, cfg, tau, O.OracleBuilders())
            pre, wdw = L.synth.compress(wf, PRESUM, (WDW_FROM, WDW_N))
        if args.workload == "sipm":
            import torch
            Ps = sipm_setup(L, O.OracleBuilders())
            wf = sipm_events(torch, n_s, 99, "cpu").numpy().view("uint16")

        def step():
            if args.workload == "sipm":
                O.dsp_sipm(Ps, wf, n_threads=threads)
            elif args.workload == "compressed":
                O.dsp_icpc_compressed(Pp, Pw, pre, wdw, PRESUM, aux, n_threads=threads)
            elif args.workload == "trap_sweep":
                O.trap_sweep(sparams, wf, variants, n_threads=threads)
            else:
                O.dsp_icpc(P, wf, n_threads=threads)
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
        val = n_s * args.steps / dt
        sample = f"{n_s} synthetic waveforms per step (same generator/config as the GPU arm), float64, OpenMP over events"
        print(json.dumps({
            "impl": "reference", "metric": "waveforms/sec for dsp_icpc (8192-sample)", "value": val,
            "unit": "waveforms/s", "n_gpus": 0, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_name, "n_samples": 8192, "events_per_step": n_s,
                       "note": "CPU restatement of the reference algorithm (oracle/), not Julia: no Julia toolchain in this image"},
            "cpu_baseline": {"value": val, "unit": "waveforms/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "waveforms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    # ------------------------------------------------------------------------------------------
    # our arm
    # ------------------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: the communicator initialisation prints "NCCL version ..." to fd 1, so fd 1
        # points at stderr while the process group and its communicator come up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    # a dedicated (non-default) torch stream shared with the library, so torch's CUDA events time our kernels
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    h = L.Handle(local_rank, stream=stream.cuda_stream)

    P = L.resolve_icpc_params(cfg, tau, groups=groups or L._abi.GROUP_ALL, cuspzac_direct=args.direct)
    if args.workload == "trap_sweep":
        sparams = L.resolve_sweep_params(cfg, tau)
    elif args.workload == "compressed":
        Pp, Pw, aux = compressed_setup(L, cfg, tau)
    elif args.workload == "sipm":
        Ps = sipm_setup(L)
    else:
        h.icpc_set_params(P)

    B = args.batch
    n_pool = max(1, args.pool)
    # resident pool of distinct batches, generated on the device (each rank its own slice of the event stream)
    if args.workload == "sipm":
        pool = torch.stack([sipm_events(torch, B, 1000 + rank * n_pool + k, dev) for k in range(n_pool)])
        out_t = torch.empty((B, 4, 4, SIPM_CAP), dtype=torch.float64, device=dev)
    else:
        pool = torch.empty((n_pool, B, 8192), dtype=torch.int16, device=dev)
        for k in range(n_pool):
            L.synth.generate_device(h, pool[k].data_ptr(), B, first_event=(rank * n_pool + k) * B)
    if args.workload == "trap_sweep":
        out = torch.empty((B, 200), dtype=torch.float32, device=dev)
    else:
        out = torch.empty((B, NCOL), dtype=torch.float64, device=dev)
    h.synchronize()
    if args.workload == "compressed":
        # the DAQ's compressed format made from the full traces (input generation, outside the timed region); only the
        # compressed buffers stay resident
        torch.cuda.synchronize()
        full = pool.view(n_pool, B, 8192).to(torch.int32) & 0xFFFF
        c_pre = full.view(n_pool, B, 8192 // PRESUM, PRESUM).sum(dim=3, dtype=torch.int32).contiguous()
        c_wdw = pool[:, :, WDW_FROM:WDW_FROM + WDW_N].contiguous()
        del full, pool
        out_w = torch.empty((B, NCOL), dtype=torch.float64, device=dev)
        out_s = torch.empty((B, 5, 5), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        h.icpc_compressed_run_device(Pp, Pw, c_pre[0].data_ptr(), 4, 8192 // PRESUM, c_wdw[0].data_ptr(), 2, WDW_N, float(PRESUM),
                                     aux, B, out.data_ptr(), out_w.data_ptr(), out_s.data_ptr())
        h.synchronize()

    def step(k):
        if args.workload == "sipm":
            h.sipm_run_device(Ps, pool[k % n_pool].data_ptr(), B, SIPM_N, out.data_ptr(), out_t.data_ptr())
        elif args.workload == "compressed":
            h.icpc_compressed_run_device(None, None, c_pre[k % n_pool].data_ptr(), 4, 8192 // PRESUM, c_wdw[k % n_pool].data_ptr(),
                                         2, WDW_N, float(PRESUM), aux, B, out.data_ptr(), out_w.data_ptr(), out_s.data_ptr())
        elif args.workload == "trap_sweep":
            h.sweep_run_device(sparams, pool[k % n_pool].data_ptr(), B, 8192, variants, out.data_ptr())
        else:
            h.icpc_run_device(None, pool[k % n_pool].data_ptr(), B, 8192, out.data_ptr())

    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = h.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record(stream)
    for k in range(args.steps):
        step(k)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = h.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max * 1e-3)

    # ---- e2e: pinned host buffers through the host C-ABI call (H2D + kernel + D2H inside the timed region) ----
    Be = min(B, 65536)
    if args.workload == "compressed":
        host_pre = torch.empty((Be, 8192 // PRESUM), dtype=torch.int32).pin_memory()
        host_wdw = torch.empty((Be, WDW_N), dtype=torch.int16).pin_memory()
        host_pre.copy_(c_pre[0][:Be])
        host_wdw.copy_(c_wdw[0][:Be])
        host_out_w = torch.empty((Be, NCOL), dtype=torch.float64).pin_memory()
        host_out_s = torch.empty((Be, 25), dtype=torch.float64).pin_memory()
    else:
        host_in = torch.empty((Be, SIPM_N if args.workload == "sipm" else 8192), dtype=torch.int16).pin_memory()
        host_in.copy_(pool[0][:Be])
        host_out_t = torch.empty((Be, 16 * SIPM_CAP), dtype=torch.float64).pin_memory() if args.workload == "sipm" else None
    if args.workload == "trap_sweep":
        host_out = torch.empty((Be, 200), dtype=torch.float32).pin_memory()
    else:
        host_out = torch.empty((Be, NCOL), dtype=torch.float64).pin_memory()

    def e2e_step():
        if args.workload == "sipm":
            h.sipm_run_host(Ps, host_in.data_ptr(), Be, SIPM_N, host_out.data_ptr(), host_out_t.data_ptr())
        elif args.workload == "compressed":
            h.icpc_compressed_run_host(None, None, host_pre.data_ptr(), 4, 8192 // PRESUM, host_wdw.data_ptr(), 2, WDW_N,
                                       float(PRESUM), aux, Be, host_out.data_ptr(), host_out_w.data_ptr(), host_out_s.data_ptr())
        elif args.workload == "trap_sweep":
            h.sweep_run_host(sparams, host_in.data_ptr(), Be, 8192, variants, host_out.data_ptr())
        else:
            h.icpc_run_host(None, host_in.data_ptr(), Be, 8192, host_out.data_ptr())
    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * Be * args.e2e_steps / float(te.item())
    ho = host_out.double()
    checksum = float(ho[torch.isfinite(ho)].sum().item())  # the result is really read on the host

    if rank == 0:
        peaks, peak_kind = _peaks()
        achieved = B * bytes_per_wf / (ms_max * 1e-3 / args.steps) / 1e9
        line = {
            "metric": "waveforms/sec for dsp_icpc (8192-sample)", "value": value, "unit": "waveforms/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_name, "n_samples": {"compressed": [8192 // PRESUM, WDW_N], "sipm": SIPM_N}.get(args.workload, 8192),
                       "events_per_step_per_gpu": B, "resident_pool_batches": n_pool,
                       "dsp_config": "reference example config (test/test_dsp_icpc.jl:50-161), tie-free windows, tau=500us, default filter pars",
                       "cuspzac": "direct FIR" if args.direct else "structured",
                       "l2": f"inputs larger than L2: each step reads a different {B * bytes_in / 1e9:.1f} GB batch",
                       "parallelism": f"event-sharded x{world}, no data-path collective"},
            "clocks": clocks,
            "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": "waveforms/s", "h2d_bytes_per_step": Be * bytes_in,
                    "d2h_bytes_per_step": Be * (out_bytes if args.workload != "compressed" else (2 * NCOL + 25) * 8), "events_per_step_per_gpu": Be, "checksum": checksum},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"],
                         "traffic": (NCU_DRAM_BYTES_PER_WF[args.workload] * B if args.workload in NCU_DRAM_BYTES_PER_WF
                                     and args.groups is None else None),
                         "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, scaled by events)",
                         "peak_source": peak_kind,
                         "algorithmic_bytes_per_waveform": bytes_per_wf,
                         "kernel": "sweep_kernel" if args.workload == "trap_sweep" else "sipm_kernel" if args.workload == "sipm" else "icpc_kernel"
                                   + (" x2 (presummed + windowed) + window_stats_kernel" if args.workload == "compressed" else "")},
        }
        if not args.no_cpu and world == 1:
            from oracle import oracle as O
            Po = L.resolve_icpc_params(cfg, tau, groups=groups or L._abi.GROUP_ALL, builders=O.OracleBuilders())
            if args.workload == "compressed":
                Po = compressed_setup(L, cfg, tau, O.OracleBuilders())
            if args.workload == "sipm":
                Po = sipm_setup(L, O.OracleBuilders())
            so = L.resolve_sweep_params(cfg, tau, builders=O.OracleBuilders()) if args.workload == "trap_sweep" else None
            threads = host_threads()
            n_s = args.cpu_sample or 128 * threads
            v = cpu_port_throughput(L, O, Po, n_s, args.workload, variants, so, threads=threads)
            line["cpu_baseline"] = {"value": v, "unit": "waveforms/s", "cores": threads, "kind": "port",
                                    "sample": f"{n_s} waveforms of the same synthetic stream, CPU restatement of the reference "
                                              "algorithm (oracle/, float64, direct-form CUSP/ZAC FIRs), OpenMP over events; not Julia"}
        print(json.dumps(line))
    h.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
