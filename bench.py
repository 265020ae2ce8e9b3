#!/usr/bin/env python
"""Benchmark of the estimator hot path: solved estimation windows per second.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload vdp|gnss] [--impl ours|reference]

A "step" solves one batch of B independent windows (default: the van der Pol MHE shape of
BASELINE.json configs[1], B = 4096 windows per GPU) with the whole-solve kernel.  Inputs are
resident in HBM for `value`; `e2e` repeats the measurement through the host-buffer API with the
H2D/D2H copies inside the timed region.  For N > 1 (torchrun, one process per GPU) every rank
solves its own B windows (weak scaling) and the estimates are all-gathered over NCCL at the end
of each step (the only collective on the path).  One JSON line is printed by rank 0.

`--impl reference` times the CPU path (the NumPy restatement in oracle/, because the
reference's CasADi/IPOPT stack is not installable here) on all host cores, on a bounded sample
of the same workload.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic flop per Gauss-Newton iteration and compulsory bytes per window
FLOP_PER_ITER = {"vdp": 146480.0, "gnss": 297342.0}
BYTES_PER_WINDOW = {"vdp": 2104.0 + 360.0, "gnss": 25184.0 + 464.0}
WORKLOAD_DESC = {
    "vdp": "van_der_pol MHE windows (K=21 nodes, n=2, m=1, 50 full-state epochs, Q=1e-2 I, nz=42)",
    "gnss": "stationary-receiver GNSS MHE windows (K=11, n=5, m=3, 51 epochs x 12 pseudoranges, nz=55)",
}


def make_problem(workload):
    from nlp_filter_b200 import shapes
    return shapes.vdp_mhe() if workload == "vdp" else shapes.gnss()


def make_batch(desc, workload, B, seed):
    from nlp_filter_b200 import synth
    return synth.vdp_mhe_batch(desc, B, seed=seed) if workload == "vdp" else synth.gnss_batch(desc, B, seed=seed)


# ---------------------------------------------------------------------------------------------
# CPU path (oracle port) -- the only place bench.py executes oracle/
# ---------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(workload, B, seed):
    from oracle import nls  # noqa: F401  (test/bench infrastructure)
    d = make_problem(workload)
    _CPU["d"] = d
    _CPU["bt"] = make_batch(d, workload, B, seed)


def _cpu_solve(i):
    from oracle import nls
    d, bt = _CPU["d"], _CPU["bt"]
    win = nls.Window(d, U=bt["U"][i], data=bt["data"][i])
    z, info = nls.solve(d, win, bt["z0"][i])
    return info["iters"]


def cpu_path_windows_per_sec(workload, n_windows, cores, seed=2000, repeats=1):
    """Solve n_windows windows of the workload with the CPU restatement on `cores` processes."""
    ctx = mp.get_context("fork")
    _cpu_init(workload, n_windows, seed)
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_solve, range(min(n_windows, cores)))          # warm the workers
        best = 0.0
        for _ in range(repeats):
            t0 = time.perf_counter()
            iters = pool.map(_cpu_solve, range(n_windows), chunksize=max(1, n_windows // (cores * 4)))
            dt = time.perf_counter() - t0
            best = max(best, n_windows / dt)
    return best, float(np.mean(iters))


# ---------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(np.max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n = args.cpu_windows or (256 if args.workload == "vdp" else 64)
    vals = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        v, it = cpu_path_windows_per_sec(args.workload, n, cores, seed=2000 + s)
        if s >= args.warmup:
            vals.append((v, time.perf_counter() - t0, it))
    value = float(np.mean([v[0] for v in vals]))
    line = {
        "impl": "reference", "metric": "solved_windows_per_sec", "value": value, "unit": "windows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * n / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "windows_per_step": n},
        "cpu_baseline": {"value": value, "unit": "windows/s", "cores": cores, "kind": "port",
                         "sample": "%d windows per step of the same synthetic workload, NumPy restatement "
                                   "(oracle/nls.py) in a %d-process pool; CasADi/IPOPT not installable" % (n, cores)},
        "e2e": {"value": value, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cpu_base = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = args.cpu_windows or (512 if args.workload == "vdp" else 128)
        v, it = cpu_path_windows_per_sec(args.workload, n, cores)        # before CUDA is initialised (fork)
        cpu_base = {"value": v, "unit": "windows/s", "cores": cores, "kind": "port",
                    "sample": "%d windows of the same synthetic workload (seed 2000), NumPy restatement "
                              "oracle/nls.py in a %d-process pool, mean %.1f iterations/window" % (n, cores, it)}

    import torch
    import torch.distributed as dist
    from nlp_filter_b200 import engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = args.windows
    desc = make_problem(args.workload)
    prob = engine.WindowProblem(desc)
    bt = make_batch(desc, args.workload, B, seed=2000 + rank)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    z0, U, data = dev(bt["z0"]), dev(bt["U"]), dev(bt["data"])
    z = torch.empty_like(z0)
    cost = torch.empty((B,), dtype=torch.float64, device="cuda")
    iters = torch.empty((B,), dtype=torch.int32, device="cuda")
    status = torch.empty((B,), dtype=torch.int32, device="cuda")
    gathered = torch.empty((world * B, desc.nz), dtype=torch.float64, device="cuda") if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")     # 256 MB > 126 MB L2
    peak_fp64 = engine.fp64_peak_tflops()

    def step_device():
        z.copy_(z0)
        prob.solve(z, U, data, None, None, None, out=(cost, iters, status))
        if world > 1:
            dist.all_gather_into_tensor(gathered, z)

    hb = engine.HostBatch(prob, B)
    hb.load(bt["z0"], bt["U"], bt["data"], None)
    h_gather = torch.empty((world * B, desc.nz), dtype=torch.float64).pin_memory() if world > 1 else None

    def step_e2e():
        hb.run()
        if world > 1:
            dist.all_gather_into_tensor(gathered, hb.d_z)

    def timed(step_fn, kernel_events=False):
        for _ in range(args.warmup):
            flush.zero_()
            step_fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = []
        for _ in range(args.steps):
            flush.zero_()                       # evict the inputs from L2 between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step_fn()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev = timed(step_device)
    it_mean = float(iters.to(torch.float64).mean().item())
    st = status.cpu().numpy()
    ms_e2e = timed(step_e2e)
    # kernel-only duration (solve kernel alone, CUDA events on the launching stream)
    kev = []
    for _ in range(max(3, args.steps)):
        flush.zero_()
        z.copy_(z0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        prob.solve(z, U, data, None, None, None, out=(cost, iters, status))
        e1.record()
        kev.append((e0, e1))
    torch.cuda.synchronize()
    ms_kernel = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        tot_it = torch.tensor([it_mean], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot_it)
        it_mean_all = float(tot_it.item()) / world
    else:
        it_mean_all = it_mean
    if rank == 0:
        total_windows = world * B * args.steps
        value = total_windows / (ms_dev * 1e-3)
        e2e = total_windows / (ms_e2e * 1e-3)
        flops_launch = FLOP_PER_ITER[args.workload] * it_mean * B
        achieved = flops_launch / (ms_kernel * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_ach = BYTES_PER_WINDOW[args.workload] * B / (ms_kernel * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "solved_windows_per_sec", "value": value, "unit": "windows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "windows_per_gpu": B, "global_windows": world * B,
                       "parallelism": "windows sharded over %d GPU(s), all-gather of estimates at the end" % world,
                       "l2": "256 MB buffer written between timed iterations (inputs are smaller than L2)",
                       "mean_iterations_per_window": it_mean_all,
                       "converged_fraction": float(np.mean(st <= 1))},
            "e2e": {"value": e2e, "unit": "windows/s", "h2d_bytes_per_step": hb.h2d_bytes, "d2h_bytes_per_step": hb.d2h_bytes,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": args.steps,
            "roofline": {"bound": "fp64", "kernel": "solve_kernel", "achieved": achieved, "peak": peak_fp64,
                         "unit": "TFLOP/s", "frac": achieved / peak_fp64 if peak_fp64 > 0 else None,
                         "traffic": traffic, "kernel_ms": ms_kernel,
                         "peak_source": "FP64 FMA probe (nlpf_fp64_probe) measured in this run; MEASURED_PEAKS.json "
                                        "has no FP64 figure",
                         "flop_per_iteration": FLOP_PER_ITER[args.workload],
                         "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s"}},
            "clocks": clocks,
        }
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vdp", choices=["vdp", "gnss"])
    ap.add_argument("--windows", type=int, default=4096, help="windows per GPU per step")
    ap.add_argument("--cpu-windows", type=int, default=0, help="size of the CPU-path sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
