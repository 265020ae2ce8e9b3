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
FLOP_PER_ITER = {"vdp": 146480.0, "gnss": 297342.0, "2rx": 1338750.0, "car": 172272.0, "ekf": 9178.0, "ls": 806.0,
                 "rc": 1517808.0, "mrx": 2052636.0}
# of which factorisation + the two triangular solves (kernel 3); the rest is kernels 1+2
FLOP_CHOL_PER_ITER = {"vdp": 28224.0, "gnss": 61508.0, "2rx": 467866.0, "car": 58320.0, "rc": 698544.0, "mrx": 731818.0}
NZ = {"vdp": 42, "gnss": 55, "car": 54, "2rx": 99}
BYTES_PER_WINDOW = {"rc": 21024.0 + 1032.0, "mrx": 62512.0 + 1048.0, "vdp": 2104.0 + 360.0, "gnss": 25184.0 + 464.0, "2rx": 7920.0 + 904.0, "car": 9840.0 + 456.0, "ekf": 448.0, "ls": 416.0}
WORKLOAD_DESC = {
    "vdp": "van_der_pol MHE windows (K=21 nodes, n=2, m=1, 50 full-state epochs, Q=1e-2 I, nz=42)",
    "gnss": "stationary-receiver GNSS MHE windows (K=11, n=5, m=3, 51 epochs x 12 pseudoranges, nz=55)",
    "2rx": "two-receiver GNSS MHE scenarios (gnss-multi-receiver.py shape: K=11, n=10, m=6, 120 pseudorange slots + 51 range + "
           "51 heading residuals, arrival cost, zA == zB folded, nz=99), 1024 scenarios in total sharded over the GPUs",
    "rc": "RC-car MHE windows (rc-car.py shape: K=21, n=6, m=2, 41 epochs x 12 pseudoranges, nz=126), start near the truth",
    "mrx": "multi-receiver MHE windows (multi-receiver.py shape: K=16, n=8, m=0, 3 global unknowns, 64 epochs x (11 pseudoranges + 11 "
           "pseudorange rates + 1 range), nz=131): flat valley, most windows run to the 300-iteration limit (DESIGN section 3)",
    "car": "vehicle MHE chains (autonomous-car.py shape: K=6, n=9, m=2, 21 epochs x 11 pseudoranges, arrival cost, bounds, nz=54), "
           "receding-horizon windows DT = 1 s, warm-started",
    "ekf": "5-state GNSS EKF (utils/ekf.py + utils/gnss.py), 256 filters x 64 predict/update steps, 12 satellites",
    "ls": "iterative GNSS least squares (utils/leastsquares.py:19-42), 16384 independent epochs x 12 satellites, ECEF",
}
METRIC = {"rc": ("solved_windows_per_sec", "windows/s"), "mrx": ("solved_windows_per_sec", "windows/s"),
          "vdp": ("solved_windows_per_sec", "windows/s"), "gnss": ("solved_windows_per_sec", "windows/s"),
          "car": ("solved_windows_per_sec", "windows/s"), "2rx": ("solved_windows_per_sec", "windows/s"),
          "ekf": ("ekf_steps_per_sec", "filter-steps/s"), "ls": ("ls_epochs_per_sec", "epochs/s")}


def make_problem(workload):
    from nlp_filter_b200 import shapes
    return {"vdp": shapes.vdp_mhe, "gnss": shapes.gnss, "2rx": shapes.two_receiver, "rc": shapes.rc, "mrx": shapes.multi_receiver}[workload]()


def make_batch(desc, workload, B, seed):
    from nlp_filter_b200 import synth
    return {"vdp": synth.vdp_mhe_batch, "gnss": synth.gnss_batch, "2rx": synth.two_receiver_batch, "rc": synth.rc_batch,
            "mrx": synth.multi_receiver_batch}[workload](desc, B, seed=seed)


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
    win = nls.Window(d, U=None if bt.get("U") is None else bt["U"][i], data=bt["data"][i],
                     xprior=None if bt.get("xprior") is None else bt["xprior"][i])
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

    def hold_load(self, step_fn, min_samples=6, max_seconds=4.0):
        """The timed regions of these workloads last 5-20 ms, nvidia-smi samples every 100 ms: keep running the same
        step (untimed) until the sampler has seen the GPU under this load a few times."""
        import torch
        if self.proc is None:
            return 0
        n0, t0, extra = len(self.rows), time.perf_counter(), 0
        while len(self.rows) - n0 < min_samples and time.perf_counter() - t0 < max_seconds:
            for _ in range(8):
                step_fn()
                extra += 1
            torch.cuda.synchronize()
        self.extra_steps = getattr(self, "extra_steps", 0) + extra
        self.load_rows = getattr(self, "load_rows", []) + self.rows[n0:]
        return extra

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
        rows = getattr(self, "load_rows", None) or self.rows
        for r in rows:
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
                "samples": len(sm), "reasons": sorted(reasons),
                "how": "nvidia-smi -lms 100 while the timed step ran back to back: the timed steps plus %d further identical "
                       "untimed steps (the timed region alone is shorter than one sampling period)" % getattr(self, "extra_steps", 0)}



# ---------------------------------------------------------------------------------------------
# secondary workloads: fused EKF and GNSS least squares (row A11)
# ---------------------------------------------------------------------------------------------
def make_ekf_inputs(Bf=256, steps=64, S=12, seed=4001):
    rng = np.random.default_rng(seed)
    pos = rng.normal(0, 5.0, (Bf, 3))
    b0, al = rng.uniform(-1e3, 1e3, Bf), rng.normal(-190, 20, Bf)
    el = np.arcsin(rng.uniform(np.sin(np.deg2rad(10)), 1.0, (Bf, S)))
    az = rng.uniform(0, 2 * np.pi, (Bf, S))
    rad = rng.uniform(2.0e7, 2.6e7, (Bf, S))
    s0 = np.stack([rad * np.cos(el) * np.sin(az), rad * np.cos(el) * np.cos(az), rad * np.sin(el)], 2)
    t = np.arange(1, steps + 1)
    sat = s0[:, None, :, :] + rng.normal(0, 1e3, (Bf, 1, S, 3)) * t[None, :, None, None]
    z = np.linalg.norm(pos[:, None, None, :] - sat, axis=3) + b0[:, None, None] + al[:, None, None] * t[None, :, None] \
        + rng.normal(0, 10.0, (Bf, steps, S))
    x0 = np.concatenate([pos + rng.normal(0, 3, (Bf, 3)), (b0 + 5)[:, None], (al + 1)[:, None]], 1)
    return dict(x0=x0, S0=np.tile(np.eye(5), (Bf, 1, 1)), u=np.zeros((Bf, steps, 3)), z=z, sat=sat,
                nsat=np.full((Bf, steps), S, dtype=np.int32), Q=np.array([0.0001, 0.0001, 0.0001, 0.1, 0.001]))


def make_ls_inputs(B=16384, S=12, seed=4002):
    rng = np.random.default_rng(seed)
    p_ref = np.array([-2699899.3380952473, -4292848.690053425, 3855186.0888884724])
    rx = p_ref[None, :] + rng.normal(0, 50.0, (B, 3))
    dirs = rng.normal(size=(B, S, 3)) + 1.5 * (p_ref / np.linalg.norm(p_ref))[None, None, :]
    dirs /= np.linalg.norm(dirs, axis=2, keepdims=True)
    sat = rx[:, None, :] + dirs * rng.uniform(2.0e7, 2.5e7, (B, S, 1))
    pr = np.linalg.norm(sat - rx[:, None, :], axis=2) + rng.uniform(-1e3, 1e3, B)[:, None] + rng.normal(0, 3.0, (B, S))
    return dict(sat=sat, pr=pr, nsat=np.full(B, S, dtype=np.int32))


def _cpu_ekf(i):
    from oracle import ekf_ls
    e = _CPU["ekf"]
    ekf_ls.ekf_run(0, e["x0"][i], e["S0"][i], e["u"][i], e["z"][i], e["sat"][i], e["nsat"][i], e["Q"], 100.0, 1.0)
    return e["z"].shape[1]


def _cpu_ls(i):
    from oracle import ekf_ls
    l = _CPU["ls"]
    return ekf_ls.iterative_ls(l["sat"][i], l["pr"][i])[2]


def cpu_aux_per_sec(workload, n_units, cores):
    """CPU restatement (oracle/ekf_ls.py) of the EKF / LS workload on `cores` processes; returns units/s."""
    ctx = mp.get_context("fork")
    if workload == "ekf":
        _CPU["ekf"] = make_ekf_inputs(Bf=n_units)
        fn, per = _cpu_ekf, _CPU["ekf"]["z"].shape[1]
    else:
        _CPU["ls"] = make_ls_inputs(B=n_units)
        fn, per = _cpu_ls, 1
    with ctx.Pool(cores) as pool:
        pool.map(fn, range(min(n_units, cores)))
        t0 = time.perf_counter()
        pool.map(fn, range(n_units), chunksize=max(1, n_units // (cores * 4)))
        dt = time.perf_counter() - t0
    return n_units * per / dt


def run_aux(args):
    """EKF / LS lines (single GPU): same JSON shape as the headline line."""
    cores = os.cpu_count() or 1
    cpu_base = None
    if not args.no_cpu_baseline:
        n = args.cpu_windows or (64 if args.workload == "ekf" else 2048)
        v = cpu_aux_per_sec(args.workload, n, cores)
        cpu_base = {"value": v, "unit": METRIC[args.workload][1], "cores": cores, "kind": "port",
                    "sample": "%d %s of the same synthetic workload, NumPy restatement oracle/ekf_ls.py, %d-process pool"
                              % (n, "filters x 64 steps" if args.workload == "ekf" else "epochs", cores)}
    import torch
    from nlp_filter_b200 import engine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(0)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    pin = lambda a: torch.as_tensor(np.ascontiguousarray(a)).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    peak_fp64 = engine.fp64_peak_tflops()
    if args.workload == "ekf":
        e = make_ekf_inputs(Bf=256 if args.windows == 4096 else args.windows)      # BASELINE size unless --windows says otherwise
        Bf, steps = e["z"].shape[:2]
        units = Bf * steps
        d = {k: dev(e[k]) for k in ("x0", "S0", "u", "z", "sat", "Q")}
        nsat = dev(e["nsat"])
        h = {k: pin(e[k]) for k in ("x0", "S0", "u", "z", "sat")}
        mu, S = torch.empty_like(d["x0"]), torch.empty_like(d["S0"])
        h_out = torch.empty((Bf, steps, 5), dtype=torch.float64).pin_memory()

        def step_device():
            mu.copy_(d["x0"]); S.copy_(d["S0"])
            return engine.ekf_run(engine.EKF_GNSS5, mu, S, d["u"], d["z"], d["sat"], nsat, d["Q"], 100.0, 1.0)[2]

        def step_e2e():
            for k in ("x0", "S0", "u", "z", "sat"):
                d[k].copy_(h[k], non_blocking=True)
            h_out.copy_(step_device(), non_blocking=True)
        h2d = sum(h[k].numel() * 8 for k in h)
        d2h = h_out.numel() * 8
        iters_mean = 1.0
    else:
        l = make_ls_inputs(B=16384 if args.windows == 4096 else args.windows)
        units = l["pr"].shape[0]
        d = {k: dev(l[k]) for k in ("sat", "pr")}
        nsat = dev(l["nsat"])
        h = {k: pin(l[k]) for k in ("sat", "pr")}
        h_out = torch.empty((units, 4), dtype=torch.float64).pin_memory()
        last = {}

        def step_device():
            x, it = engine.ls_solve(d["sat"], d["pr"], nsat)
            last["it"] = it
            return x

        def step_e2e():
            for k in ("sat", "pr"):
                d[k].copy_(h[k], non_blocking=True)
            h_out.copy_(step_device(), non_blocking=True)
        h2d = sum(h[k].numel() * 8 for k in h)
        d2h = h_out.numel() * 8

    def timed(fn):
        for _ in range(args.warmup):
            flush.zero_(); fn()
        torch.cuda.synchronize()
        evs = []
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)

    sampler = ClockSampler(0)
    sampler.start()
    ms_dev = timed(step_device)
    ms_e2e = timed(step_e2e)
    sampler.hold_load(step_device)
    clocks = sampler.stop()
    if args.workload == "ls":
        iters_mean = float(last["it"].to(torch.float64).mean().item())
    ms_kernel = ms_dev / args.steps
    flops = FLOP_PER_ITER[args.workload] * iters_mean * units
    achieved = flops / (ms_kernel * 1e-3) / 1e12
    metric, unit = METRIC[args.workload]
    line = {"metric": metric, "value": units * args.steps / (ms_dev * 1e-3), "unit": unit, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload] if args.windows == 4096 else
                       WORKLOAD_DESC[args.workload].replace("256 filters", "%d filters" % args.windows).replace("16384 independent", "%d independent" % args.windows),
                       "units_per_step": units,
                       "l2": "256 MB buffer written between timed iterations"},
            "stats": {"mean_iterations": iters_mean},
            "e2e": {"value": units * args.steps / (ms_e2e * 1e-3), "unit": unit, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": args.steps,
            "roofline": {"bound": "fp64", "bound_note": "compute-bound on the FP64 FMA pipe; the contract's vocabulary is hbm|tensor: B200 has no FP64 tcgen05 kind and its FP64 mma.sync peak (nlpf_dmma_probe: 36.7 TFLOP/s) equals the vector peak used here", "kernel": "ekf_kernel" if args.workload == "ekf" else "ls_kernel<4>",
                         "achieved": achieved, "peak": peak_fp64, "unit": "TFLOP/s", "frac": achieved / peak_fp64,
                         "traffic": None, "kernel_ms": ms_kernel,
                         "peak_source": "FP64 FMA probe measured in this run",
                         "note": "latency-bound at this batch size (%d units): launch + dependent chains" % units},
            "clocks": clocks}
    if cpu_base is not None:
        line["cpu_baseline"] = cpu_base
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# vehicle chains (BASELINE configs[3]): B runs x W chained windows per step, all on the device
# ---------------------------------------------------------------------------------------------
def _cpu_car_chain(b):
    from oracle import nls
    d, ch = _CPU["d"], _CPU["ch"]
    z = np.zeros(d.nz)
    xp = ch["xhat0"][b].copy()
    it = 0
    for w in range(ch["data"].shape[0]):
        z, info = nls.solve(d, nls.Window(d, U=ch["U"][w, b], data=ch["data"][w, b], xprior=xp), z, _CPU["opts"].for_window(w))
        xp = (ch["PhiDT"] @ z.reshape(d.K, d.nv))[0]
        it += info["iters"]
    return it


def car_opts(args):
    from nlp_filter_b200.desc import SolveOpts
    return SolveOpts(mu0_warm=args.mu0_warm)


def car_config(args, B, W):
    from nlp_filter_b200 import shapes
    d = shapes.car()
    return {"workload": WORKLOAD_DESC["car"], "runs": B, "chained_windows_per_run": W, "windows_per_step": B * W,
            "solver": "log barrier from mu0 = 0.1 in the first window of a run, from mu0_warm = %g in the warm-started windows after it "
                      "(nlpf_opts.mu0_warm; 0 = 0.1 everywhere, as IPOPT's default mu_init)" % args.mu0_warm,
            "l2": "256 MB buffer written between timed iterations; chain inputs (%d MB) exceed L2" % (W * B * d.Ldata * 8 // 2 ** 20)}


def cpu_car_windows_per_sec(runs, windows, cores, opts):
    from nlp_filter_b200 import shapes, synth
    d = shapes.car()
    _CPU["d"], _CPU["ch"], _CPU["opts"] = d, synth.car_chain_batch(d, runs, windows, seed=4000), opts
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        its = pool.map(_cpu_car_chain, range(runs))
        dt = time.perf_counter() - t0
    return runs * windows / dt, float(np.sum(its)) / (runs * windows)


def _traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full summary, or None"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_%s.json" % workload)) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except Exception:
        return None


def run_car(args):
    cores = os.cpu_count() or 1
    cpu_base = None
    if not args.no_cpu_baseline:
        runs, wins = (args.cpu_windows or 2 * cores), 3
        v, it = cpu_car_windows_per_sec(runs, wins, cores, car_opts(args))
        cpu_base = {"value": v, "unit": "windows/s", "cores": cores, "kind": "port",
                    "sample": "%d runs x %d chained windows of the same synthetic workload, NumPy restatement oracle/nls.py, "
                              "%d-process pool, mean %.1f iterations/window" % (runs, wins, cores, it)}
    import torch
    from nlp_filter_b200 import engine, shapes, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(0)
    B, W = args.windows if args.windows != 4096 else 2048, args.chain
    d = shapes.car()
    ch = synth.car_chain_batch(d, B, W, seed=4000)
    prob = engine.WindowProblem(d)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    U_all, data_all, xh0 = dev(ch["U"]), dev(ch["data"]), dev(ch["xhat0"])
    z = torch.zeros((B, d.nz), dtype=torch.float64, device="cuda")
    xp = torch.empty_like(xh0)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    peak_fp64 = engine.fp64_peak_tflops()
    stats = {}
    opts = car_opts(args)

    def step_device():
        z.zero_()
        xp.copy_(xh0)
        stats["it"], stats["bad"] = engine.solve_chain(prob, z, U_all, data_all, xp, ch["PhiDT"], opts=opts)

    hc = engine.HostChains(prob, B, W, nchunk=args.chain_chunks)     # groups of runs: copy of group c+1 under the solve of group c
    hc.load(ch["U"], ch["data"], ch["xhat0"])

    def step_e2e():
        hc.run(ch["PhiDT"], opts)

    def timed(fn):
        for _ in range(args.warmup):
            flush.zero_(); fn()
        torch.cuda.synchronize()
        evs = []
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)

    sampler = ClockSampler(0)
    sampler.start()
    ms_dev = timed(step_device)
    it_mean = float(stats["it"].to(torch.float64).mean().item()) / W
    nbad = int(stats["bad"].sum().item())
    ms_e2e = timed(step_e2e)
    sampler.hold_load(step_device, min_samples=4)
    clocks = sampler.stop()
    units = B * W
    ms_step = ms_dev / args.steps
    achieved = FLOP_PER_ITER["car"] * it_mean * units / (ms_step * 1e-3) / 1e12
    line = {"metric": "solved_windows_per_sec", "value": units * args.steps / (ms_dev * 1e-3), "unit": "windows/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": car_config(args, B, W),
            "stats": {"mean_iterations_per_window": it_mean, "windows_not_converged": nbad},
            "e2e": {"value": units * args.steps / (ms_e2e * 1e-3), "unit": "windows/s",
                    "h2d_bytes_per_step": int(hc.h2d_bytes), "d2h_bytes_per_step": int(hc.d2h_bytes),
                    "ms_per_step": ms_e2e / args.steps,
                    "scope": "pinned host buffers (runs in %d groups, each run-major) -> H2D -> fused-chain launch per group -> D2H of the final estimates; "
                             "a group's copy overlaps the previous group's solve; NumPy -> pinned staging (HostChains.load, once) is outside" % hc.nchunk},
            "gpu_launches": args.steps * (W * 2 if os.environ.get("NLPF_CHAIN_LOOP") else 1),
            "roofline": {"bound": "fp64", "bound_note": "compute-bound on the FP64 FMA pipe; the contract's vocabulary is hbm|tensor: B200 has no FP64 tcgen05 kind and its FP64 mma.sync peak (nlpf_dmma_probe: 36.7 TFLOP/s) equals the vector peak used here", "kernel": "wsolve_kernel<CfgCar> (fused chain: all windows of a run in one warp, one launch per step)",
                         "achieved": achieved, "peak": peak_fp64, "unit": "TFLOP/s",
                         "frac": achieved / peak_fp64, "traffic": _traffic("car") if (B, W) == (2048, 25) else None, "kernel_ms": ms_step,
                         "peak_source": "FP64 FMA probe measured in this run"},
            "clocks": clocks}
    if cpu_base is not None:
        line["cpu_baseline"] = cpu_base
    print(json.dumps(line))
    return 0

# ---------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference arm: the CPU path on all host cores.  CasADi/IPOPT (where the reference's NLP arithmetic
    lives) cannot be installed here, so this is the NumPy restatement in oracle/ (kind "port"); each step solves a
    bounded sample -- the FIRST n windows of the very batch the GPU arm solves (same generator, same seed) -- and
    `config` is the GPU arm's, so the two lines describe one workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    n = args.cpu_windows or {"vdp": 256, "gnss": 64, "2rx": 48, "ekf": 64, "ls": 2048, "car": 0, "rc": 32, "mrx": 16}.get(args.workload, 0)
    vals = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        if args.workload == "car":
            v, it = cpu_car_windows_per_sec(2 * cores, 3, cores, car_opts(args))
            n = 2 * cores * 3
        elif args.workload in ("ekf", "ls"):
            v, it = cpu_aux_per_sec(args.workload, n, cores), 0.0
        else:
            v, it = cpu_path_windows_per_sec(args.workload, n, cores, seed=2000)
        if s >= args.warmup:
            vals.append((v, time.perf_counter() - t0, it))
    value = float(np.mean([v[0] for v in vals]))
    if args.workload in ("vdp", "gnss", "2rx", "rc", "mrx"):
        _, config = shared_config(args, world)
    elif args.workload == "car":
        config = car_config(args, args.windows if args.windows != 4096 else 2048, args.chain)
    else:
        config = {"workload": WORKLOAD_DESC[args.workload]}
    line = {
        "impl": "reference", "metric": METRIC[args.workload][0], "value": value, "unit": METRIC[args.workload][1],
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * n / value, "higher_is_better": True, "scaling": "strong" if args.workload == "2rx" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": METRIC[args.workload][1], "cores": cores, "kind": "port",
                         "sample": "each step: the first %d units of the same synthetic workload (seed 2000), NumPy restatement "
                                   "(oracle/nls.py, oracle/ekf_ls.py) in a %d-process pool; CasADi/IPOPT not installable" % (n, cores)},
        "e2e": {"value": value, "unit": METRIC[args.workload][1], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def shared_config(args, world):
    """The `config` dict both arms print (the CPU arm's bounded sample is described in its cpu_baseline)."""
    strong = args.workload == "2rx"
    B = (1024 // world if args.windows == 4096 else args.windows) if strong else args.windows
    if args.windows == 4096 and args.workload in ("rc", "mrx"):
        B = {"rc": 1024, "mrx": 256}[args.workload]          # windows of these classes are 10-100x the work of the headline ones
    return B, {"workload": WORKLOAD_DESC[args.workload], "windows_per_gpu": B, "global_windows": world * B,
               "parallelism": "windows sharded over %d GPU(s), all-gather of estimates at the end" % world
                              + (" of every step, overlapped with the next step's solve (two buffers)" if world > 1 else ""),
               "l2": "256 MB buffer written between timed iterations (inputs are smaller than L2)"}


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cpu_base = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = args.cpu_windows or {"vdp": 512, "gnss": 128, "2rx": 64, "rc": 32, "mrx": 16}[args.workload]
        v, it = cpu_path_windows_per_sec(args.workload, n, cores)        # before CUDA is initialised (fork)
        cpu_base = {"value": v, "unit": "windows/s", "cores": cores, "kind": "port",
                    "sample": "the first %d windows of the same synthetic batch (seed 2000), NumPy restatement "
                              "oracle/nls.py in a %d-process pool, mean %.1f iterations/window" % (n, cores, it)}

    import torch
    import torch.distributed as dist
    from nlp_filter_b200 import engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    strong = args.workload == "2rx"          # BASELINE configs[4]: 1024 scenarios in total, sharded over the GPUs
    B, config = shared_config(args, world)
    desc = make_problem(args.workload)
    prob = engine.WindowProblem(desc)
    bt = make_batch(desc, args.workload, B, seed=2000 + rank)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    z0, data = dev(bt["z0"]), dev(bt["data"])
    U = dev(bt["U"]) if bt.get("U") is not None else None
    xprior = dev(bt["xprior"]) if bt.get("xprior") is not None else None
    z = torch.empty_like(z0)
    cost = torch.empty((B,), dtype=torch.float64, device="cuda")
    iters = torch.empty((B,), dtype=torch.int32, device="cuda")
    status = torch.empty((B,), dtype=torch.int32, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")     # 256 MB > 126 MB L2
    peak_fp64 = engine.fp64_peak_tflops()

    # Multi-GPU: the gather of step s (NCCL's own stream) overlaps the solve of step s+1 -- sharding.GatherPipe: two
    # gather buffers alternate, a buffer is reused only after its gather has completed, and the timed region ends only
    # after the last gather has (pipe.finish()).  The end-to-end leg also brings every gathered batch to the host.
    from nlp_filter_b200 import sharding
    pipes = {"dev": sharding.GatherPipe(B, (desc.nz,)), "e2e": sharding.GatherPipe(B, (desc.nz,), to_host=True)}
    zbuf = [z, torch.empty_like(z0)] if world > 1 else [z]

    def step_device():
        pipe = pipes["dev"]
        i = pipe.slot()
        zi = zbuf[i]
        zi.copy_(z0)
        prob.solve(zi, U, data, xprior, None, None, out=(cost, iters, status))
        pipe.gather(i, zi)

    hb = engine.HostBatch(prob, B, nchunk=(args.e2e_chunks if B >= 1024 else 1) if args.e2e_chunks > 0 else None)
    hb.load(bt["z0"], bt["U"], bt["data"], bt.get("xprior"))
    hb.run()
    torch.cuda.synchronize()
    e2e_graph = False
    if not args.no_e2e_graph:
        try:                                    # one launch per step instead of ~12 per chunk issued from Python
            hb.capture()
            e2e_graph = True
        except Exception as exc:                # (a launch-overhead optimisation: the plain path computes the same thing)
            sys.stderr.write("bench.py: CUDA-graph capture of the end-to-end step failed (%s); issuing it call by call\n" % exc)
            hb._graph = None

    def step_e2e():
        pipe = pipes["e2e"]
        i = pipe.slot()
        hb.run()
        if world > 1:
            pipe.gather(i, hb.device_z())

    def timed(step_fn, pipe):
        """ONE CUDA-event pair around all K steps, L2 flushes included; the flushes alone are timed the same way and
        subtracted (they are scaffolding of the measurement, not part of the path).  Nothing of a step -- solve,
        collective, copies -- can hide in an untimed gap (round-1 review: per-step events with untimed flushes
        between them let the asynchronous gather finish outside the clock)."""
        for _ in range(args.warmup):
            flush.zero_()
            step_fn()
        pipe.finish()
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            flush.zero_()
        f1.record()
        torch.cuda.synchronize()
        ms_flush = f0.elapsed_time(f1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s_ in range(args.steps):
            flush.zero_()                       # evict the inputs from L2 between timed iterations
            step_fn()
        pipe.finish()                           # every gather (and host copy) of the timed steps completes inside the region
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms_region = e0.elapsed_time(e1)
        t = torch.tensor([ms_region - ms_flush, ms_region, ms_flush], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, ms_dev_region, ms_flush = timed(step_device, pipes["dev"])
    it_mean = float(iters.to(torch.float64).mean().item())
    st = status.cpu().numpy()
    ms_e2e, ms_e2e_region, _ = timed(step_e2e, pipes["e2e"])
    # kernel-only duration (solve launch alone -- for the linear-measurement classes that is the start-cost, ranking and
    # solve kernels of one nlpf_solve call; CUDA events on the launching stream)
    kev = []
    for _ in range(max(3, args.steps)):
        flush.zero_()
        z.copy_(z0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        prob.solve(z, U, data, xprior, None, None, out=(cost, iters, status))
        e1.record()
        kev.append((e0, e1))
    torch.cuda.synchronize()
    ms_kernel = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    if rank == 0 and world == 1:
        def _load():
            z.copy_(z0)
            prob.solve(z, U, data, xprior, None, None, out=(cost, iters, status))
        sampler.hold_load(_load)
    clocks = sampler.stop() if rank == 0 else None
    # per-phase split of the fused kernel: clock counters inside the warp-per-window kernel (one extra, untimed solve)
    phases = None
    if rank == 0 and args.workload in FLOP_CHOL_PER_ITER:
        prob.profile(True)
        z.copy_(z0)
        prob.solve(z, U, data, xprior, None, None, out=(cost, iters, status))
        pc = prob.profile(False)
        if pc["window_cycles"] > 0:
            f_it, f_ch = FLOP_PER_ITER[args.workload], FLOP_CHOL_PER_ITER[args.workload]
            sh_e, sh_l = pc["eval_cycles"] / pc["window_cycles"], pc["ldl_cycles"] / pc["window_cycles"]
            fl_e, fl_l = (f_it - f_ch) * pc["evals"], f_ch * pc["factorisations"]
            ach_e = fl_e / (ms_kernel * 1e-3 * sh_e) / 1e12 if sh_e > 0 else None
            ach_l = fl_l / (ms_kernel * 1e-3 * sh_l) / 1e12 if sh_l > 0 else None
            phases = {"method": "clock64 around the two phases inside the kernel (nlpf_problem_profile), summed over warps; a "
                                "phase's time = kernel time x its share of the warp-cycles spent on windows; flops by the "
                                "SURVEY 8(d) convention: (flop/iteration - factorisation flop) per evaluation, factorisation "
                                "flop per factorisation, counted per call (rejected steps included)",
                      "eval": {"share": sh_e, "calls": pc["evals"], "achieved": ach_e, "frac": ach_e / peak_fp64 if ach_e else None},
                      "ldl": {"share": sh_l, "calls": pc["factorisations"], "achieved": ach_l, "frac": ach_l / peak_fp64 if ach_l else None},
                      "other_share": 1.0 - sh_e - sh_l}

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
        # run-dependent numbers live in "stats", not in "config": both arms print the SAME config dict
        stats = {"mean_iterations_per_window": it_mean_all, "converged_fraction": float(np.mean(st <= 1))}
        gather_d2h = world * B * desc.nz * 8 if world > 1 else 0
        line = {
            "metric": "solved_windows_per_sec", "value": value, "unit": "windows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "stats": stats,
            "timing": {"method": "one CUDA-event pair around all timed steps (L2 flushes inside), max over ranks; the flushes "
                                 "alone, timed the same way, are subtracted",
                       "region_ms": ms_dev_region, "flush_ms": ms_flush, "e2e_region_ms": ms_e2e_region},
            "e2e": {"value": e2e, "unit": "windows/s", "h2d_bytes_per_step": hb.h2d_bytes, "d2h_bytes_per_step": hb.d2h_bytes + gather_d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "scope": "pinned host buffers -> H2D -> solve -> D2H of estimates, cost, iterations, status"
                             + (" -> all-gather -> D2H of the gathered estimates" if world > 1 else "")
                             + "; the NumPy -> pinned staging (HostBatch.load, once) and the NumPy views of the results are outside; %d chunks, %s"
                             % (hb.nchunk, "replayed as one CUDA graph per step (HostBatch.capture)" if e2e_graph else "issued call by call")},
            "gpu_launches": args.steps * (3 if (args.workload == "vdp" and B > 1776) else 1),
            "roofline": {"bound": "fp64", "bound_note": "compute-bound on the FP64 FMA pipe; the contract's vocabulary is hbm|tensor: B200 has no FP64 tcgen05 kind and its FP64 mma.sync peak (nlpf_dmma_probe: 36.7 TFLOP/s) equals the vector peak used here", "kernel": {"vdp": "wsolve_kernel<CfgVdp>", "gnss": "wsolve_kernel<CfgGnss>", "2rx": "solve_kernel_prb", "rc": "solve_kernel_prb", "mrx": "solve_kernel_prb"}[args.workload], "achieved": achieved, "peak": peak_fp64,
                         "unit": "TFLOP/s", "frac": achieved / peak_fp64 if peak_fp64 > 0 else None,
                         "traffic": traffic, "kernel_ms": ms_kernel,
                         "peak_source": "FP64 FMA probe (nlpf_fp64_probe) measured in this run; MEASURED_PEAKS.json "
                                        "has no FP64 figure",
                         "flop_per_iteration": FLOP_PER_ITER[args.workload],
                         "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s"}},
            "clocks": clocks,
        }
        if phases is not None:
            line["roofline"]["phases"] = phases
            line["roofline"]["eval_frac"] = phases["eval"]["frac"]
            line["roofline"]["ldl_frac"] = phases["ldl"]["frac"]
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------
# kernels (1)+(2) alone: nlpf_assemble (residuals, Jacobians, J^T W J, J^T W r; H written to HBM)
# ---------------------------------------------------------------------------------------------
def run_assemble(args):
    import torch
    from nlp_filter_b200 import engine, shapes, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(0)
    wl = args.workload.split("-")[1]
    B = args.windows
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    if wl == "car":
        d = shapes.car()
        ch = synth.car_chain_batch(d, B, 1, seed=4000)
        bt = dict(z0=np.zeros((B, d.nz)), U=ch["U"][0], data=ch["data"][0], xprior=ch["xhat0"])
    else:
        d = make_problem(wl)
        bt = make_batch(d, wl, B, seed=2000)
    prob = engine.WindowProblem(d)
    z, U, data = dev(bt["z0"]), dev(bt["U"]), dev(bt["data"])
    xp = dev(bt["xprior"]) if bt.get("xprior") is not None else None
    prob.solve(z, U, data, xp)                     # evaluate at the converged estimates (a realistic point, inside the bounds)
    torch.cuda.synchronize()
    nz = d.nz
    H = torch.empty((B, nz, nz), dtype=torch.float64, device="cuda")
    g = torch.empty((B, nz), dtype=torch.float64, device="cuda")
    cost = torch.empty((B, 2), dtype=torch.float64, device="cuda")
    import ctypes
    from nlp_filter_b200 import _capi
    ptr = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())

    def step():
        _capi.check(prob._lib.nlpf_assemble(prob._h, B, ptr(z), ptr(U), ptr(data), ptr(xp), None, 0.0, ptr(H), ptr(g), ptr(cost),
                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    peak_fp64 = engine.fp64_peak_tflops()
    sampler = ClockSampler(0)
    sampler.start()
    for _ in range(args.warmup):
        flush.zero_(); step()
    torch.cuda.synchronize()
    evs = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    sampler.hold_load(step)
    clocks = sampler.stop()
    ms_k = ms / args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_w = BYTES_PER_WINDOW[wl] - 8.0 * (nz + 3) + 8.0 * (nz * nz + nz + 2)      # inputs once; H (nz^2), g, {J, merit} out
    flop_w = FLOP_PER_ITER[wl] - FLOP_CHOL_PER_ITER[wl]
    hbm_ach = bytes_w * B / (ms_k * 1e-3) / 1e9
    fl_ach = flop_w * B / (ms_k * 1e-3) / 1e12
    line = {"metric": "assembled_windows_per_sec", "value": B * args.steps / (ms * 1e-3), "unit": "windows/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_k, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "kernels (1)+(2) alone (nlpf_assemble: residuals, analytic Jacobians, J^T W J + exact second-order "
                                   "dynamics term, J^T W r; H written to HBM as nz x nz) on " + WORKLOAD_DESC["car" if wl == "car" else wl],
                       "windows_per_step": B, "l2": "256 MB buffer written between timed iterations"},
            "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "kernel": "ws_assemble_kernel<Cfg%s>" % wl.capitalize(), "achieved": hbm_ach, "peak": hbm_peak,
                         "unit": "GB/s", "frac": hbm_ach / hbm_peak, "traffic": None, "kernel_ms": ms_k,
                         "bytes_per_window": bytes_w, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s",
                         "fp64": {"achieved": fl_ach, "peak": peak_fp64, "unit": "TFLOP/s", "frac": fl_ach / peak_fp64,
                                  "flop_per_window": flop_w, "peak_source": "FP64 FMA probe measured in this run"},
                         "note": "arithmetic intensity %.1f flop/B against a ridge of %.1f: both fractions reported (SURVEY 8d)"
                                 % (flop_w / bytes_w, peak_fp64 * 1e3 / hbm_peak)},
            "clocks": clocks}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# long horizon (row A9): one window = 200 epochs = 40 collocation segments sharing their end nodes
# ---------------------------------------------------------------------------------------------
def _cpu_long(b):
    from oracle import multiseg as oms
    seg, S, bt = _CPU["seg"], _CPU["S"], _CPU["bt"]
    return oms.solve(seg, S, bt["z0"][b], bt["U"][b], bt["data"][b])[1]["iters"]


def run_long(args):
    from nlp_filter_b200 import shapes, synth
    S = 40
    B = 512 if args.windows == 4096 else args.windows
    seg = shapes.vdp_segment()
    cores = os.cpu_count() or 1
    cpu_base = None
    if not args.no_cpu_baseline:
        n = args.cpu_windows or 2 * cores
        _CPU["seg"], _CPU["S"], _CPU["bt"] = seg, S, synth.vdp_multiseg_batch(seg, S, n, seed=6000)
        with mp.get_context("fork").Pool(cores) as pool:
            t0 = time.perf_counter()
            its = pool.map(_cpu_long, range(n))
            dt = time.perf_counter() - t0
        cpu_base = {"value": n / dt, "unit": "windows/s", "cores": cores, "kind": "port",
                    "sample": "the first %d windows of the same synthetic batch, dense NumPy restatement oracle/multiseg.py "
                              "(402 x 402 Cholesky per iteration) in a %d-process pool, mean %.1f iterations" % (n, cores, float(np.mean(its)))}
    import torch
    from nlp_filter_b200 import engine, multiseg
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(0)
    bt = synth.vdp_multiseg_batch(seg, S, B, seed=6000)
    ms = multiseg.MultiSegmentProblem(seg, S)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    z0, U, data = dev(bt["z0"]), dev(bt["U"]), dev(bt["data"])
    h_in = [torch.as_tensor(np.ascontiguousarray(bt[k])).pin_memory() for k in ("z0", "U", "data")]
    h_out = torch.empty((B, ms.nzt), dtype=torch.float64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    peak_fp64 = engine.fp64_peak_tflops()
    variant = multiseg.CYCLIC if os.environ.get("NLPF_LONG_CR") else multiseg.RICCATI
    last = {}

    def step_device():
        z = z0.clone()
        last["out"] = ms.solve(z, U, data, variant=variant)

    def step_e2e():
        z0.copy_(h_in[0], non_blocking=True); U.copy_(h_in[1], non_blocking=True); data.copy_(h_in[2], non_blocking=True)
        step_device()
        h_out.copy_(last["out"][0], non_blocking=True)

    def timed(fn):
        for _ in range(args.warmup):
            flush.zero_(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()                        # (inputs are re-staged by every step; the solve iterates on device-resident blocks)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    sampler = ClockSampler(0)
    sampler.start()
    ms_dev = timed(step_device)
    z, J, it, st = last["out"]
    it_mean = float(it.to(torch.float64).mean().item())
    conv = float((st <= 1).to(torch.float64).mean().item())
    ms_e2e = timed(step_e2e)
    sampler.hold_load(step_device, min_samples=4)
    clocks = sampler.stop()
    K, n, E, M = seg.K, seg.n, seg.E, seg.M
    bs = (K - 1) * n
    f_seg = 2 * E * K * n + 62 * M + E * K * (K + 1) * n * n + 2 * E * K * n + 2 * K * n * n + 4 * K * K * n * n + 2 * K * n ** 3 + 4 * K * K * n + 2 * K * n * n
    f_blk = bs ** 3 / 3.0 + 4.0 * bs ** 3 + 6.0 * bs * bs      # Cholesky, S^-1 [L^T | y] (two triangular sweeps of bs columns), L (S^-1 L^T), sweeps
    flop_iter = S * f_seg + (S + 1) * f_blk
    ms_step = ms_dev / args.steps
    achieved = flop_iter * it_mean * B / (ms_step * 1e-3) / 1e12
    line = {"metric": "solved_windows_per_sec", "value": B * args.steps / (ms_dev * 1e-3), "unit": "windows/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "van der Pol MHE over a 200-epoch horizon: ONE window = 40 collocation segments (N_s = 5, 5 epochs each) "
                                   "sharing their end nodes, 402 unknowns, block-tridiagonal normal equations (41 blocks of 10) solved by "
                                   + ("block cyclic reduction" if variant == multiseg.CYCLIC else "the Riccati recursion"),
                       "windows_per_step": B, "epochs_per_step": B * S * E,
                       "l2": "per-iteration working set (%d MB of segment matrices and blocks) is re-written every iteration" % ((B * S * (K * n) ** 2 * 8 * 2 + B * (S + 1) * bs * bs * 16) >> 20)},
            "stats": {"mean_iterations_per_window": it_mean, "converged_fraction": conv},
            "e2e": {"value": B * args.steps / (ms_e2e * 1e-3), "unit": "windows/s", "h2d_bytes_per_step": int(sum(t.numel() * 8 for t in h_in)),
                    "d2h_bytes_per_step": int(h_out.numel() * 8), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(args.steps * (1 + it_mean * 1.3) * 4),
            "roofline": {"bound": "fp64", "kernel": "assemble_kernel_gen + ms_blocks_kernel + blocktri_%s_kernel per iteration" % ("cr" if variant == multiseg.CYCLIC else "riccati"),
                         "achieved": achieved, "peak": peak_fp64, "unit": "TFLOP/s", "frac": achieved / peak_fp64, "traffic": None,
                         "kernel_ms": ms_step, "flop_per_iteration": flop_iter, "peak_source": "FP64 FMA probe measured in this run",
                         "note": "launch-per-phase iteration driven from the host (4 kernels + element-wise bookkeeping per sweep): latency-bound at this batch size"},
            "clocks": clocks}
    if cpu_base is not None:
        line["cpu_baseline"] = cpu_base
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vdp", choices=["vdp", "gnss", "2rx", "car", "ekf", "ls",
                                                          "assemble-vdp", "assemble-gnss", "assemble-car", "vdp-long", "rc", "mrx"])
    ap.add_argument("--mu0-warm", type=float, default=1e-6, help="car workload: barrier parameter at which the warm-started windows 2.. of a run start (0: mu0)")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="chunks of the end-to-end leg (0: HostBatch's own choice)")
    ap.add_argument("--no-e2e-graph", action="store_true", help="issue the end-to-end step call by call instead of replaying a CUDA graph")
    ap.add_argument("--chain-chunks", type=int, default=8, help="car workload, end-to-end leg: groups of runs whose copies overlap the previous group's solve")
    ap.add_argument("--chain", type=int, default=200, help="chained windows per run (car workload; BASELINE configs[3]: horizon 200)")
    ap.add_argument("--windows", type=int, default=4096, help="windows per GPU per step")
    ap.add_argument("--cpu-windows", type=int, default=0, help="size of the CPU-path sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.workload in ("ekf", "ls"):
        return run_aux(args)
    if args.workload == "car":
        return run_car(args)
    if args.workload.startswith("assemble-"):
        return run_assemble(args)
    if args.workload == "vdp-long":
        return run_long(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
