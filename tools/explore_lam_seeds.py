"""Damping schedules on several synthetic van der Pol batches (different seeds): device time, iterations.
(tools/ = development scripts, not part of the product or the tests)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from nlp_filter_b200 import engine, shapes, synth, desc as D

SETS = {"round1": dict(lam0=1e-3, lam_jump=1e-30, lam_jump_indef=1e-30, lam_dec_first=10.0),
        "default": dict(),
        "ji.03": dict(lam_jump_indef=0.03), "ji.2": dict(lam_jump_indef=0.2), "ji.3": dict(lam_jump_indef=0.3), "ji.5": dict(lam_jump_indef=0.5),
        "ji1": dict(lam_jump_indef=1.0), "ji.3 df100": dict(lam_jump_indef=0.3, lam_dec_first=100.0),
        "ji.3 lj1e-3": dict(lam_jump_indef=0.3, lam_jump=1e-3), "ji.1 lj1e-2": dict(lam_jump=1e-2)}
d = shapes.vdp_mhe()
prob = engine.WindowProblem(d)
tot = {k: 0.0 for k in SETS}
for seed in (2000, 1, 2, 3, 4, 5):
    bt = synth.vdp_mhe_batch(d, 4096, seed=seed)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    z0, U, data = dev(bt["z0"]), dev(bt["U"]), dev(bt["data"])
    for name, kw in SETS.items():
        o = D.SolveOpts(**kw)
        best = 1e9
        for rep in range(4):
            z = z0.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = prob.solve(z, U, data, opts=o)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        it = out[2].cpu().numpy(); st = out[3].cpu().numpy()
        tot[name] += best
        print("seed %d %-14s %.3f ms  iters mean %.2f max %d  status %s" % (seed, name, best, it.mean(), it.max(), np.bincount(st).tolist()), flush=True)
print({k: round(v, 3) for k, v in tot.items()})
