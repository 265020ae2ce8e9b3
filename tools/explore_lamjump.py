"""lam_jump (damping after a step that did not decrease the merit) on every class: device time and iterations
(tools/ = development scripts)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from nlp_filter_b200 import engine, shapes, synth, desc as D

def timed(fn, reps=4):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out

dev = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a)).cuda()
for lj in (1e-4, 1e-3, 1e-2, 1e-1):
    line = ["lam_jump %g:" % lj]
    for name, mk, syn, B in (("vdp", shapes.vdp_mhe, synth.vdp_mhe_batch, 4096), ("gnss", shapes.gnss, synth.gnss_batch, 4096),
                             ("rc", shapes.rc, synth.rc_batch, 1024), ("2rx", shapes.two_receiver, synth.two_receiver_batch, 1024)):
        d = mk(); bt = syn(d, B); prob = engine.WindowProblem(d)
        z0, U, data, xp = dev(bt["z0"]), dev(bt.get("U")), dev(bt["data"]), dev(bt.get("xprior"))
        o = D.SolveOpts(lam_jump=lj)
        ms, out = timed(lambda: prob.solve(z0.clone(), U, data, xprior=xp, opts=o))
        it = out[2].cpu().numpy(); st = out[3].cpu().numpy()
        line.append("%s %.3f ms it %.2f max %d bad %d |" % (name, ms, it.mean(), it.max(), int((st > 1).sum())))
        prob.close()
    d = shapes.car(); B, W = 2048, 25
    ch = synth.car_chain_batch(d, B, W, seed=4000)
    prob = engine.WindowProblem(d)
    U_all, data_all, xh0 = dev(ch["U"]), dev(ch["data"]), dev(ch["xhat0"])
    o = D.SolveOpts(lam_jump=lj, mu0_warm=1e-6)
    def run():
        z = torch.zeros((B, d.nz), dtype=torch.float64, device="cuda"); xp = xh0.clone()
        return engine.solve_chain(prob, z, U_all, data_all, xp, ch["PhiDT"], opts=o)
    ms, out = timed(run, 2)
    line.append("car x25 %.2f ms it %.2f bad %d" % (ms, float(out[0].sum()) / (B * W), int(out[1].sum())))
    prob.close()
    print(" ".join(line), flush=True)
