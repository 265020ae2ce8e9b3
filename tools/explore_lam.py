"""Damping-schedule sweep on the GPU: iterations, evaluations, factorisations and device time per batch
(tools/ = development scripts, not part of the product or the tests)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from nlp_filter_b200 import engine, shapes, synth, desc as D

SETS = [dict(), dict(lam0=1e-3, lam_jump=1e-30, lam_jump_indef=1e-30, lam_dec_first=10.0)]       # defaults; the round-1 schedule
for ji in (1e-4, 1e-3, 1e-2, 3e-2, 1e-1, 1.0):
    for df in (10.0, 1000.0):
        SETS.append(dict(lam_jump_indef=ji, lam_dec_first=df))
for l0 in (1e-6, 1e-3):
    SETS.append(dict(lam0=l0, lam_jump_indef=1e-2))
which = sys.argv[1:] or ["vdp", "gnss", "rc", "2rx"]
table = {"vdp": (shapes.vdp_mhe, synth.vdp_mhe_batch, 4096), "gnss": (shapes.gnss, synth.gnss_batch, 4096),
         "rc": (shapes.rc, synth.rc_batch, 1024), "2rx": (shapes.two_receiver, synth.two_receiver_batch, 1024)}
for name in which:
    mk, syn, B = table[name]
    d = mk()
    bt = syn(d, B)
    prob = engine.WindowProblem(d)
    dev = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a)).cuda()
    z0, U, data, xp = dev(bt["z0"]), dev(bt.get("U")), dev(bt["data"]), dev(bt.get("xprior"))
    sets = SETS if name in ("vdp", "rc") else SETS[:2] + [dict(lam_dec_first=1000.0)]
    for kw in sets:
        o = D.SolveOpts(**kw)
        best = 1e9
        for rep in range(4):
            z = z0.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = prob.solve(z, U, data, xprior=xp, opts=o)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        it = out[2].cpu().numpy(); st = out[3].cpu().numpy()
        print("%s %-70s %.3f ms  iters mean %.2f max %d  status %s" % (name, kw, best, it.mean(), it.max(), np.bincount(st).tolist()), flush=True)
    prob.close()
