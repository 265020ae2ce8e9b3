"""Diagnostics for the 2048 x 200 vehicle chains: which windows end at the iteration limit, what the oracle does there."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nlp_filter_b200 import engine, shapes, synth
from oracle import nls
d = shapes.car()
B, W = 2048, 200
ch = synth.car_chain_batch(d, B, W, seed=4000)
prob = engine.WindowProblem(d)
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
z = torch.zeros((B, d.nz), dtype=torch.float64, device="cuda")
xp = dev(ch["xhat0"])
it_sum, bad, Z = engine.solve_chain(prob, z, dev(ch["U"]), dev(ch["data"]), xp, ch["PhiDT"], keep=True)
torch.cuda.synchronize()
Z = Z.cpu().numpy(); bad = bad.cpu().numpy()
print("bad windows", bad.sum(), "runs", np.nonzero(bad)[0].tolist())
PhiDT = ch["PhiDT"]
V0 = Z.reshape(W, B, d.K, d.nv)[:, :, 0, :]
err = np.linalg.norm(V0[:, :, :2] - ch["truth"][:, :, :2], axis=2)
print("pos err median %.3f max(after 20) %.3f" % (np.median(err), np.max(err[20:])))
for b in np.nonzero(bad)[0][:6]:
    zs = np.concatenate([np.zeros((1, d.nz)), Z[:-1, b]], 0)
    xps = np.concatenate([ch["xhat0"][b][None], np.einsum("k,wkn->wn", PhiDT[0], zs[1:].reshape(W - 1, d.K, d.nv))], 0)
    zz, J, iters, status = engine.solve_windows_host(prob, zs, ch["U"][:, b], ch["data"][:, b], xps)
    ws = np.nonzero(status > 1)[0]
    rel = np.max(np.abs(zz - Z[:, b]) / (1 + np.abs(zz)), axis=1)
    print("run", b, "bad", bad[b], "replay status>1 at", ws.tolist(), "status", status[ws].tolist(), "iters", iters[ws].tolist())
    print("   replay rel diff: max over good windows %.2e at w=%d; at bad %s" % (np.max(np.where(status <= 1, rel, 0)), int(np.argmax(np.where(status <= 1, rel, 0))), rel[ws]))
    for w in ws[:2]:
        zprev = Z[w - 1, b] if w > 0 else np.zeros(d.nz)
        xpo = (PhiDT @ zprev.reshape(d.K, d.nv))[0] if w > 0 else ch["xhat0"][b]
        zo, info = nls.solve(d, nls.Window(d, U=ch["U"][w, b], data=ch["data"][w, b], xprior=xpo), zprev)
        Vg, Vo = Z[w, b].reshape(d.K, d.nv), zo.reshape(d.K, d.nv)
        print("   w=%d oracle %s  |dpos| %.3e rel %.3e  J gpu %.9g" % (w, info, np.max(np.abs(Vg[:, [0, 1, 8]] - Vo[:, [0, 1, 8]])), np.max(np.abs(Vg - Vo) / (1 + np.abs(Vo))), J[w]))
        print("     state at node 0 gpu", np.round(Vg[0], 4), " truth", np.round(ch["truth"][w, b], 4))
        w2 = int(np.argmax(np.where(status <= 1, rel, 0)))
    w2 = int(np.argmax(np.where(status <= 1, rel, 0)))
    Va, Vb = zz[w2].reshape(d.K, d.nv), Z[w2, b].reshape(d.K, d.nv)
    print("   worst good window w=%d status %d iters %d: per-component max |diff| over nodes:" % (w2, status[w2], iters[w2]), np.max(np.abs(Va - Vb), 0))
    print("      xprior diff", np.max(np.abs(xps[w2] - (PhiDT @ Z[w2 - 1, b].reshape(d.K, d.nv))[0])), "J replay %.12g" % J[w2])
    zprev = Z[w2 - 1, b]; xpo = (PhiDT @ zprev.reshape(d.K, d.nv))[0]
    zo, info = nls.solve(d, nls.Window(d, U=ch["U"][w2, b], data=ch["data"][w2, b], xprior=xpo), zprev)
    Vo = zo.reshape(d.K, d.nv)
    print("      oracle", info, " |oracle - chain| per comp", np.max(np.abs(Vo - Vb), 0), " |oracle - replay|", np.max(np.abs(Vo - Va), 0))
