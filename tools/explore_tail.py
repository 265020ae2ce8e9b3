"""How much of the van der Pol kernel time is the tail?  Solve the batch, then replace the windows that needed more than
CUT iterations by copies of ordinary ones and time again (tools/ = development scripts)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from nlp_filter_b200 import engine, shapes, synth, desc as D

d = shapes.vdp_mhe()
prob = engine.WindowProblem(d)
bt = synth.vdp_mhe_batch(d, 4096)
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()

def run(z0, U, data, label):
    z0d, Ud, dd = dev(z0), dev(U), dev(data)
    best = 1e9
    for rep in range(5):
        z = z0d.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = prob.solve(z, Ud, dd); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    it = out[2].cpu().numpy()
    print("%-40s %.3f ms  iters mean %.2f max %d  hist %s" % (label, best, it.mean(), it.max(), np.bincount(it).tolist()), flush=True)
    return it

it = run(bt["z0"], bt["U"], bt["data"], "full batch")
for cut in (14, 11, 9, 8, 7):
    z0, U, data = bt["z0"].copy(), bt["U"].copy(), bt["data"].copy()
    long_ = np.nonzero(it > cut)[0]
    ordinary = np.nonzero(it == 7)[0]
    src = ordinary[np.arange(len(long_)) % len(ordinary)]
    z0[long_], U[long_], data[long_] = z0[src], U[src], data[src]
    run(z0, U, data, "windows with > %d iterations replaced (%d)" % (cut, len(long_)))
# all windows identical (perfect balance): the throughput-bound time at this mean work
src = np.nonzero(it == 7)[0][:1]
idx = np.repeat(src, 4096)
run(bt["z0"][idx], bt["U"][idx], bt["data"][idx], "4096 copies of one 7-iteration window")
