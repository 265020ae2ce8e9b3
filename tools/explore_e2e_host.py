"""Host-side cost of one end-to-end step (HostBatch.run) against its device time, plain and as a CUDA graph
(tools/ = development scripts)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from nlp_filter_b200 import engine, shapes, synth

d = shapes.vdp_mhe()
bt = synth.vdp_mhe_batch(d, 4096)
prob = engine.WindowProblem(d)
ref = None
for graph in (False, True):
    for nchunk in (1, 2, 3, 4, 6, 8):
        hb = engine.HostBatch(prob, 4096, nchunk=nchunk)
        hb.load(bt["z0"], bt["U"], bt["data"])
        for _ in range(3):
            hb.run()
        if graph:
            hb.capture()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            hb.run()
        t_host = (time.perf_counter() - t0) / 50
        torch.cuda.synchronize()
        t_all = (time.perf_counter() - t0) / 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(10):
            torch.cuda.synchronize()
            e0.record(); hb.run(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        z = hb.results()[0]
        if ref is None: ref = z.copy()
        print("graph %d chunks %d: host issue %.3f ms per step, back-to-back %.3f ms per step, one step alone (events) %.3f ms, equal to first: %s" % (graph, nchunk, t_host * 1e3, t_all * 1e3, best, np.array_equal(z, ref)), flush=True)
