"""World-size-2 test (gloo, CPU) of the multi-rank host logic: contiguous sharding of the window
batch and the single end-of-step gather.  The per-rank solve is the CPU oracle here (the CUDA
path needs a GPU); what is under test is that shards + gather reproduce the single-rank batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nlp_filter_b200 import shapes, sharding, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import nls
    d = shapes.vdp_mhe()
    bt = synth.vdp_mhe_batch(d, B, seed=77)                 # every rank builds the same global batch ...
    lo, hi = sharding.shard_bounds(B, rank, world)          # ... and solves only its shard
    local = np.zeros((hi - lo, d.nz))
    for i in range(lo, hi):
        local[i - lo], _ = nls.solve(d, nls.Window(d, U=bt["U"][i], data=bt["data"][i]), bt["z0"][i])
    full = sharding.gather_estimates(torch.as_tensor(local), B)
    if rank == 0:
        np.save(out_path, full.numpy())
    if B % world == 0:
        # the pipelined form bench.py uses: three "steps" whose gathers overlap the next step; every gathered batch
        # (device buffer and host copy) must equal the plain gather of that step's estimates
        pipe = sharding.GatherPipe(hi - lo, (d.nz,), device="cpu", to_host=True)
        for step in range(3):
            i = pipe.slot()
            pipe.gather(i, torch.as_tensor(local) + float(step))
        pipe.finish()
        for step, i in ((1, 1), (2, 0)):               # slots alternate 0, 1, 0: slot 0 now holds step 2, slot 1 step 1
            dev_buf, host_buf = pipe.last(i)
            assert torch.equal(dev_buf, full + float(step)) and torch.equal(host_buf, full + float(step))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [6, 5])
def test_two_rank_shards_reproduce_single_rank_batch(tmp_path, B):
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), B, out), nprocs=2, join=True)
    got = np.load(out)
    from oracle import nls
    d = shapes.vdp_mhe()
    bt = synth.vdp_mhe_batch(d, B, seed=77)
    ref = np.array([nls.solve(d, nls.Window(d, U=bt["U"][i], data=bt["data"][i]), bt["z0"][i])[0] for i in range(B)])
    assert got.shape == ref.shape and np.array_equal(got, ref)


def test_shard_bounds_cover_batch_exactly():
    for B in (1, 7, 4096, 1024):
        for world in (1, 2, 4, 8):
            cuts = [sharding.shard_bounds(B, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            assert max(c[1] - c[0] for c in cuts) - min(c[1] - c[0] for c in cuts) <= 1
