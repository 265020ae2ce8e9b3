"""Sharding of a batch of independent windows over ranks (one process per GPU) and the single
end-of-step collective.  No data-path collective: every window is solved entirely on one GPU."""
import torch
import torch.distributed as dist


def shard_bounds(B, rank, world):
    """Contiguous split of B units over `world` ranks; the first B % world ranks take one more."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_estimates(local, B_total, world=None):
    """all-gather of per-rank results (rows = windows) into the global (B_total, ...) tensor,
    in rank order; ranks may hold unequal shard sizes."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = world or dist.get_world_size()
    sizes = [shard_bounds(B_total, r, world)[1] - shard_bounds(B_total, r, world)[0] for r in range(world)]
    if len(set(sizes)) == 1:
        out = torch.empty((B_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)


class GatherPipe(object):
    """The end-of-step collective of a multi-rank run, pipelined: the all-gather of step s (on the backend's own
    stream) overlaps the solve of step s + 1.  Two gather buffers alternate; a buffer is reused only after its
    gather has completed, and finish() waits for everything still in flight.  With to_host=True every gathered batch
    is also copied to pinned host memory (the end-to-end leg of bench.py).  world == 1: every call is a no-op.

    Usage per step:   i = pipe.slot();  ... solve into a buffer ...;  pipe.gather(i, estimates)
    and once after the last step:   pipe.finish()."""

    def __init__(self, rows_per_rank, row_shape, dtype=torch.float64, device="cuda", to_host=False):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.k = 0
        self.pending = [None, None]
        if self.world > 1:
            shape = (self.world * rows_per_rank,) + tuple(row_shape)
            self.gb = [torch.empty(shape, dtype=dtype, device=device) for _ in range(2)]
            pin = to_host and torch.device(device).type == "cuda"
            self.hb = [torch.empty(shape, dtype=dtype).pin_memory() if pin else torch.empty(shape, dtype=dtype)
                       for _ in range(2)] if to_host else None
        else:
            self.gb, self.hb = [None, None], None

    def _complete(self, i):
        if self.pending[i] is None:
            return
        self.pending[i].wait()                     # the current stream waits for the collective
        if self.hb is not None:                    # gathered estimates -> host memory
            self.hb[i].copy_(self.gb[i], non_blocking=True)
        self.pending[i] = None

    def slot(self):
        i = (self.k & 1) if self.world > 1 else 0
        self.k += 1
        if self.world > 1:
            self._complete(i)
        return i

    def gather(self, i, src):
        if self.world > 1:
            self.pending[i] = dist.all_gather_into_tensor(self.gb[i], src.contiguous(), async_op=True)

    def finish(self):
        if self.world > 1:
            for i in range(2):
                self._complete(i)

    def last(self, i):
        """Gathered (world * rows, ...) tensor of slot i (device), and its host copy or None."""
        return self.gb[i], (self.hb[i] if self.hb is not None else None)
