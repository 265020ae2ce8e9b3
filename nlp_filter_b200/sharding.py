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
