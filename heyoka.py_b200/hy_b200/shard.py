"""Trajectory-range sharding helpers for the multi-GPU ensemble path
(shards only, no data-path collective)."""


def shard_bounds(total, rank, world):
    """Contiguous trajectory range [lo, hi) of `rank` out of `world` shards;
    sizes differ by at most one and cover [0, total) exactly."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("invalid rank/world")
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def device_for_iteration(i, n_devices):
    """ensemble_propagate_*: iteration i runs on device i mod n_devices."""
    return i % max(1, int(n_devices))


def reduce_throughput(local_seconds, local_units, dist=None, device=None):
    """Whole-job throughput = units of all ranks / max-over-ranks time.
    `dist` is torch.distributed (initialised) or None for a single process."""
    import torch

    t = torch.tensor([float(local_seconds)], dtype=torch.float64, device=device)
    u = torch.tensor([float(local_units)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item()) / float(t.item()), float(t.item()), float(u.item())
