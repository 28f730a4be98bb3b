"""Host-side multi-GPU plumbing: contiguous index ranges and the over-ranks reduction of a
benchmark report. No collective touches the step path (trajectories are independent)."""


def shard_range(n, world, rank):
    """Contiguous [lo, hi) of rank `rank` when [0, n) is split over `world` GPUs
    (the same ceil-division rule as ode_b200_integrate_batch_multi)."""
    per = (n + world - 1) // world
    return min(n, per * rank), min(n, per * (rank + 1))


def reduce_report(dist, device, my_ms, my_work):
    """(max over ranks of the device time, sum over ranks of each work counter)."""
    import torch
    if dist is None:
        return float(my_ms), [int(w) for w in my_work]
    t = torch.tensor([float(my_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    w = torch.tensor([int(x) for x in my_work], dtype=torch.int64, device=device)
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return float(t[0]), [int(x) for x in w]
