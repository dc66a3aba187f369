"""Multi-GPU plumbing for the batch path: clips are independent, so a batch shards contiguously over ranks with no
data-path collective.  torch.distributed is used only for the barrier and the max-over-ranks timing."""
from __future__ import annotations


def shard_range(n_clips: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of rank `rank`: sizes differ by at most one clip, union is exactly range(n_clips)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_clips, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def whole_job_rate(units_this_rank: int, seconds_this_rank: float, dist=None, device=None) -> float:
    """Whole-job throughput: sum of units over ranks / max of the per-rank times (device times, never wall clock)."""
    import torch
    u = torch.tensor([float(units_this_rank)], dtype=torch.float64, device=device)
    t = torch.tensor([float(seconds_this_rank)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(u.item() / t.item())
