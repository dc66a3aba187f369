"""Multi-GPU plumbing for the batch path: clips are independent, so a batch shards contiguously over ranks with no
data-path collective.  torch.distributed is used only for the barrier, the max-over-ranks timing and the
optional gather of the output shards (NCCL all-gather over NVLink / NVSwitch; not part of the hot path)."""
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


def gather_output(local, n_clips: int, dist):
    """Optional: every rank receives the whole batch's features.  `local` is this rank's output shard
    [hi - lo, ...] for `shard_range(n_clips, rank, world)`; returns [n_clips, ...] in clip order on every rank.

    One `all_gather_into_tensor` (NCCL on GPUs: ring / NVLS over NVSwitch).  When the batch divides evenly the gathered
    buffer *is* the result (no copy); otherwise shards are padded to the largest shard (sizes differ by at most one clip)
    and stitched.  The mel tensor of BASELINE configs[3] is 0.98 GB per GPU, so the gather (>= 7.6 ms at 900 GB/s) costs
    ~17x the kernel that produced it: it is an opt-in convenience, reported separately by `bench.py --gather`."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_range(n_clips, rank, world)
    if tuple(local.shape[:1]) != (hi - lo,):
        raise ValueError(f"rank {rank}: shard has {local.shape[0]} clips, expected {hi - lo}")
    if world == 1:
        return local
    per = -(-n_clips // world)
    rest = tuple(local.shape[1:])
    send = local.contiguous()
    if hi - lo < per:   # pad the short shards by one clip
        send = torch.zeros((per,) + rest, dtype=local.dtype, device=local.device)
        send[: hi - lo] = local
    buf = torch.empty((world * per,) + rest, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, send)
    if n_clips % world == 0:
        return buf
    parts = []
    for r in range(world):
        a, b = shard_range(n_clips, r, world)
        parts.append(buf[r * per: r * per + (b - a)])
    return torch.cat(parts, dim=0)
