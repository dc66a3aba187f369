"""World-size-2 gloo test (CPU) of the multi-GPU plumbing: contiguous clip shards, no data-path collective, the
whole-job rate is sum(units) / max(time)."""
import os
import socket

import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    from mel_spec_b200.shard import shard_range, whole_job_rate
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(1025, rank, world)
    dist.barrier()
    rate = whole_job_rate((hi - lo) * 998, 0.5 + rank, dist)      # rank 1 is the slow one: 1.5 s
    # optional output gather (north_star: "only an optional NCCL gather of the output mel tensor"): uneven and even batches
    import torch
    from mel_spec_b200.shard import gather_output
    ok = True
    for n in (1025, 1024, 3, 1):
        a, b = shard_range(n, rank, world)
        local = (torch.arange(a, b, dtype=torch.float32)[:, None, None] * 10 + torch.arange(6, dtype=torch.float32).reshape(1, 2, 3))
        full = gather_output(local, n, dist)
        want = torch.arange(n, dtype=torch.float32)[:, None, None] * 10 + torch.arange(6, dtype=torch.float32).reshape(1, 2, 3)
        ok = ok and full.shape == want.shape and bool(torch.equal(full, want))
    q.put((rank, lo, hi, rate, ok))
    dist.destroy_process_group()


def test_two_rank_sharding_and_rate():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[4] for r in res), "gather_output must return the whole batch in clip order on every rank"
    res = [r[:4] for r in res]
    (r0, lo0, hi0, rate0), (r1, lo1, hi1, rate1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 513, 513, 1025)           # contiguous, disjoint, complete
    assert rate0 == rate1 == pytest.approx(1025 * 998 / 1.5)     # sum of units / max of times


def test_shard_ranges_partition():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from mel_spec_b200.shard import shard_range
    for n in (0, 1, 7, 8, 1024, 8192):
        for w in (1, 2, 4, 8):
            cuts = [shard_range(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
