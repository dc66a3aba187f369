#!/usr/bin/env python
"""Optional output gather through the C ABI (melspec_nccl_unique_id / _init / melspec_gather_nccl) on N GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/gather_nccl_check.py
Every rank computes its own shard of clips (no collective on the hot path), the shards are gathered with one ncclAllGather
issued by the library, and the result is compared with torch.distributed's all_gather of the same shards.  The unique id
travels over torch.distributed (gloo) here; a Rust host would use any channel it has.  Prints one JSON line on rank 0."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import mel_spec_b200 as ms
from bench import synth_batch_torch


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    L = ms.lib()
    h = ms.CudaMelSpectrogram(400, 160, 16000.0, 80, device=local)
    ident = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_uint8 * 128)()
        assert L.melspec_nccl_unique_id(buf) == 0, ms.last_error()
        ident = torch.tensor(list(buf), dtype=torch.uint8)
    ident = ident.to(dev)
    dist.broadcast(ident, 0)
    idb = (C.c_uint8 * 128)(*ident.cpu().tolist())
    assert L.melspec_nccl_init(h._h, idb, rank, world) == 0, ms.last_error()
    clips, n = 256, 160000
    x = synth_batch_torch(torch, clips, n, dev, rank)
    F = h.num_frames(n)
    shard = torch.empty((clips, F, 80), dtype=torch.float32, device=dev)
    full = torch.empty((world * clips, F, 80), dtype=torch.float32, device=dev)
    st = torch.cuda.Stream(device=dev)
    h.compute_device(x, clips, n, n, shard, stream=st)
    assert L.melspec_gather_nccl(h._h, shard.data_ptr(), shard.numel(), full.data_ptr(), st.cuda_stream) == 0, ms.last_error()
    st.synchronize()
    ref = torch.empty_like(full)
    dist.all_gather_into_tensor(ref, shard)
    torch.cuda.synchronize()
    same = bool(torch.equal(full, ref))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(5):
            L.melspec_gather_nccl(h._h, shard.data_ptr(), shard.numel(), full.data_ptr(), st.cuda_stream)
        e1.record(st)
    st.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 5], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        recv = (world - 1) * shard.numel() * 4
        print(json.dumps({"check": "melspec_gather_nccl == torch all_gather_into_tensor on every rank", "ok": bool(ok.item()), "world": world,
                          "shard_floats": shard.numel(), "ms": float(t.item()), "GB/s_received_per_gpu": recv / (float(t.item()) * 1e-3) / 1e9}))
    assert L.melspec_nccl_destroy(h._h) == 0
    h.close()
    dist.destroy_process_group()
    if not same:
        sys.exit(1)


if __name__ == "__main__":
    main()
