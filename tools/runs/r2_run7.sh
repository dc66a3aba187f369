#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gather_nccl.py tests/test_shard_gloo.py -q 2>&1 | tail -5 > gpurun_out/r2/gather_test.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/gather_nccl_check.py > gpurun_out/r2/gather_n2.json 2> gpurun_out/r2/gather_n2.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/r2/bench_n2.json 2> gpurun_out/r2/bench_n2.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 ) > gpurun_out/r2/bench_ref_n2.json 2> gpurun_out/r2/bench_ref_n2.err
cat gpurun_out/r2/gather_test.txt; cat gpurun_out/r2/gather_n2.json; tail -3 gpurun_out/r2/gather_n2.err; cut -c1-1500 gpurun_out/r2/bench_n2.json; tail -4 gpurun_out/r2/bench_n2.err; cut -c1-400 gpurun_out/r2/bench_ref_n2.json; tail -4 gpurun_out/r2/bench_ref_n2.err
