#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/r2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:melspec400 -c 1 -f -o gpurun_out/r2/full400f python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2/f400f.log 2>&1
tail -2 gpurun_out/r2/f400f.log
