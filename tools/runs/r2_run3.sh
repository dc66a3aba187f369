#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/gputests3.txt
python bench.py --no-extra --no-cpu-baseline > gpurun_out/r2/bench_cfg2_i16.json 2> gpurun_out/r2/bench3.err
ncu --set full --clock-control none --import-source on -k regex:melspec400 -c 1 -f -o gpurun_out/r2/full400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --e2e-steps 1 > gpurun_out/r2/f400.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:melspec512 -c 1 -f -o gpurun_out/r2/full512k python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2/f512.log 2>&1
tail -5 gpurun_out/r2/gputests3.txt; cut -c1-300 gpurun_out/r2/bench_cfg2_i16.json; tail -3 gpurun_out/r2/bench3.err
