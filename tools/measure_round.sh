#!/bin/bash
# Round measurement pass (run on the GPU box through gpurun): bench lines, launch lists and ncu captures into gpurun_out/m/.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/m
python bench.py > gpurun_out/m/bench_cfg2.json 2> gpurun_out/m/bench.err
python bench.py --workload cfg3 > gpurun_out/m/bench_cfg3.json 2>> gpurun_out/m/bench.err
python bench.py --workload cfg4shard --no-cpu-baseline > gpurun_out/m/bench_cfg4shard.json 2>> gpurun_out/m/bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/m/bench_reference.json 2>> gpurun_out/m/bench.err
python tools/stream_bench.py > gpurun_out/m/stream_cfg5.json 2>> gpurun_out/m/bench.err
python tools/bench_next_rows.py > gpurun_out/m/next_rows.jsonl 2>> gpurun_out/m/bench.err
# launch lists: only kernels of the library (the torch kernels of the same command generate the synthetic PCM before the
# timed region and would exhaust any launch-count limit)
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -k regex:melspec -c 60 --csv --log-file gpurun_out/m/launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/m/l2.log 2>&1
ncu --metrics $M --clock-control none -k regex:melspec -c 60 --csv --log-file gpurun_out/m/launches_cfg3.csv python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/m/l3.log 2>&1
if [ "$1" = "full" ]; then
ncu --set full --clock-control none --import-source on -k regex:melspec400 -c 1 -f -o gpurun_out/m/full400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/m/f400.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:melspec512 -c 1 -f -o gpurun_out/m/full512 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/m/f512.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:melspec_cmn -c 1 -f -o gpurun_out/m/fullcmn python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/m/fcmn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:melspec_generic -c 1 -f -o gpurun_out/m/fullgeneric python tools/bench_next_rows.py > gpurun_out/m/fgen.log 2>&1
fi
cut -c1-300 gpurun_out/m/bench_cfg2.json
