#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/gputests4.txt
python tools/dbg_i16.py > gpurun_out/r2/dbg_i16.txt 2>&1
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1"
P="import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d['ms_per_step'],5), 'ms', round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], 'e2e', round(d['e2e']['ms_per_step'],2), d['e2e']['matches_device_path'])"
for i in 1 2; do
MELSPEC_B200_LIB=$PWD/mel-spec_b200/lib/libmelspec_r1.so $B --workload cfg3 2>/dev/null | python -c "$P" r1_cfg3
MELSPEC_CMN_FUSED=1 $B --workload cfg3 2>/dev/null | python -c "$P" new_cfg3_mode1
MELSPEC_CMN_FUSED=2 $B --workload cfg3 2>/dev/null | python -c "$P" new_cfg3_mode2
MELSPEC_CMN_FUSED=0 $B --workload cfg3 2>/dev/null | python -c "$P" new_cfg3_mode0
$B --workload cfg2 --no-extra 2>/dev/null | python -c "$P" new_cfg2
done > gpurun_out/r2/ab4.txt 2>&1
python tools/bench_next_rows.py > gpurun_out/r2/next_rows4.jsonl 2> gpurun_out/r2/next4.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -k regex:melspec -c 20 --csv --log-file gpurun_out/r2/launches_cfg3.csv python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2/l3.log 2>&1
tail -4 gpurun_out/r2/gputests4.txt; cat gpurun_out/r2/dbg_i16.txt | tail -5; cat gpurun_out/r2/ab4.txt; tail -3 gpurun_out/r2/launches_cfg3.csv | cut -c1-300
