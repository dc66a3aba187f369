#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/r2
timeout 120 python -m pytest tests/test_generic_plan.py -m gpu -x -q -k "nemo_ragged" 2>&1 | tail -5 > gpurun_out/r2/gputests9a.txt
if ! grep -q " passed" gpurun_out/r2/gputests9a.txt; then cat gpurun_out/r2/gputests9a.txt; echo RAGGED_NEMO_FAILED; exit 1; fi
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/gputests9.txt
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-extra"
P="import sys,json; d=json.loads(sys.stdin.readline()); print(sys.argv[1], round(d['ms_per_step'],5), 'ms', round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], 'e2e', round(d['e2e']['ms_per_step'],2), d['e2e']['matches_device_path'], d.get('e2e_int16_pcm',{}).get('matches_f32_path_bit_exact'))"
for i in 1 2; do
for W in 12 16; do
MELSPEC_WARPS=$W timeout 300 $B --workload cfg2 2>/dev/null | python -c "$P" cfg2_w$W
MELSPEC_WARPS=$W timeout 300 $B --workload cfg4shard 2>/dev/null | python -c "$P" cfg4shard_w$W
done
done > gpurun_out/r2/ab9.txt 2>&1
MELSPEC_B200_LIB=$PWD/mel-spec_b200/lib/libmelspec_r1.so timeout 300 python tools/bench512.py 2>&1 | tail -1 >> gpurun_out/r2/ab9.txt
timeout 300 python tools/bench512.py 2>&1 | tail -1 >> gpurun_out/r2/ab9.txt
tail -6 gpurun_out/r2/gputests9.txt; cat gpurun_out/r2/ab9.txt
