#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/gputests2.txt
for i in 1 2; do
for L in mel-spec_b200/lib/libmelspec_r1.so mel-spec_b200/lib/libmelspec_twsmem.so mel-spec_b200/lib/libmelspec_b200.so; do
  for W in cfg2 cfg3; do
  MELSPEC_B200_LIB=$PWD/$L python bench.py --workload $W --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$L', '$W', round(d['ms_per_step'],5), 'ms', round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
  done
done
done > gpurun_out/r2/ab2.txt 2>&1
tail -5 gpurun_out/r2/gputests2.txt; cat gpurun_out/r2/ab2.txt
