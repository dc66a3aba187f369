#!/bin/bash
# A/B of two builds of the library on the same box: tools/ab_bench.sh <libA.so> <libB.so> [workload] [reps]
# Alternates the two builds `reps` times (device-resident bench only) and prints ms_per_step of each run.
A=$1; B=$2; W=${3:-cfg2}; R=${4:-3}
cd ${GRAFT_REPO_ROOT:-.}
for i in $(seq 1 $R); do
  for L in $A $B; do
    MELSPEC_B200_LIB=$PWD/$L python bench.py --workload $W --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-extra 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$L', '$W', round(d['ms_per_step'],5), 'ms', round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
  done
done
