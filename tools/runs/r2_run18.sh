#!/bin/bash
# general plan, pair form: parity tests, then the general rows of bench_next_rows with the one-frame kernel (MELSPEC_GENERIC_PAIR=0),
# the run-time pair kernel (=2) and the compiled-in sizes (=1); then the plan-512 variants of r2_run17
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run18*.txt
timeout 600 python -m pytest tests/test_generic_plan.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 > $O/run18_tests.txt
MELSPEC_FORCE_GENERIC=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $O/run18_tests.txt
for m in 0 2 1; do
  MELSPEC_GENERIC_PAIR=$m timeout 600 python tools/bench_generic.py >> $O/run18_rows.txt 2>&1
done
MELSPEC_GENERIC_PAIR=2 MELSPEC_PAIR_MIN_WARPS=3 timeout 600 python tools/bench_generic.py >> $O/run18_rows.txt 2>&1
MELSPEC_GENERIC_PAIR=1 MELSPEC_PAIR_MIN_WARPS=3 timeout 600 python tools/bench_generic.py >> $O/run18_rows.txt 2>&1
cat $O/run18_tests.txt $O/run18_rows.txt
tools/r2_run17.sh
