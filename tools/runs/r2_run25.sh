#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests/test_generic_plan.py -m gpu -x -q 2>&1 | tail -5 > $O/run25_tests.txt
MELSPEC_PAIR_LB=1 MELSPEC_FORCE_GENERIC=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $O/run25_tests.txt
timeout 600 python tools/bench_generic.py > $O/run25_rows.txt 2>&1
MELSPEC_PAIR_LB=1 timeout 600 python tools/bench_generic.py >> $O/run25_rows.txt 2>&1
MELSPEC_PAIR_INPLACE=0 timeout 600 python tools/bench_generic.py >> $O/run25_rows.txt 2>&1
cat $O/run25_tests.txt $O/run25_rows.txt
