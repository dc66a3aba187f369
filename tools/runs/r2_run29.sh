#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
MELSPEC_GENERIC_PAIR=0 timeout 900 python -m pytest tests/test_generic_plan.py -m gpu -x -q 2>&1 | tail -4 > $O/run29_tests.txt
MELSPEC_GENERIC_PAIR=0 MELSPEC_FORCE_GENERIC=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $O/run29_tests.txt
timeout 900 python -m pytest tests/test_generic_plan.py -m gpu -x -q 2>&1 | tail -4 >> $O/run29_tests.txt
MELSPEC_GENERIC_PAIR=0 timeout 600 python tools/bench_generic.py > $O/run29_rows.txt 2>&1
timeout 600 python tools/bench_generic.py >> $O/run29_rows.txt 2>&1
cat $O/run29_tests.txt $O/run29_rows.txt
