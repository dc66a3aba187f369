#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/run26_tests.txt
timeout 600 python tools/bench_generic.py > $O/run26_rows.txt 2>&1
MELSPEC_GENERIC_PAIR=0 timeout 600 python tools/bench_generic.py >> $O/run26_rows.txt 2>&1
cat $O/run26_tests.txt $O/run26_rows.txt
