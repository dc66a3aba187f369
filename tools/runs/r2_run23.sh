#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
MELSPEC_FORCE_GENERIC=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/run23_tests.txt
MELSPEC_FORCE_GENERIC=1 MELSPEC_GENERIC_PAIR=2 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $O/run23_tests.txt
timeout 600 python tools/bench_generic.py > $O/run23_rows.txt 2>&1
MELSPEC_GENERIC_PAIR=2 timeout 600 python tools/bench_generic.py >> $O/run23_rows.txt 2>&1
MELSPEC_GENERIC_PAIR=0 timeout 600 python tools/bench_generic.py >> $O/run23_rows.txt 2>&1
cat $O/run23_tests.txt $O/run23_rows.txt
