#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests/test_generic_plan.py -m gpu -x -q 2>&1 | tail -5 > $O/run20_tests.txt
MELSPEC_FORCE_GENERIC=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $O/run20_tests.txt
timeout 600 python tools/bench_generic.py > $O/run20_rows.txt 2>&1
MELSPEC_GENERIC_PAIR=2 timeout 600 python tools/bench_generic.py >> $O/run20_rows.txt 2>&1
cat $O/run20_tests.txt $O/run20_rows.txt
