#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests/test_generic_plan.py -m gpu -x -q 2>&1 | tail -5 > $O/run24_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" >> $O/run24_tests.txt 2>&1
timeout 600 python tools/bench_generic.py > $O/run24_rows.txt 2>&1
MELSPEC_PAIR_INPLACE=0 timeout 600 python tools/bench_generic.py >> $O/run24_rows.txt 2>&1
cat $O/run24_tests.txt $O/run24_rows.txt
