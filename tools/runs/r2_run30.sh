#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run30_k5.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/run30_tests.txt
for i in 1 2; do
MELSPEC_KSPEC5=0 timeout 300 python tools/bench_next_rows.py 2>/dev/null | grep "f-3a" | cut -c1-160 >> $O/run30_k5.txt
timeout 300 python tools/bench_next_rows.py 2>/dev/null | grep "f-3a" | cut -c1-160 >> $O/run30_k5.txt
done
cat $O/run30_tests.txt $O/run30_k5.txt
