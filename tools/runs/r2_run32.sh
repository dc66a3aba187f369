#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run32_ab.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/run32_tests.txt
for i in 1 2; do
for t in 0 1; do for k in 0 1; do
echo -n "TILE_ORDER=$t KSPEC5=$k " >> $O/run32_ab.txt
MELSPEC_TILE_ORDER=$t MELSPEC_KSPEC5=$k timeout 300 python tools/bench_next_rows.py 2>/dev/null | grep "f-3a" | cut -c1-130 >> $O/run32_ab.txt
done; done; done
timeout 300 python tools/bench512.py >> $O/run32_ab.txt 2>&1
cat $O/run32_tests.txt $O/run32_ab.txt
