#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run49.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $O/run49.txt
timeout 300 python tools/dbg_mm2.py 128 1 600 202 2>&1 | tail -1 | cut -c1-160 >> $O/run49.txt
timeout 300 python tools/dbg_mm2.py 80 1 600 202 2>&1 | tail -1 | cut -c1-160 >> $O/run49.txt
for i in 1 2; do timeout 300 python tools/bench512.py >> $O/run49.txt 2>&1; done
timeout 600 python bench.py > $O/run49_bench.json 2>$O/run49_bench.err
tail -c 3000 $O/run49_bench.json >> $O/run49.txt
cat $O/run49.txt
