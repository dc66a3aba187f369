#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/run27_tests.txt
timeout 600 python bench.py --no-cpu-baseline > $O/run27_bench.json 2> $O/run27_bench.err
for i in 1 2; do
MELSPEC_KSPEC4=0 timeout 300 python tools/bench_next_rows.py 2>/dev/null | grep "large-v3" | cut -c1-140 >> $O/run27_k4.txt
timeout 300 python tools/bench_next_rows.py 2>/dev/null | grep "large-v3" | cut -c1-140 >> $O/run27_k4.txt
done
cat $O/run27_tests.txt $O/run27_k4.txt; tail -3 $O/run27_bench.err; cut -c1-200 $O/run27_bench.json
