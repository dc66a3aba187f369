#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/run52_tests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" >> $O/run52_tests.txt 2>&1
bash tools/measure_round.sh full > $O/run52_measure.log 2>&1
cat $O/run52_tests.txt
for T in memcheck racecheck synccheck; do grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/m/san_$T.log | tail -2; done
