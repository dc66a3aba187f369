#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run53.txt
A=${1:-build/lib_et0.so}; B=${2:-build/lib_et1.so}
MELSPEC_B200_LIB=$PWD/$B timeout 300 python tools/dbg_mm2.py 128 1 300 202 2>&1 | tail -1 | cut -c1-160 >> $O/run53.txt
MELSPEC_B200_LIB=$PWD/$B timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_onset_parity.py tests/test_generic_plan.py -m gpu -x -q 2>&1 | tail -2 >> $O/run53.txt
for i in 1 2 3; do for L in $A $B; do
  MELSPEC_B200_LIB=$PWD/$L timeout 300 python tools/bench512.py >> $O/run53.txt 2>&1
done; done
cat $O/run53.txt
