#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run58.txt
A=$1; B=$2
cp $B mel-spec_b200/lib/libmelspec_b200.so; touch mel-spec_b200/lib/libmelspec_b200.so
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | cut -c1-200 >> $O/run58.txt
timeout 300 python tools/dbg_mm2.py 80 0 300 202 2>&1 | tail -1 | cut -c1-160 >> $O/run58.txt
for i in 1 2 3; do for L in $A $B; do
  MELSPEC_B200_LIB=$PWD/$L timeout 300 python tools/bench512.py >> $O/run58.txt 2>&1
done; done
cat $O/run58.txt
