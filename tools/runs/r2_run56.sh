#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run56.txt
cp build/lib_v3.so mel-spec_b200/lib/libmelspec_b200.so; touch mel-spec_b200/lib/libmelspec_b200.so
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | cut -c1-200 >> $O/run56.txt
for i in 1 2 3; do for L in build/lib_et0.so build/lib_v3nolead.so build/lib_v3.so; do
  MELSPEC_B200_LIB=$PWD/$L timeout 300 python tools/bench512.py >> $O/run56.txt 2>&1
done; done
cat $O/run56.txt
