#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run47.txt
for L in build/lib_g0.so mel-spec_b200/lib/libmelspec_b200.so build/lib_g2.so; do
  echo "== $L" >> $O/run47.txt
  MELSPEC_B200_LIB=$PWD/$L timeout 300 python tools/dbg_mm2.py 128 1 600 202 2>&1 | tail -1 | cut -c1-200 >> $O/run47.txt
  MELSPEC_B200_LIB=$PWD/$L timeout 300 python tools/bench512.py >> $O/run47.txt 2>&1
done
for L in build/lib_g0.so mel-spec_b200/lib/libmelspec_b200.so build/lib_g2.so; do
  echo "== $L" >> $O/run47.txt
  MELSPEC_B200_LIB=$PWD/$L timeout 300 python tools/bench512.py >> $O/run47.txt 2>&1
done
cat $O/run47.txt
