#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run48.txt
MELSPEC_B200_LIB=$PWD/build/lib_g3.so timeout 300 python tools/dbg_mm2.py 128 1 600 202 2>&1 | tail -1 | cut -c1-160 >> $O/run48.txt
MELSPEC_B200_LIB=$PWD/build/lib_g3.so timeout 300 python tools/dbg_mm2.py 80 1 600 202 2>&1 | tail -1 | cut -c1-160 >> $O/run48.txt
for i in 1 2; do for L in build/lib_g0.so build/lib_g2.so build/lib_g3.so; do
  MELSPEC_B200_LIB=$PWD/$L timeout 300 python tools/bench512.py >> $O/run48.txt 2>&1
done; done
cat $O/run48.txt
