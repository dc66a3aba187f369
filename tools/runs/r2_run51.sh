#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run51.txt
for L in build/lib_e2.so build/lib_late.so build/lib_e2late.so; do
MELSPEC_B200_LIB=$PWD/$L timeout 300 python tools/dbg_mm2.py 128 1 600 202 2>&1 | tail -1 | cut -c1-160 >> $O/run51.txt
done
for i in 1 2 3; do for L in build/lib_au0.so build/lib_e2.so build/lib_late.so build/lib_e2late.so; do
  MELSPEC_B200_LIB=$PWD/$L timeout 300 python tools/bench512.py >> $O/run51.txt 2>&1
done; done
cat $O/run51.txt
