#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run50.txt
MELSPEC_B200_LIB=$PWD/build/lib_au1.so timeout 300 python tools/dbg_mm2.py 128 1 600 202 2>&1 | tail -1 | cut -c1-160 >> $O/run50.txt
MELSPEC_B200_LIB=$PWD/build/lib_au1.so timeout 300 python tools/dbg_mm2.py 80 1 600 202 2>&1 | tail -1 | cut -c1-160 >> $O/run50.txt
for i in 1 2 3; do for L in build/lib_au0.so build/lib_au1.so; do
  MELSPEC_B200_LIB=$PWD/$L timeout 300 python tools/bench512.py >> $O/run50.txt 2>&1
done; done
MELSPEC_B200_LIB=$PWD/build/lib_au1.so timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> $O/run50.txt
cat $O/run50.txt
