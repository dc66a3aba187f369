#!/bin/bash
# plan-512 variants: window starts / mel indices in registers (v10), prescale table only in scaled passes (v01), both (v11)
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run17.txt
for i in 1 2; do
  for v in 00 10 01 11; do
    echo -n "v$v " >> $O/run17.txt
    MELSPEC_B200_LIB=$PWD/build/lib_v$v.so timeout 300 python tools/bench512.py >> $O/run17.txt 2>&1
  done
done
cat $O/run17.txt
