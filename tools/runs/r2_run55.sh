#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run55.txt
for i in 1 2 3; do for L in "$@"; do
  MELSPEC_B200_LIB=$PWD/$L timeout 300 python tools/bench512.py >> $O/run55.txt 2>&1
done; done
cat $O/run55.txt
