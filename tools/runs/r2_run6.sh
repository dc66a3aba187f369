#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/gputests6.txt
for i in 1 2; do
for L in libmelspec_r1.so libmelspec_b200.so; do
  MELSPEC_B200_LIB=$PWD/mel-spec_b200/lib/$L python tools/bench512.py 2>&1 | tail -1
done
done > gpurun_out/r2/ab512b.txt
tail -5 gpurun_out/r2/gputests6.txt; cat gpurun_out/r2/ab512b.txt
