#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/r2
for i in 1 2; do
timeout 200 python tools/bench512.py 2>&1 | tail -1
MELSPEC_B200_LIB=$PWD/mel-spec_b200/lib/libmelspec_fast.so timeout 200 python tools/bench512.py 2>&1 | tail -1
done > gpurun_out/r2/ab15.txt
MELSPEC_B200_LIB=$PWD/mel-spec_b200/lib/libmelspec_fast.so timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> gpurun_out/r2/ab15.txt
cat gpurun_out/r2/ab15.txt
