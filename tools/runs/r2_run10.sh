#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/r2
timeout 120 python -m pytest tests/test_boundary_r2.py -m gpu -x -q -k "nemo_fused" 2>&1 | tail -12 > gpurun_out/r2/gputests10a.txt
cat gpurun_out/r2/gputests10a.txt
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/gputests10.txt
for i in 1 2; do
timeout 200 python tools/bench512.py 2>&1 | tail -1
MELSPEC_NORM_FUSED=0 timeout 200 python tools/bench512.py 2>&1 | tail -1
done > gpurun_out/r2/ab10.txt
tail -6 gpurun_out/r2/gputests10.txt; cat gpurun_out/r2/ab10.txt
