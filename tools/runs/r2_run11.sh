#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/r2
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2/gputests11.txt
for i in 1 2; do
timeout 200 python tools/bench512.py 2>&1 | tail -1
done > gpurun_out/r2/ab11.txt
timeout 200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kaldi_fused or kaldi_batch" 2>&1 | tail -6 > gpurun_out/r2/race11.txt
tail -6 gpurun_out/r2/gputests11.txt; cat gpurun_out/r2/ab11.txt; cat gpurun_out/r2/race11.txt
