#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/r2
timeout 900 tools/ab_bench.sh mel-spec_b200/lib/libmelspec_b200.so mel-spec_b200/lib/libmelspec_fast.so cfg2 3 > gpurun_out/r2/ab13.txt 2>&1
cat gpurun_out/r2/ab13.txt
MELSPEC_B200_LIB=$PWD/mel-spec_b200/lib/libmelspec_fast.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_onset_parity.py -m gpu -x -q 2>&1 | tail -3
