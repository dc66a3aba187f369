#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
MELSPEC_B200_LIB=$PWD/${1:-build/lib_et1.so} timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_onset_parity.py tests/test_generic_plan.py -m gpu -x -q 2>&1 | tail -60 | cut -c1-220 > $O/run54.txt
cat $O/run54.txt
