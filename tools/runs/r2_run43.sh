#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run43.txt
for i in 1 2 3 4; do timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mel_major_large" 2>&1 | grep -E "^E  |passed|failed|assert" | cut -c1-220 | head -12 >> $O/run43.txt; done
cat $O/run43.txt
