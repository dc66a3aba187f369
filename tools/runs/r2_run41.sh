#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
MELSPEC_B200_LIB=$PWD/build/lib_z64.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_onset_parity.py -m gpu -x -q 2>&1 | tail -3 > $O/run41.txt
timeout 600 tools/ab_bench.sh mel-spec_b200/lib/libmelspec_b200.so build/lib_z64.so cfg2 4 >> $O/run41.txt 2>&1
timeout 600 tools/ab_bench.sh mel-spec_b200/lib/libmelspec_b200.so build/lib_z64.so cfg4shard 2 >> $O/run41.txt 2>&1
cat $O/run41.txt
