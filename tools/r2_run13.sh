#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/r2
timeout 900 tools/ab_bench.sh mel-spec_b200/lib/libmelspec_b200.so mel-spec_b200/lib/libmelspec_meta.so cfg2 3 > gpurun_out/r2/ab13.txt 2>&1
cat gpurun_out/r2/ab13.txt
