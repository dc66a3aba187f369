#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2/gpu.txt
MELSPEC_B200_LIB=$PWD/mel-spec_b200/lib/libmelspec_r1.so timeout 600 python -m pytest tests/test_onset_parity.py -q 2>&1 | tail -40 > gpurun_out/r2/onset_r1lib.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2/gputests.txt
tools/ab_bench.sh mel-spec_b200/lib/libmelspec_r1.so mel-spec_b200/lib/libmelspec_b200.so cfg2 3 > gpurun_out/r2/ab_prescale.txt 2>&1
( time python bench.py ) > gpurun_out/r2/bench_full.json 2> gpurun_out/r2/bench_full.err
tail -5 gpurun_out/r2/onset_r1lib.txt; tail -5 gpurun_out/r2/gputests.txt; cat gpurun_out/r2/ab_prescale.txt; tail -4 gpurun_out/r2/bench_full.err; cut -c1-600 gpurun_out/r2/bench_full.json
