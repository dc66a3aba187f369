#!/bin/bash
# plan-512 compile-time projection schedule: parity tests, then A/B against the table-driven loop (MELSPEC_KSCHED=0)
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/run16_tests.txt
for i in 1 2; do
  MELSPEC_KSCHED=0 timeout 300 python tools/bench512.py >> $O/run16_bench512.txt 2>&1
  timeout 300 python tools/bench512.py >> $O/run16_bench512.txt 2>&1
done
cat $O/run16_tests.txt $O/run16_bench512.txt
