#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
for T in racecheck memcheck initcheck synccheck; do
  timeout 600 compute-sanitizer --tool $T python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mel_major_large" > $O/san44_$T.log 2>&1
  echo "== $T"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Uninitialized|Invalid|passed|failed" $O/san44_$T.log | head -12
  grep -B2 -A12 "Race reported\|Uninitialized __shared__\|Invalid __" $O/san44_$T.log | head -60
done
