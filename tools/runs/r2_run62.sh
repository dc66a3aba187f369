#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run62.txt
cp build/lib_rf1.so mel-spec_b200/lib/libmelspec_b200.so; touch mel-spec_b200/lib/libmelspec_b200.so
timeout 60 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | cut -c1-200 >> $O/run62.txt
for i in 1 2; do for L in build/lib_rf0.so build/lib_rf1.so; do
  MELSPEC_B200_LIB=$PWD/$L timeout 20 python tools/bench512.py 2>&1 | tail -1 >> $O/run62.txt
done; done
cat $O/run62.txt
