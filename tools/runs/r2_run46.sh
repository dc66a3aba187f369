#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run46.txt
for cfg in "X=1" "MELSPEC_TILE_ORDER=0" "MELSPEC_KSPEC5=0" "MELSPEC_KSPEC5=0 MELSPEC_TILE_ORDER=0"; do
  env $cfg timeout 300 python tools/dbg_mm2.py 128 1 300 202 >> $O/run46.txt 2>&1
done
timeout 300 python tools/dbg_mm2.py 128 0 300 202 >> $O/run46.txt 2>&1
timeout 300 python tools/dbg_mm2.py 80 1 300 202 >> $O/run46.txt 2>&1
timeout 300 python tools/dbg_mm2.py 80 0 300 202 >> $O/run46.txt 2>&1
timeout 300 python tools/dbg_mm2.py 128 1 300 204 >> $O/run46.txt 2>&1
timeout 300 python tools/dbg_mm2.py 80 0 300 998 >> $O/run46.txt 2>&1
cat $O/run46.txt
