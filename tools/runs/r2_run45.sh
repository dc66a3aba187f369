#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run45.txt
for cfg in "X=1" "MELSPEC_NO_PRESCALE=1" "MELSPEC_TILE_ORDER=0" "MELSPEC_KSPEC5=0" "MELSPEC_KSPEC5=0 MELSPEC_TILE_ORDER=0"; do
  echo "== $cfg" >> $O/run45.txt
  env $cfg timeout 300 python tools/dbg_mm.py 2>&1 | tail -8 | cut -c1-200 >> $O/run45.txt
done
cat $O/run45.txt
