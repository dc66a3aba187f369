#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
MELSPEC_KSPEC5=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:melspec400 -s 2 -c 1 -f -o $O/mm python tools/prof_mm.py > $O/mm.log 2>&1
python tools/ncu_summary.py $O/mm.ncu-rep > $O/mm_digest.txt 2>&1
python tools/ncu_ophist.py $O/mm.ncu-rep 171008 >> $O/mm_digest.txt 2>&1
rm -f $O/mm.ncu-rep
head -60 $O/mm_digest.txt
