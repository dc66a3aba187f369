#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:generic -s 2 -c 1 -f -o $O/gen1024 python tools/prof_generic.py 1024 256 128 > $O/gen1024.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:generic -s 2 -c 1 -f -o $O/gen480 python tools/prof_generic.py 480 160 80 > $O/gen480.log 2>&1
tail -3 $O/gen1024.log $O/gen480.log
