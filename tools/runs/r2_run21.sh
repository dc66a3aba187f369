#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests/test_generic_plan.py -m gpu -x -q 2>&1 | tail -5 > $O/run21_tests.txt
MELSPEC_FORCE_GENERIC=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $O/run21_tests.txt
MELSPEC_FORCE_GENERIC=1 MELSPEC_GENERIC_PAIR=2 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $O/run21_tests.txt
timeout 600 python tools/bench_generic.py > $O/run21_rows.txt 2>&1
MELSPEC_GENERIC_PAIR=2 timeout 600 python tools/bench_generic.py >> $O/run21_rows.txt 2>&1
cat $O/run21_tests.txt $O/run21_rows.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:generic -s 2 -c 1 -f -o $O/gen1024b python tools/prof_generic.py 1024 256 128 > $O/gen1024.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:generic -s 2 -c 1 -f -o $O/gen480b python tools/prof_generic.py 480 160 80 > $O/gen480.log 2>&1
