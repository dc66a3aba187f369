#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/r2
for i in 1 2; do
for L in libmelspec_r1.so libmelspec_twsmem.so libmelspec_tw512smem.so libmelspec_nohint.so libmelspec_b200.so; do
  MELSPEC_B200_LIB=$PWD/mel-spec_b200/lib/$L python tools/bench512.py 2>&1 | tail -1
done
done > gpurun_out/r2/ab512.txt
python bench.py --no-extra --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['e2e_int16_pcm'])" > gpurun_out/r2/i16flag.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:melspec512 -c 1 -f -o gpurun_out/r2/full512w python tools/bench512.py > gpurun_out/r2/f512w.log 2>&1
cat gpurun_out/r2/ab512.txt; cat gpurun_out/r2/i16flag.txt
