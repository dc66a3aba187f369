#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for L in build/lib_mmh.so mel-spec_b200/lib/libmelspec_b200.so; do
MELSPEC_B200_LIB=$PWD/$L timeout 600 ncu --metrics $M --clock-control none -k regex:melspec400 --csv --log-file $O/layouts2.csv python tools/prof_layouts.py > $O/layouts2.log 2>&1
echo $L
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2/layouts2.csv')) if len(r)>10 and r[0].isdigit()]
by={}
for r in rows: by.setdefault((int(r[0]), r[4]), {})[r[12]]=r[14]
for k in sorted(by):
    v=by[k]; print(k[0], k[1][:40], ' '.join(f"{m.split('.')[0][-20:]}={v[m]}" for m in sorted(v)))
P
done
