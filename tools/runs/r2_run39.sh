#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run39.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/run39_tests.txt
for cfg in "MELSPEC_TILE_ORDER=0" "MELSPEC_MM_SYNC=0" "MELSPEC_MM_SYNC=4" "MELSPEC_MM_SYNC=8" "MELSPEC_MM_SYNC=16" "X=1"; do
  echo -n "$cfg " >> $O/run39.txt
  env $cfg timeout 300 python tools/bench512.py >> $O/run39.txt 2>&1
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:melspec512 --csv --log-file $O/nemo.csv python tools/prof_nemo.py > $O/nemo.log 2>&1
python - >> $O/run39.txt <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2/nemo.csv')) if len(r)>10 and r[0].isdigit()]
by={}
for r in rows: by.setdefault((int(r[0]), r[4]), {})[r[12]]=r[14]
for k in sorted(by):
    v=by[k]; print(k[0], k[1][:44], ' '.join(f"{m.split('.')[0][-20:]}={v[m]}" for m in sorted(v)))
P
cat $O/run39_tests.txt $O/run39.txt
