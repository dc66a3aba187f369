#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:melspec400 --csv --log-file $O/layouts.csv python tools/prof_layouts.py > $O/layouts.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2/layouts.csv')) if len(r)>10 and r[0].isdigit()]
by={}
for r in rows: by.setdefault((int(r[0]), r[4]), {})[r[12]]=r[14]
for k in sorted(by):
    v=by[k]; print(k[0], k[1][:40], ' '.join(f"{m.split('.')[0][-28:]}={v[m]}" for m in sorted(v)))
P
