#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run57.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | cut -c1-200 >> $O/run57.txt
timeout 300 python tools/dbg_mm2.py 128 1 300 202 2>&1 | tail -1 | cut -c1-160 >> $O/run57.txt
for i in 1 2; do timeout 300 python tools/bench512.py >> $O/run57.txt 2>&1; done
timeout 600 python bench.py > $O/run57_bench.json 2>$O/run57_bench.err
python - >> $O/run57.txt <<'PY'
import json
d=json.loads(open('gpurun_out/r2/run57_bench.json').read().strip().splitlines()[-1])
print('bench', d['ms_per_step'], d['value'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'], 'e2e', d['e2e']['value'])
for k in ('cfg3','cfg4shard','cfg5_stream'):
    v=d['extra'][k]; print(k, v['ms_per_step'], v['roofline']['frac'])
PY
cat $O/run57.txt
