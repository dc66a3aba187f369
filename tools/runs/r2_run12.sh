#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out/r2
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2/gputests12.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke12.txt 2>&1
timeout 400 python tools/bench_next_rows.py 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    r = json.loads(ln)
    if 'general' in r['row']: print(round(r['ms_per_step'], 4), round(r['value'] / 1e6, 1), r['row'][:70])
" > gpurun_out/r2/general12.txt
tail -4 gpurun_out/r2/gputests12.txt; cat gpurun_out/r2/smoke12.txt; cat gpurun_out/r2/general12.txt
