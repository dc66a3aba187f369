#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests/test_vad.py tests/test_formats.py -m gpu -x -q 2>&1 | tail -5 > $O/run40.txt
timeout 300 python tools/bench_next_rows.py 2>/dev/null | grep "f-4\|f-3" | cut -c1-150 >> $O/run40.txt
cat $O/run40.txt
