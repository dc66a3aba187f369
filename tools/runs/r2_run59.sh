#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_final.json 2> $O/bench_n2_final.err
tail -c 600 $O/bench_n2_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_n2_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling','gpu_launches')}, d['roofline']['frac'], d['config']['workload'], d['e2e']['value'], d['e2e'].get('frac_of_copy_ceiling'), d['clocks'])
PY
