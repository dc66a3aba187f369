#!/bin/bash
# (NOT MEASURED: the first version of this script read the wrong stdout line and the round's GPU budget ended with it.)
# e2e at N = 4 against the pipeline chunk size of melspec_compute_host (MELSPEC_HOST_CHUNK_MB), same box, alternating
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/r2; mkdir -p $O; rm -f $O/run61.txt
port=29520
for MB in 32 128 32 128; do
  port=$((port+1))
  MELSPEC_HOST_CHUNK_MB=$MB timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 2>/dev/null \
   | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); e=d['e2e']; print('chunk_mb $MB', 'e2e_ms', round(e['ms_per_step'],2), 'ceiling_ms', round(e['copy_ceiling_ms'],2), 'frac', round(e['frac_of_copy_ceiling'],3), 'i16_ms', round(d['e2e_int16_pcm']['ms_per_step'],2))" >> $O/run61.txt
done
cat $O/run61.txt
