#!/bin/bash
# Round measurement pass (run on the GPU box through gpurun): bench lines, launch lists and ncu captures into gpurun_out/m/.
#   tools/measure_round.sh [full]      (full: also the `ncu --set full` captures and the sanitizer runs)
# Every command runs under its own `timeout`.
set -x
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out/m
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv > $O/gpu.txt
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench.err                                   # cfg2 + extra (cfg3, cfg4shard, cfg5 stream) + cpu_baseline
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2>> $O/bench.err
timeout 600 python tools/bench_next_rows.py > $O/next_rows.jsonl 2>> $O/bench.err
timeout 300 python tools/bench512.py > $O/bench512.txt 2>> $O/bench.err
# general plan: pair form (default) against the one-frame kernel
timeout 300 python tools/bench_generic.py > $O/bench_generic.txt 2>> $O/bench.err
MELSPEC_GENERIC_PAIR=0 timeout 300 python tools/bench_generic.py >> $O/bench_generic.txt 2>> $O/bench.err
# worst-case parity of the two-frames-per-transform packing: the new tests against this build and against the round-1 build
timeout 300 python -m pytest tests/test_onset_parity.py -m gpu -q -s 2>&1 | grep -E "worst|passed|failed" > $O/onset_parity.txt
# (the round-1 library lacks the ABI-2 entry points the Python loader now binds, so this block only runs against an old checkout's
# build placed there by hand; profiles/r2_ab_r1_vs_r2_cfg2.txt and the lower half of r2_onset_parity.txt are from when it still loaded)
if [ -f mel-spec_b200/lib/libmelspec_r1.so ]; then
  echo "--- the same tests against the round-1 library (before the pair prescale)" >> $O/onset_parity.txt
  MELSPEC_B200_LIB=$PWD/mel-spec_b200/lib/libmelspec_r1.so timeout 300 python -m pytest tests/test_onset_parity.py -m gpu -q 2>&1 | grep -E "^E   +Assertion|^FAILED|passed|failed" >> $O/onset_parity.txt
  # A/B of the headline kernel: round-1 build vs this build, alternating
  timeout 600 tools/ab_bench.sh mel-spec_b200/lib/libmelspec_r1.so mel-spec_b200/lib/libmelspec_b200.so cfg2 3 > $O/ab_r1_vs_r2.txt 2>&1
fi
# launch lists: only kernels of the library (the torch kernels of the same command generate the synthetic PCM before the
# timed region and would exhaust any launch-count limit)
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for W in cfg2 cfg3 cfg4shard; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:melspec -c 30 --csv --log-file $O/launches_$W.csv python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/l_$W.log 2>&1
done
if [ "$1" = "full" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:melspec400 -c 1 -f -o $O/full400 python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/f400.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:melspec512 -c 1 -f -o $O/full512 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/f512.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:generic_pair -s 2 -c 1 -f -o $O/fullgen1024 python tools/prof_generic.py 1024 256 128 > $O/fgen1024.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:generic_pair -s 2 -c 1 -f -o $O/fullgen480 python tools/prof_generic.py 480 160 80 > $O/fgen480.log 2>&1
# digests of the general-plan captures are made here; their .ncu-rep files stay on the box (gpurun_out/ is capped at 64 MiB)
for G in 1024 480; do
  P=$(( G == 1024 ? 318464 : 510976 ))   # frame pairs per launch
  { python tools/ncu_summary.py $O/fullgen$G.ncu-rep; python tools/ncu_smem.py $O/fullgen$G.ncu-rep $P | tail -12; python tools/ncu_ophist.py $O/fullgen$G.ncu-rep $P; } > $O/ncu_full_generic_pair_fft$G.txt 2>&1
  rm -f $O/fullgen$G.ncu-rep
done
for T in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $T python -m pytest tests/test_gpu_parity.py tests/test_onset_parity.py tests/test_boundary_r2.py -m gpu -x -q tests/test_generic_plan.py -k "synthetic_batch or ragged_lengths or kaldi_fused or kaldi_batch or nemo_features or click_and_silence or int16_host or spectrogram_add_reference or pair_form_unaligned or generic_batch_layouts or nemo_ragged" > $O/san_$T.log 2>&1
done
fi
cut -c1-300 $O/bench_n1.json
