#!/bin/bash
# Copy the results of tools/measure_round.sh full (gpurun_out/m) into profiles/ under the round's tag and rebuild the ncu digests.
#   tools/refresh_profiles.sh r2
T=${1:-r2}; M=gpurun_out/m; cd "$(dirname "$0")/.."
python tools/update_traffic.py $M $T
cp $M/bench_n1.json profiles/${T}_bench_n1.json
cp $M/bench_reference.json profiles/${T}_bench_reference.json
cp $M/next_rows.jsonl profiles/${T}_next_rows.jsonl
cp $M/bench512.txt profiles/${T}_bench512.txt
cp $M/bench_generic.txt profiles/${T}_bench_generic.txt
cp $M/ncu_full_generic_pair_fft1024.txt profiles/${T}_ncu_full_generic_pair_fft1024.txt
cp $M/ncu_full_generic_pair_fft480.txt profiles/${T}_ncu_full_generic_pair_fft480.txt
head -4 $M/onset_parity.txt > profiles/${T}_onset_parity.txt.new && tail -n +5 profiles/${T}_onset_parity.txt >> profiles/${T}_onset_parity.txt.new && mv profiles/${T}_onset_parity.txt.new profiles/${T}_onset_parity.txt
{ python tools/ncu_summary.py $M/full400.ncu-rep; python tools/ncu_smem.py $M/full400.ncu-rep 171008 | tail -12; python tools/ncu_ophist.py $M/full400.ncu-rep 171008; } > profiles/${T}_ncu_full_melspec400.txt 2>&1
{ python tools/ncu_summary.py $M/full512.ncu-rep; python tools/ncu_smem.py $M/full512.ncu-rep 256000 | tail -12; python tools/ncu_ophist.py $M/full512.ncu-rep 256000; } > profiles/${T}_ncu_full_melspec512_kaldi_cmn.txt 2>&1
{ echo "# sanitizer runs of the final build of the round (tools/measure_round.sh full; the test selection is in that script)"; for S in memcheck racecheck synccheck; do echo "== $S"; grep -E "COMPUTE-SANITIZER|passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" $M/san_$S.log | sort -u; done; } > profiles/${T}_sanitizers.txt
