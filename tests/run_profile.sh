#!/bin/bash
# ncu passes of one short bench run (B200_PROFILING.md recipe): launch list, then --set full captures of the dominant
# kernel (ntt_pass_kernel: the six passes of one LDE) and of the 8-column Merkle leaf kernel.  Outputs in gpurun_out/.
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-prove --no-sweep"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 6 -c 6 -f -o gpurun_out/prof_ntt_${TAG} $CMD > gpurun_out/ncu_full_ntt_${TAG}.log 2>&1
echo "ntt full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:merkle_leaves_cols -s 2 -c 2 -f -o gpurun_out/prof_merkle_${TAG} $CMD > gpurun_out/ncu_full_merkle_${TAG}.log 2>&1
echo "merkle full rc=$?"
tail -2 gpurun_out/ncu_full_ntt_${TAG}.log
