#!/bin/bash
# ncu evidence for the bench command (round 1): launch list + one full capture of the dominant kernel
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 8 -c 3 -o gpurun_out/prof_ntt_r01 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/ncu_full.log
