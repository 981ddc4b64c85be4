#!/bin/bash
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:merkle_leaves_cols -s 2 -c 2 -o gpurun_out/prof_merkle_r01 $CMD > gpurun_out/ncu_full2.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/ncu_full2.log
