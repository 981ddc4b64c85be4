#!/bin/bash
set -x
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --log-n 20 --cpu-log-n 16 > gpurun_out/bench_l20.json 2> gpurun_out/bench_l20.err; tail -5 gpurun_out/bench_l20.err; cat gpurun_out/bench_l20.json
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_l24.json 2> gpurun_out/bench_l24.err; tail -5 gpurun_out/bench_l24.err; cat gpurun_out/bench_l24.json
