#!/bin/bash
# On the GPU box: rebuild the library with different tile load / store unroll factors and time the LDE each time
set -u
for v in "2 2" "1 2" "2 1" "1 1" "2 2 -DNTT_NOINLINE_MUL"; do
  set -- $v
  SB_NVCC_EXTRA="-DNTT_LOAD_UNROLL=$1 -DNTT_STORE_UNROLL=$2 ${3:-}" python -m stark_pure_rust_b200.build --force > /dev/null 2> gpurun_out/build_u$1_$2.err || { echo "build failed $v"; continue; }
  python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-prove --no-sweep > gpurun_out/unroll_$1_$2${3:+_ni}.json 2> /dev/null
  python - <<PY
import json
d = json.loads(open("gpurun_out/unroll_$1_$2${3:+_ni}.json").read())
b = d["breakdown"]
print("load_unroll=$1 store_unroll=$2 ${3:-}: step %.2f ms lde %.2f ms ntt_pass %.2f ms executed_frac %.3f" % (d["ms_per_step"], b["lde_ms"], b["kernel_ms_per_step"]["ntt_pass"], d["int_pipe"]["executed_frac"]))
PY
done
python -m stark_pure_rust_b200.build --force > /dev/null 2>&1
