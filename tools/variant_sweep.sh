#!/bin/bash
# On the GPU box: rebuild the library with each set of extra nvcc flags given as arguments, run a parity subset and the bench
# usage: variant_sweep.sh "<pytest -k expression>" "<flags 1>" "<flags 2>" ...
set -u
K="$1"; shift
for v in "$@"; do
  SB_NVCC_EXTRA="$v" python -m stark_pure_rust_b200.build --force > /dev/null 2> gpurun_out/build_variant.err || { echo "build failed: $v"; continue; }
  python -m pytest tests/test_gpu_parity.py -q -x -k "$K" 2>&1 | tail -1
  python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-prove --no-sweep > gpurun_out/variant.json 2> /dev/null
  python - <<PY
import json
d = json.loads(open("gpurun_out/variant.json").read())
b = d["breakdown"]
print("[$v] step %.2f ms | lde %.2f ntt_pass %.2f | merkle8 %.2f merkle1 %.2f fri %.2f | leaves %.2f nodes %.2f fold %.2f | executed_frac %.3f merkle_alu_frac %.3f" % (
    d["ms_per_step"], b["lde_ms"], b["kernel_ms_per_step"]["ntt_pass"], b["merkle8_ms"], b["merkle1_ms"], b["fri_ms"],
    b["kernel_ms_per_step"]["merkle_leaves"], b["kernel_ms_per_step"]["merkle_nodes"], b["kernel_ms_per_step"]["fri_fold"], d["int_pipe"]["executed_frac"], b["merkle_alu_frac"]))
PY
done
python -m stark_pure_rust_b200.build --force > /dev/null 2>&1
