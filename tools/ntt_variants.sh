#!/bin/bash
# One GPU call: parity and timing of the NTT pass kernel variants (SB_NTT_MAXQ = radix of the register rounds; 0 = rolled radix-2)
set -u
mkdir -p gpurun_out
for q in 0 1 2 3; do
  SB_NTT_MAXQ=$q python -m pytest tests/test_gpu_parity.py -q -x -k "best_fft or lde_batch or expand_root" > gpurun_out/variant_q${q}_pytest.log 2>&1
  echo "maxq=$q parity: $(tail -1 gpurun_out/variant_q${q}_pytest.log)"
  SB_NTT_MAXQ=$q python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-prove --no-sweep > gpurun_out/variant_q${q}_bench.json 2> gpurun_out/variant_q${q}_bench.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/variant_q${q}_bench.json").read())
    b = d["breakdown"]
    print("maxq=${q} step %.2f ms lde %.2f ms ntt_pass %.2f ms m8 %.2f m1 %.2f fri %.2f executed_frac %.3f" % (d["ms_per_step"], b["lde_ms"], b["kernel_ms_per_step"]["ntt_pass"], b["merkle8_ms"], b["merkle1_ms"], b["fri_ms"], d["int_pipe"]["executed_frac"]))
except Exception as e:
    print("maxq=${q} bench failed", e)
PY
done
