#!/bin/bash
# sharded modes only: usage run_scale2.sh "<list of N>" <log_n> <cols> <tag>
NS=${1:-"4 8"}; L=${2:-26}; C=${3:-8}; TAG=${4:-r01b}
for N in $NS; do
  if [ "$N" = "1" ]; then LAUNCH="python"; else LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800+N))"; fi
  $LAUNCH bench.py --gpus $N --mode sharded --log-n $L --cols $C --steps 3 --warmup 2 > gpurun_out/scale_sharded_${TAG}_L${L}_n$N.json 2> gpurun_out/scale_sharded_${TAG}_L${L}_n$N.err
  echo "sharded N=$N rc=$?"
  $LAUNCH bench.py --gpus $N --mode sharded-ntt --log-n $L --steps 3 --warmup 2 > gpurun_out/scale_sharded_ntt_${TAG}_L${L}_n$N.json 2> gpurun_out/scale_sharded_ntt_${TAG}_L${L}_n$N.err
  echo "sharded-ntt N=$N rc=$?"
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/scale_sharded*_r01b_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n_gpus", d["n_gpus"], "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"])
    except Exception as e:
        print(f, "unreadable", e)
PY
