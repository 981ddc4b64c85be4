#!/bin/bash
# multi-GPU bench lines: replicas (weak scaling, the driver's default) and sharded (one job, strong scaling)
# usage: run_scale.sh "<list of N>" <sharded log_n> <sharded cols> <tag>
NS=${1:-"1 2"}; L=${2:-26}; C=${3:-8}; TAG=${4:-r01}
for N in $NS; do
  if [ "$N" = "1" ]; then LAUNCH="python"; else LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N))"; fi
  $LAUNCH bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-prove --no-sweep > gpurun_out/scale_replicas_${TAG}_n$N.json 2> gpurun_out/scale_replicas_${TAG}_n$N.err
  echo "replicas N=$N rc=$?"; tail -c 400 gpurun_out/scale_replicas_${TAG}_n$N.json | head -c 10 >/dev/null
  $LAUNCH bench.py --gpus $N --mode sharded --log-n $L --cols $C --steps 3 --warmup 2 > gpurun_out/scale_sharded_${TAG}_L${L}_n$N.json 2> gpurun_out/scale_sharded_${TAG}_L${L}_n$N.err
  echo "sharded N=$N rc=$?"
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/scale_*_n*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "n_gpus", d["n_gpus"], "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], d["breakdown"].get("exchange_ms"), d["breakdown"].get("exchange_gbs_rank0"))
    except Exception as e:
        print(f, "unreadable", e)
PY
