#!/bin/bash
# multi-GPU bench lines: replicas (weak scaling, the driver's default), sharded (one commitment, strong scaling) and
# sharded-ntt (one transform).  usage: run_scale.sh "<list of N>" <sharded log_n> <sharded cols> <tag>
NS=${1:-"1 2"}; L=${2:-26}; C=${3:-8}; TAG=${4:-r01}
for N in $NS; do
  if [ "$N" = "1" ]; then LAUNCH="python"; else LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N))"; fi
  $LAUNCH bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-sweep > gpurun_out/scale_replicas_${TAG}_n$N.json 2> gpurun_out/scale_replicas_${TAG}_n$N.err
  echo "replicas N=$N rc=$?"
  $LAUNCH bench.py --gpus $N --mode sharded --log-n $L --cols $C --steps 3 --warmup 2 > gpurun_out/scale_sharded_${TAG}_L${L}_n$N.json 2> gpurun_out/scale_sharded_${TAG}_L${L}_n$N.err
  echo "sharded N=$N rc=$?"
  $LAUNCH bench.py --gpus $N --mode sharded-ntt --log-n $L --steps 3 --warmup 2 > gpurun_out/scale_sharded_ntt_${TAG}_L${L}_n$N.json 2> gpurun_out/scale_sharded_ntt_${TAG}_L${L}_n$N.err
  echo "sharded-ntt N=$N rc=$?"
done
python - "$TAG" <<'PY'
import glob, json, sys
for f in sorted(glob.glob("gpurun_out/scale_*_%s_*.json" % sys.argv[1])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        pr = d.get("prove") or {}
        k = [x for x in pr if x.startswith("synthetic")]
        extra = (" proofs/s %.1f" % pr[k[0]]["proofs_per_s_all_gpus"]) if k and "proofs_per_s_all_gpus" in pr[k[0]] else ""
        print(f, "n_gpus", d["n_gpus"], "ms/step %.2f" % d["ms_per_step"], "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], extra)
    except Exception as e:
        print(f, "unreadable", e)
PY
