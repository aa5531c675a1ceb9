#!/bin/bash
# N-GPU bench lines of the multi-GPU BASELINE.json configs: gpu_scale.sh TAG N "cfg5 cfg4 ..."
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-sc}; N=${2:-2}; WLS=${3:-"cfg5"}; T0=$(date +%s)
for W in $WLS; do
  if [ "$N" = "1" ]; then
    timeout 400 python bench.py --workload $W --steps 20 --warmup 5 --no-kernel-table --no-cpu-baseline > $O/bench_${TAG}_${W}_n$N.json 2> $O/bench_${TAG}_${W}_n$N.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --workload $W --gpus $N --steps 20 --warmup 5 --no-kernel-table > $O/bench_${TAG}_${W}_n$N.json 2> $O/bench_${TAG}_${W}_n$N.err
  fi
  echo "bench $W n=$N rc=$? $(( $(date +%s) - T0 )) s"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_${TAG}_${W}_n$N.json") if l.startswith("{")][-1]); print("[$W, $N GPUs] ms_per_step", round(d["ms_per_step"],3), "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "batch/gpu", d["config"]["batch_per_gpu"], d["scaling"], d.get("parity",{}).get("ok"), d["clocks"])
except Exception as e: print("failed", e)
PY
  tail -2 $O/bench_${TAG}_${W}_n$N.err | cut -c1-300
done
