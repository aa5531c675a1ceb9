#!/bin/bash
# 2-GPU call: distributed parity test, then bench lines with the in-graph overlapped all-reduce and the blocking one.
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-n2}; N=${2:-2}
timeout 500 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q > $O/t_dist_$TAG.log 2>&1; echo "dist rc=$?"; tail -4 $O/t_dist_$TAG.log | cut -c1-300
for V in "" "--overlap-allreduce"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-kernel-table $V > $O/bench_${TAG}_n$N$V.json 2> $O/bench_${TAG}_n$N$V.err; echo "bench $V rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_${TAG}_n$N$V.json") if l.startswith("{")][-1]); print("[$N GPUs $V] ms_per_step", round(d["ms_per_step"],3), "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["config"].get("grad_allreduce"))
except Exception as e: print("failed", e)
PY
  tail -2 $O/bench_${TAG}_n$N$V.err | cut -c1-300
done
