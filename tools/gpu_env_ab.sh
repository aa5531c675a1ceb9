#!/bin/bash
# A/B of one environment knob inside the replayed step: gpu_env_ab.sh TAG KNOB "v1 v2 ..." [--with-tests]
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-ab}; KNOB=$2; VALS=$3; T0=$(date +%s)
if [ "$4" = "--with-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/t_gpu_$TAG.log 2>&1; echo "rc=$?" >> $O/t_gpu_$TAG.log; tail -3 $O/t_gpu_$TAG.log | cut -c1-200
fi
for V in $VALS; do
  for REP in 1 2; do
    env $KNOB=$V timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table --no-parity > $O/bench_${TAG}_${V}_$REP.json 2> $O/bench_${TAG}_${V}_$REP.err
    python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_${TAG}_${V}_$REP.json") if l.startswith("{")][-1]); print("$KNOB=$V rep $REP: ms_per_step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
except Exception as e: print("failed", e)
PY
  done
done
echo "all done $(( $(date +%s) - T0 )) s"
