#!/bin/bash
# A/B bench lines under environment variants: tools/gpu_ab.sh TAG "VAR=val VAR2=val" "..." 
mkdir -p gpurun_out; O=gpurun_out; TAG=$1; shift
i=0
for V in "$@"; do
  env $V timeout 300 python bench.py --steps 20 --warmup 5 --no-kernel-table > $O/ab_${TAG}_$i.json 2> $O/ab_${TAG}_$i.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/ab_${TAG}_$i.json") if l.startswith("{")][-1]); print("[$V] ms_per_step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
except Exception as e: print("[$V] failed", e)
PY
  i=$((i+1))
done
