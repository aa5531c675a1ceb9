#!/bin/bash
# Marginal in-graph cost of one encoder / decoder layer: bench with layer-count overrides (never bench lines).
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-abl}
for V in "5 5" "4 5" "5 4" "5 1" "1 5"; do
  set -- $V
  POET_BENCH_ENC_LAYERS=$1 POET_BENCH_DEC_LAYERS=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-kernel-table > $O/abl_${TAG}_$1_$2.json 2> $O/abl_${TAG}_$1_$2.err
  python - <<PY
import json
d=json.loads([l for l in open("$O/abl_${TAG}_$1_$2.json") if l.startswith("{")][-1])
print("enc $1 dec $2: ms_per_step", round(d["ms_per_step"],3))
PY
done
