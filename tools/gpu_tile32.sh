#!/bin/bash
# D = 32 tile backward: op parity (ld = 3 shapes forced onto the tile kernel too), REF-pyramid 8-head micro A/B
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-t32}; T0=$(date +%s)
POET_MSDA_TILE_MAX_LD=3 timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "msda" > $O/t_msda_${TAG}.log 2>&1; echo "rc=$?" >> $O/t_msda_${TAG}.log
tail -3 $O/t_msda_${TAG}.log | cut -c1-300
for V in 1 0; do
  POET_MSDA_TILE=$V timeout 120 python tools/msda_micro.py $TAG ref8 1.0 2>&1 | tail -1
  POET_MSDA_TILE=$V timeout 120 python tools/msda_micro.py $TAG ref8 0.2 2>&1 | tail -1
  POET_MSDA_TILE=$V timeout 120 python tools/msda_micro.py $TAG cfg5 1.0 2>&1 | tail -1
done
echo "all done $(( $(date +%s) - T0 )) s"
