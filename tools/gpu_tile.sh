#!/bin/bash
# Tile MSDA backward: op parity, A/B micro timings and short bench lines per variant.  Output in gpurun_out/.
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-tile}; VARS=${2:-"0 1"}; T0=$(date +%s)
for V in $VARS; do
  POET_MSDA_TILE_VARIANT=$V timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "msda" > $O/t_msda_${TAG}_$V.log 2>&1; echo "rc=$?" >> $O/t_msda_${TAG}_$V.log
  tail -3 $O/t_msda_${TAG}_$V.log | cut -c1-300
  POET_MSDA_TILE_VARIANT=$V timeout 120 python tools/msda_micro.py $TAG cfg2 1.0 2>&1 | tail -1
  POET_MSDA_TILE_VARIANT=$V timeout 120 python tools/msda_micro.py $TAG cfg5 1.0 2>&1 | tail -1
  POET_MSDA_TILE_VARIANT=$V timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${TAG}_$V.json 2> $O/bench_${TAG}_$V.err; echo "bench variant=$V rc=$?"
  python tools/show_bench.py $O/bench_${TAG}_$V.json 2 2>/dev/null | cut -c1-250
done
echo "all done $(( $(date +%s) - T0 )) s"
