#!/bin/bash
# Tile MSDA kernels: op parity, A/B micro timings and short bench lines.  usage: gpu_tile.sh TAG ENVVAR   (ENVVAR=1 / 0 is compared)
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-tile}; KNOB=${2:-POET_MSDA_FWD_TILE}; T0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "msda" > $O/t_msda_${TAG}.log 2>&1; echo "rc=$?" >> $O/t_msda_${TAG}.log
tail -3 $O/t_msda_${TAG}.log | cut -c1-300
for V in 1 0; do
  env $KNOB=$V timeout 120 python tools/msda_micro.py $TAG cfg2 1.0 2>&1 | tail -1
  env $KNOB=$V timeout 120 python tools/msda_micro.py $TAG cfg2 0.2 2>&1 | tail -1
  env $KNOB=$V timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${TAG}_$V.json 2> $O/bench_${TAG}_$V.err; echo "bench $KNOB=$V rc=$?"
  python tools/show_bench.py $O/bench_${TAG}_$V.json 2 2>/dev/null | cut -c1-250
done
echo "all done $(( $(date +%s) - T0 )) s"
