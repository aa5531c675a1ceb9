#!/bin/bash
# A/B of the small-row GEMM kernel: parity suite with it on, bench lines with it on and off.
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-k}; T0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q > $O/t_gpu_$TAG.log 2>&1; echo "rc=$?" >> $O/t_gpu_$TAG.log
tail -15 $O/t_gpu_$TAG.log | cut -c1-400
echo "tests done $(( $(date +%s) - T0 )) s"
for V in 1 0; do
  POET_GEMM_SMALL=$V timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_${TAG}_small$V.json 2> $O/bench_${TAG}_small$V.err; echo "bench small=$V rc=$?"
  python tools/show_bench.py $O/bench_${TAG}_small$V.json 24 2>/dev/null
done
echo "all done $(( $(date +%s) - T0 )) s"
