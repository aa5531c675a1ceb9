#!/bin/bash
# Short gpurun call: GPU parity tests, micro-benchmarks, one bench line.  Output in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
TAG=${1:-q}
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > $O/t_gpu_$TAG.log 2>&1; echo "rc=$?" >> $O/t_gpu_$TAG.log
tail -4 $O/t_gpu_$TAG.log
echo "tests done $(( $(date +%s) - T0 )) s"
timeout 240 python tools/kernel_micro.py $TAG > $O/micro_$TAG.txt 2>&1
cat $O/micro_$TAG.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
python tools/show_bench.py $O/bench_$TAG.json 16 2>/dev/null
echo "all done $(( $(date +%s) - T0 )) s"
