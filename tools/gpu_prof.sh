#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests -m gpu -x -q > $O/t_gpu_r13.log 2>&1; echo "rc=$?" >> $O/t_gpu_r13.log; tail -4 $O/t_gpu_r13.log
for mb in 1 2 4; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-table --micro-batches $mb > $O/bench_mb$mb.json 2> $O/bench_mb$mb.err
  python -c "import json; d=json.loads(open('$O/bench_mb$mb.json').read().strip().splitlines()[-1]); print('mb=$mb', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"
done
echo "all done $(( $(date +%s) - T0 )) s"
