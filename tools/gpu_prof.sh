#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > $O/t_gpu_r31.log 2>&1; echo "rc=$?" >> $O/t_gpu_r31.log; tail -3 $O/t_gpu_r31.log
timeout 200 python tools/kernel_micro.py tail1 2>&1 | grep -E "^gemm"
for ts in 1 0; do
POET_GEMM_TAIL_SPLIT=$ts timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-kernel-table > $O/bench_ts$ts.json 2> $O/bench_ts$ts.err
python -c "import json; d=json.loads([l for l in open('$O/bench_ts$ts.json') if l.startswith('{')][-1]); print('tail_split=$ts', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"
done
