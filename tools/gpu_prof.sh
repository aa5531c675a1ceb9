#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-kernel-table > $O/bench_e2e.json 2> $O/bench_e2e.err
python -c "import json; d=json.loads([l for l in open('$O/bench_e2e.json') if l.startswith('{')][-1]); print('n1', round(d['value'],1), round(d['ms_per_step'],3), d['e2e'])"
tail -3 $O/bench_e2e.err
