#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "edge" > $O/t_edge.log 2>&1; echo "rc=$?" >> $O/t_edge.log; tail -25 $O/t_edge.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-kernel-table --criterion --optimizer > $O/bench_crit.json 2> $O/bench_crit.err
python -c "import json; d=json.loads([l for l in open('$O/bench_crit.json') if l.startswith('{')][-1]); print('crit+opt', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"
tail -3 $O/bench_crit.err
