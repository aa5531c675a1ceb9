#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > $O/t_gpu_r21.log 2>&1; echo "rc=$?" >> $O/t_gpu_r21.log; tail -3 $O/t_gpu_r21.log
for pdl in 1 0; do
POET_PDL=$pdl timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-kernel-table > $O/bench_pdl$pdl.json 2> $O/bench_pdl$pdl.err
python -c "import json; d=json.loads([l for l in open('$O/bench_pdl$pdl.json') if l.startswith('{')][-1]); print('pdl=$pdl', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"
tail -2 $O/bench_pdl$pdl.err
done
POET_PDL=1 timeout 200 python tools/kernel_micro.py pdl1 2>&1 | grep -E "decoder|proj fwd|LN"
POET_PDL=0 timeout 200 python tools/kernel_micro.py pdl0 2>&1 | grep -E "decoder|proj fwd|LN"
