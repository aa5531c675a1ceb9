#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "oracle_all_grads or golden" > $O/t_mask.log 2>&1; echo "rc=$?" >> $O/t_mask.log; tail -3 $O/t_mask.log
for fm in 1 0 1 0; do
POET_FUSE_ROW_MASK=$fm timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-kernel-table > $O/bench_fm$fm.json 2> $O/bench_fm$fm.err
python -c "import json; d=json.loads([l for l in open('$O/bench_fm$fm.json') if l.startswith('{')][-1]); print('fuse_mask=$fm', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"
done
