#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "mha" > $O/t_mha.log 2>&1; echo "rc=$?" >> $O/t_mha.log; tail -2 $O/t_mha.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2 rc=$?"
tail -3 $O/bench_n2.err
python -c "import json; d=json.loads([l for l in open('$O/bench_n2.json') if l.startswith('{')][-1]); print('n2', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks'])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-table > $O/bench_n1b.json 2> $O/bench_n1b.err
python -c "import json; d=json.loads([l for l in open('$O/bench_n1b.json') if l.startswith('{')][-1]); print('n1', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks'])"
echo "all done $(( $(date +%s) - T0 )) s"
