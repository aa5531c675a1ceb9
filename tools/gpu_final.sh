#!/bin/bash
# End-of-round record: GPU parity suite, default bench line (N=1) and, when two GPUs are visible, the N=2 line.
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests -m gpu -q > $O/t_gpu_final.log 2>&1; echo "rc=$?" >> $O/t_gpu_final.log; tail -3 $O/t_gpu_final.log
timeout 600 python bench.py --steps 30 --warmup 5 > $O/bench_final_n1.json 2> $O/bench_final_n1.err; echo "n1 rc=$?"
python tools/show_bench.py $O/bench_final_n1.json 8 | cut -c1-600
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 > $O/bench_final_n2.json 2> $O/bench_final_n2.err; echo "n2 rc=$?"
  python -c "import json; d=json.loads([l for l in open('$O/bench_final_n2.json') if l.startswith('{')][-1]); print('n2', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks'])"
fi
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_final_ref.json 2> $O/bench_final_ref.err; echo "ref rc=$?"; cut -c1-300 $O/bench_final_ref.json
echo "all done $(( $(date +%s) - T0 )) s"
