#!/bin/bash
# Round-2 gpurun call: GPU parity suite, default bench line, optional extras selected by words in $2..
#   tools/gpu_r02.sh TAG [tests] [bench] [launches] [micro] [ncu_msda] [ncu_gemm] [cfgs] [n2]
mkdir -p gpurun_out
O=gpurun_out
TAG=${1:-a}; shift
WHAT=" ${*:-tests bench} "
T0=$(date +%s)
has() { [[ "$WHAT" == *" $1 "* ]]; }
if has tests; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $O/t_gpu_$TAG.log 2>&1; echo "rc=$?" >> $O/t_gpu_$TAG.log
  tail -15 $O/t_gpu_$TAG.log | cut -c1-400
  echo "tests done $(( $(date +%s) - T0 )) s"
fi
if has smoke; then
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke_$TAG.log
fi
if has micro; then
  timeout 300 python tools/kernel_micro.py $TAG > $O/micro_$TAG.txt 2>&1; cat $O/micro_$TAG.txt
fi
if has bench; then
  timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; tail -3 $O/bench_$TAG.err
  python tools/show_bench.py $O/bench_$TAG.json 18 2>/dev/null
fi
if has launches; then
  timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$TAG.csv python tools/profile_step.py > $O/ncu_launch_$TAG.log 2>&1
  python tools/launch_shares.py $O/launches_$TAG.csv | head -40
fi
if has ncu_msda; then
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"msda_fwd_slab|msda_bwd" -c 6 -o $O/prof_${TAG}_msda -f python tools/profile_step.py > $O/ncu_msda_$TAG.log 2>&1; echo "ncu msda rc=$?"
fi
if has ncu_gemm; then
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -s 12 -c 16 -o $O/prof_${TAG}_gemm -f python tools/profile_step.py > $O/ncu_gemm_$TAG.log 2>&1; echo "ncu gemm rc=$?"
fi
if has cfgs; then
  for W in cfg1 cfg3 cfg5; do
    timeout 900 python bench.py --workload $W --steps 10 --warmup 3 --no-kernel-table > $O/bench_${TAG}_$W.json 2> $O/bench_${TAG}_$W.err; echo "$W rc=$?"
    python tools/show_bench.py $O/bench_${TAG}_$W.json 0 2>/dev/null | cut -c1-400
  done
fi
if has n2 && [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q > $O/t_dist_$TAG.log 2>&1; echo "dist rc=$?"; tail -5 $O/t_dist_$TAG.log | cut -c1-300
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-kernel-table > $O/bench_${TAG}_n2.json 2> $O/bench_${TAG}_n2.err; echo "n2 rc=$?"
  python tools/show_bench.py $O/bench_${TAG}_n2.json 0 2>/dev/null | cut -c1-400
fi
echo "all done $(( $(date +%s) - T0 )) s"
