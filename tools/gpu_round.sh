#!/bin/bash
# One gpurun call: GPU parity tests, bench (default + A/B env knobs), micro-benchmarks, ncu captures.
# Everything lands in gpurun_out/.  Each stage is wrapped in its own timeout.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
T0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q > $O/t_gpu.log 2>&1; echo "rc=$?" >> $O/t_gpu.log
tail -5 $O/t_gpu.log
if ! grep -q "^rc=0" $O/t_gpu.log; then
  timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q > $O/t_ops_all.log 2>&1; echo "rc=$?" >> $O/t_ops_all.log
  POET_GEMM_TMA_EPI=0 timeout 400 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "gemm or linear" > $O/t_ops_noepi.log 2>&1; echo "rc=$?" >> $O/t_ops_noepi.log
  POET_GEMM_WGRAD_BN=128 POET_GEMM_L2_PREFETCH=0 timeout 400 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "gemm or linear" > $O/t_ops_bn128.log 2>&1; echo "rc=$?" >> $O/t_ops_bn128.log
  POET_MSDA_SLAB=0 timeout 400 python -m pytest tests/test_gpu_ops.py -m gpu -q -k msda > $O/t_ops_noslab.log 2>&1; echo "rc=$?" >> $O/t_ops_noslab.log
  tail -3 $O/t_ops_all.log $O/t_ops_noepi.log $O/t_ops_bn128.log $O/t_ops_noslab.log
fi
echo "tests done $(( $(date +%s) - T0 )) s"
timeout 240 python tools/kernel_micro.py default > $O/micro_default.txt 2>&1
POET_GEMM_TMA_EPI=0 timeout 240 python tools/kernel_micro.py no_tma_epi > $O/micro_noepi.txt 2>&1
POET_GEMM_L2_PREFETCH=0 POET_GEMM_WGRAD_BN=128 POET_MSDA_SLAB=0 timeout 240 python tools/kernel_micro.py nopf_bn128_noslab > $O/micro_nopf.txt 2>&1
cat $O/micro_default.txt
echo "micro done $(( $(date +%s) - T0 )) s"
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python tools/show_bench.py $O/bench_default.json 2>/dev/null | head -40
POET_GEMM_TMA_EPI=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-table > $O/bench_noepi.json 2> $O/bench_noepi.err
POET_GEMM_L2_PREFETCH=0 POET_GEMM_WGRAD_BN=128 POET_MSDA_SLAB=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-table > $O/bench_nopf.json 2> $O/bench_nopf.err
for f in $O/bench_noepi.json $O/bench_nopf.json; do python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'])"; done
echo "bench done $(( $(date +%s) - T0 )) s"
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r01b.csv python tools/profile_step.py > $O/ncu_launch.log 2>&1
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"msda_fwd_slab|msda_bwd" -c 3 -o $O/prof_r01b_msda -f python tools/profile_step.py > $O/ncu_msda.log 2>&1
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -s 10 -c 14 -o $O/prof_r01b_gemm -f python tools/profile_step.py > $O/ncu_gemm.log 2>&1
echo "all done $(( $(date +%s) - T0 )) s"
