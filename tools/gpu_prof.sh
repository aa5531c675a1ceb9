#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "gemm or linear" > $O/t_gemm_r8.log 2>&1; echo "rc=$?" >> $O/t_gemm_r8.log; tail -3 $O/t_gemm_r8.log
for dbg in 0 8 16 15; do POET_GEMM_DEBUG=$dbg python tools/gemm_bisect.py child 25600x1024x256 25600x256x256 25600x256x1024 25600x768x256; done > $O/gemm_bisect_r8.txt 2>&1
cat $O/gemm_bisect_r8.txt
timeout 240 python tools/kernel_micro.py r8 > $O/micro_r8.txt 2>&1
cat $O/micro_r8.txt
echo "all done $(( $(date +%s) - T0 )) s"
