#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_criterion.py tests/test_gpu_optim.py -m gpu -x -q > $O/t_crit.log 2>&1; echo "rc=$?" >> $O/t_crit.log; tail -25 $O/t_crit.log
