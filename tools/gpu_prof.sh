#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_optim.py tests/test_gpu_model.py -m gpu -x -q > $O/t_optim.log 2>&1; echo "rc=$?" >> $O/t_optim.log; tail -30 $O/t_optim.log
