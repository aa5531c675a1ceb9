#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "backbone_mode" > $O/t_n4.log 2>&1; echo "rc=$?" >> $O/t_n4.log; tail -25 $O/t_n4.log
