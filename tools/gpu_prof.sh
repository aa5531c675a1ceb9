#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "input_proj" > $O/t_inproj.log 2>&1; echo "rc=$?" >> $O/t_inproj.log; tail -30 $O/t_inproj.log
