#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "msda" > $O/t_msda_r5.log 2>&1; echo "rc=$?" >> $O/t_msda_r5.log; tail -3 $O/t_msda_r5.log
timeout 240 python tools/kernel_micro.py r5 2>&1 | grep -E "msda|LN" 
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"msda_fwd_slab" -s 3 -c 1 -o $O/prof_r5_msda -f python tools/kernel_micro.py prof > $O/ncu_msda_r5.log 2>&1
echo "all done $(( $(date +%s) - T0 )) s"
