#!/bin/bash
# ncu evidence for profiles/: launch list of one eager step + --set full captures of the top kernels.
mkdir -p gpurun_out
O=gpurun_out
R=${1:-r01c}
T0=$(date +%s)
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$R.csv python tools/profile_step.py > $O/ncu_launch_$R.log 2>&1
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:msda_fwd_slab -c 2 -o $O/prof_${R}_msda -f python tools/profile_step.py > $O/ncu_msda_$R.log 2>&1
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:msda_bwd_kernel -c 2 -o $O/prof_${R}_msda_bwd -f python tools/profile_step.py > $O/ncu_msda_bwd_$R.log 2>&1
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -s 12 -c 16 -o $O/prof_${R}_gemm -f python tools/profile_step.py > $O/ncu_gemm_$R.log 2>&1
timeout 200 python tools/kernel_micro.py $R > $O/micro_$R.txt 2>&1
echo "all done $(( $(date +%s) - T0 )) s"
