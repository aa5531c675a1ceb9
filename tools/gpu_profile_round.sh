#!/bin/bash
# ncu evidence for profiles/: launch list of one eager step + --set full captures of the top kernels, summarised on the box
# (gpurun copies back at most 64 MiB: the big reports are reduced to text there and deleted).
mkdir -p gpurun_out
O=gpurun_out
R=${1:-r01c}
T0=$(date +%s)
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$R.csv python tools/profile_step.py > $O/ncu_launch_$R.log 2>&1
python tools/launch_shares.py $O/launches_$R.csv > $O/launch_shares_$R.txt
echo "launch list $(( $(date +%s) - T0 )) s"
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:msda_fwd_slab -c 1 -o $O/prof_${R}_msda -f python tools/profile_step.py > $O/ncu_msda_$R.log 2>&1
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:msda_bwd -s 1 -c 1 -o $O/prof_${R}_msda_bwd -f python tools/profile_step.py > $O/ncu_msda_bwd_$R.log 2>&1
timeout 500 ncu --profile-from-start off --set full --clock-control none -k regex:gemm_tc -s 12 -c 16 -o $O/prof_${R}_gemm -f python tools/profile_step.py > $O/ncu_gemm_$R.log 2>&1
echo "ncu full $(( $(date +%s) - T0 )) s"
( python tools/ncu_summary.py $O/prof_${R}_msda.ncu-rep; python tools/ncu_summary.py $O/prof_${R}_msda_bwd.ncu-rep ) > $O/ncu_msda_summary_$R.txt 2>&1
python tools/ncu_summary.py $O/prof_${R}_gemm.ncu-rep > $O/ncu_gemm_summary_$R.txt 2>&1
python tools/make_traffic.py $R $O/prof_${R}_msda.ncu-rep $O/prof_${R}_msda_bwd.ncu-rep > $O/traffic_$R.json 2>&1
python tools/ncu_hotspots.py $O/prof_${R}_msda_bwd.ncu-rep 0 40 > $O/ncu_msda_bwd_hotspots_$R.txt 2>&1
rm -f $O/prof_${R}_gemm.ncu-rep $O/prof_${R}_msda.ncu-rep
du -sh $O
echo "all done $(( $(date +%s) - T0 )) s"
