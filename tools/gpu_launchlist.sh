#!/bin/bash
# launch list of one eager step (ncu, per-launch durations) -> gpurun_out/launches_$TAG.csv + shares
mkdir -p gpurun_out; O=gpurun_out; R=${1:-ll}
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$R.csv python tools/profile_step.py > $O/ncu_launch_$R.log 2>&1
python tools/launch_shares.py $O/launches_$R.csv > $O/launch_shares_$R.txt
head -3 $O/launch_shares_$R.txt
