#!/bin/bash
# ncu --set full of the tile MSDA backward inside one eager step + kernel table of a bench run.  Output in gpurun_out/.
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-tp}; T0=$(date +%s)
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:msda_bwd -s 1 -c 2 -o $O/prof_${TAG}_msda_bwd -f python tools/profile_step.py > $O/ncu_${TAG}.log 2>&1
echo "ncu done $(( $(date +%s) - T0 )) s"
for V in 1 0; do
  POET_MSDA_TILE=$V timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${TAG}_$V.json 2> $O/bench_${TAG}_$V.err; echo "bench tile=$V rc=$?"
  python tools/show_bench.py $O/bench_${TAG}_$V.json 8 2>/dev/null | tail -9 | cut -c1-200
done
echo "all done $(( $(date +%s) - T0 )) s"
