#!/bin/bash
# End-of-round check the way the driver does it: build + smoke, GPU parity suite, default bench line, reference arm.
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-final}; T0=$(date +%s)
timeout 600 python __graft_entry__.py --smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_$TAG.log | cut -c1-200
timeout 1200 python -m pytest tests -m gpu -x -q > $O/t_gpu_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 $O/t_gpu_$TAG.log | cut -c1-200
timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
python tools/show_bench.py $O/bench_$TAG.json 6 2>/dev/null | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "reference rc=$?"; cut -c1-300 $O/bench_ref_$TAG.json
echo "all done $(( $(date +%s) - T0 )) s"
