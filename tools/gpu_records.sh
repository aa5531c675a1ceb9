#!/bin/bash
# Round records: full GPU parity suite, sanitizer over the MSDA kernels, bench lines of every BASELINE.json config on 1 GPU.
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-r02c}; T0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -x -q > $O/t_gpu_$TAG.log 2>&1; echo "rc=$?" >> $O/t_gpu_$TAG.log
tail -4 $O/t_gpu_$TAG.log | cut -c1-300
echo "tests done $(( $(date +%s) - T0 )) s"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_ops.py -m gpu -q -x \
   -k "msda and not full_size and not subprocess and not thread_kernels" > $O/sanitizer_memcheck_msda_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/sanitizer_memcheck_msda_$TAG.log | cut -c1-200
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_ops.py -m gpu -q -x \
   -k "msda_block or (msda_core and 500) or (msda_core and 333) or (msda_core and 300)" > $O/sanitizer_racecheck_msda_$TAG.log 2>&1; echo "racecheck rc=$?"; tail -3 $O/sanitizer_racecheck_msda_$TAG.log | cut -c1-200
echo "sanitizer done $(( $(date +%s) - T0 )) s"
for W in ${2:-cfg2 cfg1 cfg3 cfg4 cfg5}; do
  timeout 600 python bench.py --workload $W --steps 20 --warmup 5 > $O/bench_${TAG}_$W.json 2> $O/bench_${TAG}_$W.err; echo "bench $W rc=$?"
  python tools/show_bench.py $O/bench_${TAG}_$W.json 3 2>/dev/null | cut -c1-330
done
echo "all done $(( $(date +%s) - T0 )) s"
