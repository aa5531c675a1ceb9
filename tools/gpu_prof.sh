#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
which compute-sanitizer || ls /usr/local/cuda/bin | grep -i sanit
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "msda_block or relu_bitmask or mha or layernorm or heads or wgrad_accumulate or posenc or flatten" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/sanitizer_memcheck.log
tail -15 $O/sanitizer_memcheck.log
echo "memcheck done $(( $(date +%s) - T0 )) s"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "msda_block or relu_bitmask" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck.log
tail -15 $O/sanitizer_racecheck.log
echo "all done $(( $(date +%s) - T0 )) s"
