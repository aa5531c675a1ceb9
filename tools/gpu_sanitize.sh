#!/bin/bash
# compute-sanitizer over the op-level GPU tests (memcheck: everything but the full-size shapes; racecheck: the kernels with
# shared-memory hand-offs).  Output -> gpurun_out/sanitizer_*_$TAG.log
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-r02}
timeout 1100 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_ops.py -m gpu -q -x \
   -k "not full_size and not 25600 and not subprocess and not thread_kernels" > $O/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/sanitizer_memcheck_$TAG.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_ops.py -m gpu -q -x \
   -k "gemm_small or (tcgen05 and 3200) or msda_block or (msda_core and 500) or (msda_core and 700) or add_layernorm or mha_smallq or block_entry" > $O/sanitizer_racecheck_$TAG.log 2>&1; echo "racecheck rc=$?"; tail -4 $O/sanitizer_racecheck_$TAG.log | cut -c1-200
