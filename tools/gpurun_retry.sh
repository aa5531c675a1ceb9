#!/bin/bash
# gpurun with retries on transient "no box / draining" answers: tools/gpurun_retry.sh LOG [gpurun args...]
LOG=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient\|status=busy" "$LOG" || grep -q "retry in a few minutes" "$LOG"; then sleep 90; continue; fi
  break
done
tail -60 "$LOG"
