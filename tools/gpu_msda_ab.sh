#!/bin/bash
# micro A/B of one MSDA env knob: gpu_msda_ab.sh KNOB "v1 v2 ..."
KNOB=$1; VALS=$2
for V in $VALS; do
  env $KNOB=$V timeout 120 python tools/msda_micro.py ab cfg2 1.0 2>&1 | tail -1 | cut -c1-200
done
