#!/bin/bash
mkdir -p gpurun_out
for m in scramble scramble_resident; do
echo "== $m"
VKHR_B200_GROUP_MIB=2048 VKHR_B200_DEBUG_SINK=$m python tools/cta_trace.py 4 > gpurun_out/trace_$m.log 2>&1; grep -E "records|duration|classes" gpurun_out/trace_$m.log
done
