#!/bin/bash
mkdir -p gpurun_out
for m in full quarter null atom; do
echo "== $m"
VKHR_B200_GROUP_MIB=2048 VKHR_B200_DEBUG_SINK=$m python tools/cta_trace.py 4 | head -6
cp gpurun_out/cta_trace.json gpurun_out/cta_trace_$m.json
done
