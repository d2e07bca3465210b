#!/bin/bash
mkdir -p gpurun_out
for lib in "" vkhr_b200/lib/libvkhr_b200_occ8.so; do
for m in full null atom; do
echo "== lib=$lib $m"
VKHR_B200_LIB=$lib VKHR_B200_GROUP_MIB=2048 VKHR_B200_DEBUG_SINK=$m python tools/cta_trace.py 4 > gpurun_out/trace_$m.log 2>&1; grep -E "records|duration|classes" gpurun_out/trace_$m.log
[ -z "$lib" ] && cp gpurun_out/cta_trace.json gpurun_out/cta_trace_$m.json
done; done
