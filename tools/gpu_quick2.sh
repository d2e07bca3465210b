#!/bin/bash
mkdir -p gpurun_out
for g in 64 2048; do for m in full atom; do
VKHR_B200_GROUP_MIB=$g VKHR_B200_DEBUG_SINK=$m timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
python -c "import json;d=json.load(open('gpurun_out/bench_$m.json'));print('$m', $g, d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
done; done
