#!/bin/bash
mkdir -p gpurun_out
for m in full null compact quarter; do
VKHR_B200_DEBUG_SINK=$m timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
python -c "import json;d=json.load(open('gpurun_out/bench_$m.json'));print('$m', d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
done
