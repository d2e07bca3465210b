#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 ./tools/microbench2 > gpurun_out/microbench2.json 2>&1; cat gpurun_out/microbench2.json
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python -c "import json;d=json.load(open('gpurun_out/bench_quick.json'));print('full', d['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
VKHR_B200_DEBUG_SINK=null timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_null.json 2> gpurun_out/bench_null.err
python -c "import json;d=json.load(open('gpurun_out/bench_null.json'));print('null', d['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
tail -3 gpurun_out/bench_null.err
