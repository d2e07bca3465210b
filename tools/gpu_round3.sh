#!/bin/bash
# parity + bench + probe + ncu of the walk kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python -c "import json;d=json.load(open('gpurun_out/bench_quick.json'));print(d['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
timeout 600 python tools/perf_probe.py --out gpurun_out/probe.json > gpurun_out/probe.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk_uniform -s 2 -c 1 -o gpurun_out/prof_walk -f \
    python bench.py --steps 1 --warmup 3 --instances 8 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
