#!/bin/bash
# experiment round: parity, then the crowd bench at several L2 group sizes
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
for g in 16 32 64 96 2048; do
  VKHR_B200_GROUP_MIB=$g timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_g$g.json 2> gpurun_out/bench_g$g.err
  python -c "import json;d=json.load(open('gpurun_out/bench_g$g.json'));print($g, d['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
done
timeout 600 python tools/perf_probe.py --out gpurun_out/probe.json > gpurun_out/probe.log 2>&1; cat gpurun_out/probe.log
