#!/bin/bash
# tools/gpu_round.sh -- what one gpurun call runs: tests, smoke, microbench, probe, bench, ncu passes.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 ./tools/microbench > gpurun_out/microbench.json 2>&1
timeout 600 python tools/perf_probe.py --out gpurun_out/probe.json > gpurun_out/probe.log 2>&1; tail -3 gpurun_out/probe.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" = "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --instances 8 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk_uniform -s 2 -c 2 -o gpurun_out/prof_walk -f \
      python bench.py --steps 1 --warmup 3 --instances 8 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
fi
ls -la gpurun_out
