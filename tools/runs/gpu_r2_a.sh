#!/bin/bash
# round 2, GPU call A: parity of the instruction-diet walk, a first timing, the shared-memory atomic microbenchmark
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_a.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_a.log; tail -3 gpurun_out/pytest_gpu_a.log
timeout 300 python bench.py --no-e2e --no-cpu --no-others --steps 20 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -c 1500 gpurun_out/bench_a.json
timeout 60 ./tools/microbench3 > gpurun_out/microbench3.json 2>&1; cat gpurun_out/microbench3.json
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 15 -c 5 --csv --log-file gpurun_out/launches_a.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/ncu_a.log 2>&1
cat gpurun_out/launches_a.csv | tail -30
