#!/bin/bash
# what the driver does at round end (tests, smoke, both bench arms) + the ncu evidence for profiles/
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 12 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk_uniform -s 3 -c 1 -o gpurun_out/prof_walk64 -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | head -40
timeout 600 ncu --set full --clock-control none -k regex:k_untile_batch -s 3 -c 1 -o gpurun_out/prof_untile64 -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/ncu_untile.log 2>&1
