#!/bin/bash
# the round's last GPU call, most important artefact first (the budget may cut it short)
mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 15 -c 15 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/ncu_bench.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_walk_uniform -s 3 -c 1 -o gpurun_out/prof_walk64 -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/ncu_full.log 2>&1
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
