#!/bin/bash
# round 2, GPU call C: the whole GPU suite (with the at-size parity tests) and the new bench line at N = 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_c.log; tail -4 gpurun_out/pytest_gpu_c.log
timeout 600 python bench.py > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -c 3000 gpurun_out/bench_c.json; tail -5 gpurun_out/bench_c.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c_ref.json 2> gpurun_out/bench_c_ref.err; cat gpurun_out/bench_c_ref.json
