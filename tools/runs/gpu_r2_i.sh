#!/bin/bash
# round 2, GPU call I: sharded C++ test diagnosis, prefilter with dynamic tiles, full bench line
mkdir -p gpurun_out
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 120 ./adapter/_build/sharded_test > gpurun_out/sharded_test.log 2>&1; tail -12 gpurun_out/sharded_test.log
timeout 900 python -m pytest tests/test_prefilter_gpu.py tests/test_parity_gpu.py -m gpu -q -k "prefilter or sparse or sharded or tangent or frame" 2>&1 | tail -12 > gpurun_out/pytest_gpu_i.log; tail -5 gpurun_out/pytest_gpu_i.log
timeout 900 python bench.py > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; tail -c 2500 gpurun_out/bench_i.json; tail -3 gpurun_out/bench_i.err
