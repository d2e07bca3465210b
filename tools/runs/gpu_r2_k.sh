#!/bin/bash
# round 2, GPU call K (1 GPU): the sharded fake-rank tests FIRST in a fresh process (nothing has loaded a kernel yet), repeated
mkdir -p gpurun_out
for i in 1 2 3 4 5; do CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 120 ./adapter/_build/sharded_test > gpurun_out/sharded_test_$i.log 2>&1; tail -1 gpurun_out/sharded_test_$i.log; done
for i in 1 2 3; do timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "sharded" 2>&1 | tail -2; done > gpurun_out/pytest_gpu_k.log; cat gpurun_out/pytest_gpu_k.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_k_all.log; cat gpurun_out/pytest_gpu_k_all.log
