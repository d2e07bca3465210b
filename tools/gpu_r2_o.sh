#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 120 ./adapter/_build/sharded_test > gpurun_out/sharded_test_$i.log 2>&1; tail -1 gpurun_out/sharded_test_$i.log; done
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "sharded or fake_rank" 2>&1 | tail -3
