#!/bin/bash
# round 2, GPU call J (2 GPUs): sharded tests (repeated: the fake-rank form must not be flaky), bench.py --gpus 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_j.txt
for i in 1 2 3 4 5; do CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 120 ./adapter/_build/sharded_test > gpurun_out/sharded_test_$i.log 2>&1; tail -1 gpurun_out/sharded_test_$i.log; done
for i in 1 2 3; do timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "sharded" 2>&1 | tail -2; done > gpurun_out/pytest_gpu_j.log; cat gpurun_out/pytest_gpu_j.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_j_2gpu.json 2> gpurun_out/bench_j_2gpu.err; tail -c 3500 gpurun_out/bench_j_2gpu.json; tail -5 gpurun_out/bench_j_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_j_2gpu_ref.json 2> gpurun_out/bench_j_2gpu_ref.err; cat gpurun_out/bench_j_2gpu_ref.json | cut -c1-400
