#!/bin/bash
# round 2, GPU call L (8 GPUs): bench.py --gpus 8 (crowd strong scaling + strand_sharded)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_l.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_l_8gpu.json 2> gpurun_out/bench_l_8gpu.err; tail -c 3800 gpurun_out/bench_l_8gpu.json; tail -4 gpurun_out/bench_l_8gpu.err
