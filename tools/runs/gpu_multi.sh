#!/bin/bash
# N-GPU checks (gpurun --gpus N): strand-sharded parity + timings, and the crowd bench launched as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/sharded_check.py > gpurun_out/sharded_$N.json 2> gpurun_out/sharded_$N.err; tail -1 gpurun_out/sharded_$N.json | cut -c1-1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; tail -1 gpurun_out/bench_$N.json | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29546 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_$N.json 2> gpurun_out/bench_ref_$N.err; tail -1 gpurun_out/bench_ref_$N.json | cut -c1-300
