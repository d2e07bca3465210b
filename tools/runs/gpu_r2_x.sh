#!/bin/bash
# round 2, GPU call X (8 GPUs): bench.py --gpus 8 of the final state (crowd strong scaling + strand_sharded), then --gpus 4 on the same box
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_x_8gpu.json 2> gpurun_out/bench_x_8gpu.err; tail -4 gpurun_out/bench_x_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_x_4gpu.json 2> gpurun_out/bench_x_4gpu.err; tail -4 gpurun_out/bench_x_4gpu.err
python - <<'PY'
import json
for n in (8,4):
    d=json.loads(open('gpurun_out/bench_x_%dgpu.json'%n).read().strip().splitlines()[-1])
    print(n, 'ms/step %.4f'%d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
    print('   sharded', json.dumps(d['strand_sharded'])[:1500])
PY
