#!/bin/bash
# round 2, final 8-GPU run: bench.py --gpus 8 (crowd strong scaling + strand_sharded) with the final kernels
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_final_8gpu.json 2> gpurun_out/bench_final_8gpu.err; tail -3 gpurun_out/bench_final_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final_8gpu.json').read().strip().splitlines()[-1])
print(8, 'ms/step %.4f'%d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
print('   sharded', json.dumps(d['strand_sharded'])[:1500])
PY
