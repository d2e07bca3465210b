#!/bin/bash
# round 2, 2-GPU check of the final kernels: bench.py --gpus 2 (crowd strong scaling, 32 instances per GPU, + strand_sharded)
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_final_2gpu.json 2> gpurun_out/bench_final_2gpu.err; tail -2 gpurun_out/bench_final_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final_2gpu.json').read().strip().splitlines()[-1])
print(2, 'ms/step %.4f'%d['ms_per_step'], 'value', d['value'])
s=d['strand_sharded']; print({k:(v['ms'] if isinstance(v,dict) and 'ms' in v else v) for k,v in s.items() if k in ('allreduce','p2p','one_gpu')}, s['p2p'].get('byte_identical_to_one_gpu'))
PY
