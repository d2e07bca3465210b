#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-others $EXTRA > gpurun_out/ab5_$name.json 2> gpurun_out/ab5_$name.err
  python -c "import json;d=json.load(open('gpurun_out/ab5_$name.json'));print('$name', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['phase_ms_per_step'].items() if v})" || tail -3 gpurun_out/ab5_$name.err
}
EXTRA="" run pipe_64 VKHR_B200_WALK=red
EXTRA="" run nopipe_64 VKHR_B200_WALK=red VKHR_B200_NO_PIPELINE=1
EXTRA="" run atom_64 VKHR_B200_WALK=atom
EXTRA="--instances 8" run pipe_8 VKHR_B200_WALK=red
EXTRA="--instances 2" run pipe_2 VKHR_B200_WALK=red
EXTRA="--res 512 --instances 8" run pipe_512 VKHR_B200_WALK=red
EXTRA="--res 512 --instances 8" run atom_512 VKHR_B200_WALK=atom
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/ab5_full.json 2> gpurun_out/ab5_full.err; python -c "import json;d=json.load(open('gpurun_out/ab5_full.json'));print(d['value'], d['e2e'], d['other_configs'])"
