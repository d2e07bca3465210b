#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-others $EXTRA > gpurun_out/ab7_$name.json 2> gpurun_out/ab7_$name.err
  python -c "import json;d=json.load(open('gpurun_out/ab7_$name.json'));print('$name', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['phase_ms_per_step'].items() if v})" || tail -3 gpurun_out/ab7_$name.err
}
EXTRA="" run pipe2_64 VKHR_B200_WALK=red
EXTRA="" run pipe2_alias2 VKHR_B200_WALK=red VKHR_B200_ALIAS=2
EXTRA="--instances 8" run pipe2_8 VKHR_B200_WALK=red
VKHR_B200_WALK=red timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
