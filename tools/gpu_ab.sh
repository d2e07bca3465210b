#!/bin/bash
# A/B of the walk variants on the crowd bench (device-resident timing only) + the GPU parity tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
for v in red atom v5; do
  for ni in 0 1; do
    [ "$v" = v5 ] && [ $ni = 1 ] && continue
    VKHR_B200_WALK=$v VKHR_B200_NO_INT_INDEX=$ni timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/ab_${v}_$ni.json 2> gpurun_out/ab_${v}_$ni.err
    python -c "import json;d=json.load(open('gpurun_out/ab_${v}_$ni.json'));print('$v noint=$ni', round(d['ms_per_step'],4), d['roofline']['phase_ms_per_step'])" || tail -3 gpurun_out/ab_${v}_$ni.err
  done
done
