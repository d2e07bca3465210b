#!/bin/bash
# where does the walk's time go?  probes with the same instruction stream and different L1TEX wavefront counts
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-others $EXTRA > gpurun_out/ab3_$name.json 2> gpurun_out/ab3_$name.err
  python -c "import json;d=json.load(open('gpurun_out/ab3_$name.json'));print('$name', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['phase_ms_per_step'].items() if v})" || tail -3 gpurun_out/ab3_$name.err
}
for v in null sector line lines8 red atom; do
  EXTRA="" run ${v}_64 VKHR_B200_WALK=$v
  EXTRA="--instances 2" run ${v}_2 VKHR_B200_WALK=$v
done
