#!/bin/bash
# red / atom with the 64 instances writing into 1, 2, 4 shared volumes (L2-warm targets)
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-others $EXTRA > gpurun_out/ab4_$name.json 2> gpurun_out/ab4_$name.err
  python -c "import json;d=json.load(open('gpurun_out/ab4_$name.json'));print('$name', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['phase_ms_per_step'].items() if v})" || tail -3 gpurun_out/ab4_$name.err
}
for v in red atom; do for a in 1 2 4 8; do EXTRA="" run ${v}_alias$a VKHR_B200_WALK=$v VKHR_B200_ALIAS=$a; done; done
