#!/bin/bash
# does an L2-warm (just-cleared) volume make the atomics faster?  chunked clear/walk/finish groups, and small crowds
mkdir -p gpurun_out
run() { # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-others $EXTRA > gpurun_out/ab2_$name.json 2> gpurun_out/ab2_$name.err
  python -c "import json;d=json.load(open('gpurun_out/ab2_$name.json'));print('$name', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['phase_ms_per_step'].items() if v})" || tail -3 gpurun_out/ab2_$name.err
}
for v in atom red; do
  for c in 2 4 8 16; do EXTRA="" run ${v}_chunk$c VKHR_B200_WALK=$v VKHR_B200_CHUNK=$c; done
done
for v in atom red; do
  for i in 2 4 8 16; do EXTRA="--instances $i" run ${v}_inst$i VKHR_B200_WALK=$v; done
done
