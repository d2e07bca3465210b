#!/bin/bash
# is the atom walk bound by atomics in flight?  occupancy variants of the same kernel (48 / 64 warps per SM)
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-others $EXTRA > gpurun_out/ab6_$name.json 2> gpurun_out/ab6_$name.err
  python -c "import json;d=json.load(open('gpurun_out/ab6_$name.json'));print('$name', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['phase_ms_per_step'].items() if v})" || tail -3 gpurun_out/ab6_$name.err
}
EXTRA="" run atom_m5 VKHR_B200_WALK=atom
EXTRA="" run atom_m6 VKHR_B200_WALK=atom VKHR_B200_LIB=$PWD/tools/_libs/libvkhr_m6.so
EXTRA="" run atom_m8 VKHR_B200_WALK=atom VKHR_B200_LIB=$PWD/tools/_libs/libvkhr_m8.so
EXTRA="" run null_m8 VKHR_B200_WALK=null VKHR_B200_LIB=$PWD/tools/_libs/libvkhr_m8.so
EXTRA="" run red_m8 VKHR_B200_WALK=red VKHR_B200_NO_PIPELINE=1 VKHR_B200_LIB=$PWD/tools/_libs/libvkhr_m8.so
