#!/bin/bash
mkdir -p gpurun_out
VKHR_B200_WALK=red timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk_pipeline -s 3 -c 1 -o gpurun_out/prof_pipe64 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/ncu_pipe.log 2>&1
VKHR_B200_WALK=red VKHR_B200_ALIAS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk_pipeline -s 3 -c 1 -o gpurun_out/prof_pipe64_alias2 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/ncu_pipe_alias.log 2>&1
ls -la gpurun_out/*.ncu-rep
