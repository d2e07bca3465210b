#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/pf_time.py 256 > gpurun_out/pf_256.log 2>&1; tail -1 gpurun_out/pf_256.log
timeout 600 python tools/pf_time.py 512 > gpurun_out/pf_512.log 2>&1; tail -1 gpurun_out/pf_512.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_prefilter_tiled -s 2 -c 1 -o gpurun_out/prof_pf512 -f python tools/pf_time.py 512 > gpurun_out/ncu_pf.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_adsm -c 1 -o gpurun_out/prof_adsm256 -f python tools/pf_time.py 256 > gpurun_out/ncu_adsm.log 2>&1
ls -la gpurun_out/*.ncu-rep
