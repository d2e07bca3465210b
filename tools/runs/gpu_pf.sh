#!/bin/bash
# prefilter / ADSM: parity tests, device-resident timings, ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_prefilter_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_pf.log; tail -4 gpurun_out/pytest_pf.log
for w in 256 512 1024; do PF_NO_ADSM=$PF_NO_ADSM timeout 600 python tools/pf_time.py $w > gpurun_out/pf_$w.log 2>&1; tail -1 gpurun_out/pf_$w.log; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_prefilter_tiled -s 2 -c 1 -o gpurun_out/prof_pf512 -f python tools/pf_time.py 512 > gpurun_out/ncu_pf.log 2>&1
if [ -z "$PF_NO_ADSM" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k k_adsm -c 1 -o gpurun_out/prof_adsm256 -f python tools/pf_time.py 256 > gpurun_out/ncu_adsm.log 2>&1
fi
ls -la gpurun_out/*.ncu-rep
