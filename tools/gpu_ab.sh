#!/bin/bash
# column-form AO: parity tests on the product build, then A/B timings of the tuning builds (tools only)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_prefilter_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_pf.log; tail -4 gpurun_out/pytest_pf.log
export PF_NO_ADSM=1
for lib in libvkhr_b200 ab_minctas3 ab_tz16; do
  for w in 256 512 1024; do
    VKHR_B200_LIB=$PWD/vkhr_b200/lib/$lib.so timeout 300 python tools/pf_time.py $w > gpurun_out/pf_${lib}_$w.log 2>&1; tail -1 gpurun_out/pf_${lib}_$w.log
  done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_prefilter_tiled -s 2 -c 1 -o gpurun_out/prof_pfcol512 -f python tools/pf_time.py 512 > gpurun_out/ncu_pf.log 2>&1
ls -la gpurun_out/*.ncu-rep
