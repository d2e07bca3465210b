#!/bin/bash
# round 2, GPU call AO: neighbour absorption in the walk of large power-of-two grids (EXACT = 4): parity at size, then configs[1] / [2] / [4] with and without
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > gpurun_out/pytest_gpu_ao.log; cat gpurun_out/pytest_gpu_ao.log
cp vkhr_b200/lib/libvkhr_b200.so /tmp/product.so
for ab in product noabsorb4; do
  [ $ab = product ] || cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
  timeout 300 python tools/big_probe.py > gpurun_out/big_probe_ao_$ab.json 2>/dev/null
  echo $ab; python - <<PY
import json
d=json.load(open('gpurun_out/big_probe_ao_$ab.json'))
for k,v in d.items(): print('  ',k, round(v['ms'],4), v['phases'])
PY
done
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sectors_srcunit_tex_op_red.sum --clock-control none -k regex:k_walk_uniform -c 1 --csv --log-file gpurun_out/launches_ao.csv python tools/big_probe.py > /dev/null 2>&1
grep k_walk gpurun_out/launches_ao.csv | awk -F'","' '{print $13, $15}'
