#!/bin/bash
# round 2, GPU call AG: where the frame's time is -- probe builds without the reds / without the copy-out / without both; 3 (product) and 4 ranges per item
mkdir -p gpurun_out
B="timeout 300 python bench.py --no-e2e --no-cpu --no-others --no-sharded --steps 20 --warmup 3"
cp vkhr_b200/lib/libvkhr_b200.so /tmp/product.so
for ab in product nored nocopy neither r4 r3_cop192 r3_cop4096; do
  [ $ab = product ] || cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
  $B > gpurun_out/bench_ag_${ab}.json 2>/dev/null
done
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_ag_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_ag_')[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
for ab in nored neither; do
cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_frame -s 3 -c 1 --csv --log-file gpurun_out/probe_ag_$ab.csv \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others --no-sharded > /dev/null 2>&1
echo $ab $(grep k_frame gpurun_out/probe_ag_$ab.csv | awk -F'","' '{print $13, $15}' | tr '\n' ' ')
done
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
