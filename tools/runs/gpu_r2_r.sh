#!/bin/bash
# round 2, GPU call R: frame kernel at 5 CTAs per SM (48 registers) with 64 / 96 / 128 copier CTAs per instance
mkdir -p gpurun_out
B="timeout 300 python bench.py --no-e2e --no-cpu --no-others --no-sharded --steps 20 --warmup 3"
$B > gpurun_out/bench_r_base.json 2>/dev/null
cp vkhr_b200/lib/libvkhr_b200.so /tmp/product.so
for ab in c5_64 c5_96 c5_128; do
  cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
  $B > gpurun_out/bench_r_$ab.json 2>/dev/null
  $B --ring-mib 80 > gpurun_out/bench_r_${ab}_ring80.json 2>/dev/null
  $B --ring-mib 96 > gpurun_out/bench_r_${ab}_ring96.json 2>/dev/null
done
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_r_')[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
