#!/bin/bash
# round 2, GPU call AB: frame kernel at 5 and 6 CTAs per SM (48 / 40 registers: the spills are in the copy-out only now), ring 64 / 80 MiB
mkdir -p gpurun_out
B="timeout 300 python bench.py --no-e2e --no-cpu --no-others --no-sharded --steps 20 --warmup 3"
cp vkhr_b200/lib/libvkhr_b200.so /tmp/product.so
for ab in product cta5 cta5_cop320 cta6 cta6_if2 cta6_if2_cop384; do
  [ $ab = product ] || cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
  for ring in 64 80; do
    $B --ring-mib $ring > gpurun_out/bench_ab_${ab}_ring$ring.json 2>/dev/null
  done
done
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_ab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_ab_')[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
