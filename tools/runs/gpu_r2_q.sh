#!/bin/bash
# round 2, GPU call Q: frame kernel with dynamic range tickets inside a CTA (A/B)
mkdir -p gpurun_out
B="timeout 300 python bench.py --no-e2e --no-cpu --no-others --no-sharded --steps 20 --warmup 3"
$B > gpurun_out/bench_q_base.json 2>/dev/null
cp vkhr_b200/lib/libvkhr_b200.so /tmp/product.so
for ab in dyn2 dyn3 dyn4; do
  cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
  timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "frame_kernel or golden_small or baseline_configs" 2>&1 | tail -2
  $B > gpurun_out/bench_q_$ab.json 2>/dev/null
  $B --ring-mib 96 > gpurun_out/bench_q_${ab}_ring96.json 2>/dev/null
  $B --instances 8 --steps 50 > gpurun_out/bench_q_${ab}_8inst.json 2>/dev/null
done
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_q_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_q_')[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
