#!/bin/bash
# round 2, GPU call S: L2 cache hints (vertices evict_first, ring reds evict_last) A/B, with DRAM traffic from ncu
mkdir -p gpurun_out
B="timeout 300 python bench.py --no-e2e --no-cpu --no-others --no-sharded --steps 20 --warmup 3"
$B > gpurun_out/bench_s_base.json 2>/dev/null
cp vkhr_b200/lib/libvkhr_b200.so /tmp/product.so
for ab in l2h1 l2h2 l2h3; do
  cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
  timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "frame_kernel or golden_small" 2>&1 | tail -1
  $B > gpurun_out/bench_s_$ab.json 2>/dev/null
  $B --ring-mib 96 > gpurun_out/bench_s_${ab}_ring96.json 2>/dev/null
  $B --strategy brick8-split > gpurun_out/bench_s_${ab}_split.json 2>/dev/null
  timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_frame -s 3 -c 1 --csv --log-file gpurun_out/traffic_s_$ab.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others --no-sharded > /dev/null 2>&1
  grep k_frame gpurun_out/traffic_s_$ab.csv | awk -F'","' '{print $5, $13, $15}' | cut -c1-150
done
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_s_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_s_')[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
