#!/bin/bash
# round 2, GPU call T: more L2 cache hints (copier ring accesses evict_last = 4, output stores evict_first = 8)
mkdir -p gpurun_out
B="timeout 300 python bench.py --no-e2e --no-cpu --no-others --no-sharded --steps 20 --warmup 3"
cp vkhr_b200/lib/libvkhr_b200.so /tmp/product.so
for ab in l2h1 l2h5 l2h9 l2h13 l2h15 l2h7; do
  cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
  timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "frame_kernel or golden_small" 2>&1 | tail -1
  $B > gpurun_out/bench_t_$ab.json 2>/dev/null
  $B --ring-mib 80 > gpurun_out/bench_t_${ab}_ring80.json 2>/dev/null
  $B --instances 8 --steps 50 > gpurun_out/bench_t_${ab}_8inst.json 2>/dev/null
  timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_frame -s 3 -c 1 --csv --log-file gpurun_out/traffic_t_$ab.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others --no-sharded > /dev/null 2>&1
  grep k_frame gpurun_out/traffic_t_$ab.csv | awk -F'","' '{print $13, $15}' | tr '\n' ' '; echo
done
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_t_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_t_')[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
