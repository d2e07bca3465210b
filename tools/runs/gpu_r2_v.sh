#!/bin/bash
# round 2, GPU call V: copier reads-and-zeroes the ring with atom.exch.b128 (A/B)
mkdir -p gpurun_out
B="timeout 300 python bench.py --no-e2e --no-cpu --no-others --no-sharded --steps 20 --warmup 3"
cp vkhr_b200/lib/libvkhr_b200.so /tmp/product.so
for ab in product exch; do
  [ $ab = product ] || cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
  timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "frame_kernel or golden_small" 2>&1 | tail -1
  for ring in 48 64 96; do
    $B --ring-mib $ring > gpurun_out/bench_v_${ab}_ring$ring.json 2>/dev/null
    timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_frame -s 3 -c 1 --csv --log-file gpurun_out/traffic_v_${ab}_ring$ring.csv \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others --no-sharded --ring-mib $ring > /dev/null 2>&1
    echo $ab ring $ring: $(grep k_frame gpurun_out/traffic_v_${ab}_ring$ring.csv | awk -F'","' '{print $13, $15}' | tr '\n' ' ')
  done
done
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_v_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_v_')[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
