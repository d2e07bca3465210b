#!/bin/bash
# round 2, GPU call AD: interior walk with neighbour absorption (second sample of lane t joins the first sample of lane t+1 when the words agree)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu_ad.log; cat gpurun_out/pytest_gpu_ad.log
B="timeout 300 python bench.py --no-e2e --no-cpu --no-others --no-sharded --steps 20 --warmup 3"
cp vkhr_b200/lib/libvkhr_b200.so /tmp/product.so
for ab in product nodealt; do
  [ $ab = product ] || cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
  $B > gpurun_out/bench_ad_${ab}.json 2>/dev/null
  $B --strategy brick8-split > gpurun_out/bench_ad_${ab}_split.json 2>/dev/null
done
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_ad_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_ad_')[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], d['roofline'].get('phase_ms_per_step'))
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 3 -c 1 -o gpurun_out/prof_frame64_ad -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others --no-sharded > gpurun_out/ncu_full_ad.log 2>&1
ls -la gpurun_out/prof_frame64_ad.ncu-rep
