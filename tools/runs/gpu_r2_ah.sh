#!/bin/bash
# round 2, GPU call AH: copiers look at the previous instance's counter as they come off their walk (no second barrier, no poll round trip)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu_ah.log; cat gpurun_out/pytest_gpu_ah.log
B="timeout 300 python bench.py --no-e2e --no-cpu --no-others --no-sharded --steps 20 --warmup 3"
cp vkhr_b200/lib/libvkhr_b200.so /tmp/product.so
for ab in product cop192 copall r2; do
  [ $ab = product ] || cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
  $B > gpurun_out/bench_ah_${ab}.json 2>/dev/null
done
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_ah_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_ah_')[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 3 -c 1 -o gpurun_out/prof_frame64_ah -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others --no-sharded > gpurun_out/ncu_full_ah.log 2>&1
ls -la gpurun_out/prof_frame64_ah.ncu-rep
