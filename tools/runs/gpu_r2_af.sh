#!/bin/bash
# round 2, GPU call AF: range geometry of the frame kernel: 16 tiles x 1 range, 12 x 1 (4 and 5 CTAs per SM), 8 x 3 against 8 x 2 (product)
mkdir -p gpurun_out
B="timeout 300 python bench.py --no-e2e --no-cpu --no-sharded --steps 20 --warmup 3"
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "brick or frame or batch or crowd or golden or edge or random" 2>&1 | tail -3
cp vkhr_b200/lib/libvkhr_b200.so /tmp/product.so
for ab in product t16r1 t12r1 t12r1_cta5 t8r3; do
  [ $ab = product ] || cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
  $B > gpurun_out/bench_af_${ab}.json 2>/dev/null
done
cp vkhr_b200/lib/ab_t16r1.so vkhr_b200/lib/libvkhr_b200.so
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "brick or frame or batch or crowd or golden or edge or random" 2>&1 | tail -3
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_af_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_af_')[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])
        for k,v in d['other_configs'].items():
            if 'ms' in v: print('   ', k[:70], round(v['ms'],4))
    except Exception as e:
        print(f, 'ERR', e)
PY
