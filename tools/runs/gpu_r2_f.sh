#!/bin/bash
# round 2, GPU call F: frame kernel with 2 ranges per warp-item (A/B: 1, 4 ranges; 4 CTAs per SM), ring sizes; sparse prefilter
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_f.log; tail -4 gpurun_out/pytest_gpu_f.log
B="timeout 300 python bench.py --no-e2e --no-cpu --steps 20 --warmup 3"
$B > gpurun_out/bench_f_frame.json 2> gpurun_out/bench_f_frame.err
for r in 48 80 128; do $B --no-others --no-sharded --ring-mib $r > gpurun_out/bench_f_ring$r.json 2> gpurun_out/bench_f_ring$r.err; done
cp vkhr_b200/lib/libvkhr_b200.so /tmp/product.so
for ab in ranges1 ranges4 minctas4; do
  cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
  $B --no-others --no-sharded > gpurun_out/bench_f_$ab.json 2> gpurun_out/bench_f_$ab.err
  $B --no-others --no-sharded --ring-mib 128 > gpurun_out/bench_f_${ab}_ring128.json 2> gpurun_out/bench_f_${ab}_ring128.err
done
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_f_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_f_')[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], d['roofline']['phase_ms_per_step'])
        if d.get('other_configs'):
            for k,v in d['other_configs'].items(): print('   ', k[:60], v.get('ms'), v.get('strategy'))
        if d.get('strand_sharded'): print('   sharded', d['strand_sharded'].get('one_gpu'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
