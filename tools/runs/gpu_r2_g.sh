#!/bin/bash
# round 2, GPU call G: frame kernel v4 (block-index order, autonomous warps, static copiers); sharded tests; sparse prefilter; tangents
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu_g.log; tail -6 gpurun_out/pytest_gpu_g.log
timeout 120 compute-sanitizer --tool memcheck ./adapter/_build/sharded_test > gpurun_out/sharded_test_memcheck.log 2>&1; tail -25 gpurun_out/sharded_test_memcheck.log
B="timeout 400 python bench.py --no-e2e --no-cpu --steps 20 --warmup 3"
$B > gpurun_out/bench_g_frame.json 2> gpurun_out/bench_g_frame.err
for r in 48 96 128; do $B --no-others --no-sharded --ring-mib $r > gpurun_out/bench_g_ring$r.json 2> gpurun_out/bench_g_ring$r.err; done
cp vkhr_b200/lib/libvkhr_b200.so /tmp/product.so
for ab in ranges1 ctas5 ranges1_ctas5; do
  cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
  $B --no-others --no-sharded > gpurun_out/bench_g_$ab.json 2> gpurun_out/bench_g_$ab.err
  $B --no-others --no-sharded --ring-mib 128 > gpurun_out/bench_g_${ab}_ring128.json 2> gpurun_out/bench_g_${ab}_ring128.err
done
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_g_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_g_')[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], d['roofline']['phase_ms_per_step'])
        if d.get('other_configs'):
            for k,v in d['other_configs'].items(): print('   ', k[:70], {a:b for a,b in v.items() if a in ('ms','strategy','voxelise_ms_per_frame','voxelise_and_prefilter_ms_per_frame','error')})
        if d.get('strand_sharded'): print('   sharded', d['strand_sharded'].get('one_gpu'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 3 -c 1 -o gpurun_out/prof_frame64_v4 -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others --no-sharded > gpurun_out/ncu_full_g.log 2>&1
ls -la gpurun_out/prof_frame64_v4.ncu-rep
