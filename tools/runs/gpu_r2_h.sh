#!/bin/bash
# round 2, GPU call H: frame kernel v5 (block order, CTA-level report, static copier CTAs) + split with 4/6 CTAs per SM
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu_h.log; tail -4 gpurun_out/pytest_gpu_h.log
B="timeout 400 python bench.py --no-e2e --no-cpu --no-others --no-sharded --steps 20 --warmup 3"
$B > gpurun_out/bench_h_frame.json 2> gpurun_out/bench_h_frame.err
$B --strategy brick8-split > gpurun_out/bench_h_split.json 2> gpurun_out/bench_h_split.err
for r in 48 96 128; do $B --ring-mib $r > gpurun_out/bench_h_ring$r.json 2> gpurun_out/bench_h_ring$r.err; done
cp vkhr_b200/lib/libvkhr_b200.so /tmp/product.so
for ab in ranges1 ctas5; do
  cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
  $B > gpurun_out/bench_h_$ab.json 2> gpurun_out/bench_h_$ab.err
  $B --ring-mib 128 > gpurun_out/bench_h_${ab}_ring128.json 2> gpurun_out/bench_h_${ab}_ring128.err
done
for ab in walk4 walk6; do
  cp vkhr_b200/lib/ab_$ab.so vkhr_b200/lib/libvkhr_b200.so
  $B --strategy brick8-split > gpurun_out/bench_h_split_$ab.json 2> gpurun_out/bench_h_split_$ab.err
done
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_h_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_h_')[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], d['roofline']['phase_ms_per_step'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 3 -c 1 -o gpurun_out/prof_frame64_v5 -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others --no-sharded > gpurun_out/ncu_full_h.log 2>&1
ls -la gpurun_out/prof_frame64_v5.ncu-rep
