#!/bin/bash
# round 2, GPU call B: the frame kernel -- parity, timing against the split form, ring sizes, 4 vs 5 CTAs per SM, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not at_full_size and not sway_sequence" 2>&1 | tail -15 > gpurun_out/pytest_gpu_b.log; tail -3 gpurun_out/pytest_gpu_b.log
B="timeout 300 python bench.py --no-e2e --no-cpu --no-others --steps 20 --warmup 3"
$B > gpurun_out/bench_b_frame.json 2> gpurun_out/bench_b_frame.err
$B --strategy brick8-split > gpurun_out/bench_b_split.json 2> gpurun_out/bench_b_split.err
for r in 16 32 64 96 128; do $B --ring-mib $r > gpurun_out/bench_b_ring$r.json 2> gpurun_out/bench_b_ring$r.err; done
cp vkhr_b200/lib/libvkhr_b200.so /tmp/product.so
cp vkhr_b200/lib/ab_minctas4.so vkhr_b200/lib/libvkhr_b200.so
$B > gpurun_out/bench_b_frame_cta4.json 2> gpurun_out/bench_b_frame_cta4.err
$B --strategy brick8-split > gpurun_out/bench_b_split_cta4.json 2> gpurun_out/bench_b_split_cta4.err
cp /tmp/product.so vkhr_b200/lib/libvkhr_b200.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_b_')[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], d['roofline']['phase_ms_per_step'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 6 -c 4 --csv --log-file gpurun_out/launches_b.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/ncu_b.log 2>&1
grep -E "k_frame|k_repair" gpurun_out/launches_b.csv | awk -F'","' '{print $5, $13, $15}' | tail -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 3 -c 1 -o gpurun_out/prof_frame64 -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/ncu_full_b.log 2>&1
ls -la gpurun_out/prof_frame64.ncu-rep
