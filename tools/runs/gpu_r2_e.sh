#!/bin/bash
# round 2, GPU call E: frame kernel v3 (ticketed CTAs, the last finishers of an instance copy it out), sharded C entry
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_e.log; tail -4 gpurun_out/pytest_gpu_e.log
B="timeout 300 python bench.py --no-e2e --no-cpu --steps 20 --warmup 3"
$B > gpurun_out/bench_e_frame.json 2> gpurun_out/bench_e_frame.err
$B --strategy brick8-split --no-others --no-sharded > gpurun_out/bench_e_split.json 2> gpurun_out/bench_e_split.err
for r in 32 64 128; do $B --no-others --no-sharded --ring-mib $r > gpurun_out/bench_e_ring$r.json 2> gpurun_out/bench_e_ring$r.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_e_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('bench_e_')[1], 'ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], d['roofline']['phase_ms_per_step'])
        if d.get('other_configs'):
            for k,v in d['other_configs'].items(): print('   ', k[:60], v.get('ms'), v.get('strategy'))
        if d.get('strand_sharded'): print('   sharded', d['strand_sharded'].get('one_gpu'))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 6 -c 4 --csv --log-file gpurun_out/launches_e.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-others --no-sharded > gpurun_out/ncu_e.log 2>&1
grep -E "k_frame" gpurun_out/launches_e.csv | awk -F'","' '{print $5, $13, $15}' | tail -5
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 3 -c 1 -o gpurun_out/prof_frame64_v3 -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others --no-sharded > gpurun_out/ncu_full_e.log 2>&1
ls -la gpurun_out/prof_frame64_v3.ncu-rep
