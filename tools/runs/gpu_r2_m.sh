#!/bin/bash
# round 2, GPU call M: double-buffered staging + more copiers for the last instance: suite, bench, ncu of the frame kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_m.log; cat gpurun_out/pytest_gpu_m.log
timeout 900 python bench.py > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_m.json').read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'], d['roofline']['phase_ms_per_step'], 'e2e', d['e2e']['value'])
for k,v in d['other_configs'].items(): print('   ', k[:70], {a:b for a,b in v.items() if a in ('ms','strategy','voxelise_ms_per_frame','voxelise_and_prefilter_ms_per_frame','error')})
print(d['strand_sharded'].get('one_gpu'), d['cpu_baseline'])
PY
timeout 200 python bench.py --no-e2e --no-cpu --no-others --no-sharded --instances 8 --steps 50 > gpurun_out/bench_m_8inst.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_m_8inst.json').read().strip().splitlines()[-1]); print('8 instances: ms/step %.4f'%d['ms_per_step'], d['roofline']['phase_ms_per_step'])"
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 6 -c 4 --csv --log-file gpurun_out/launches_m.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-others --no-sharded > gpurun_out/ncu_m.log 2>&1
grep -E "k_frame|k_repair" gpurun_out/launches_m.csv | awk -F'","' '{print $5, $13, $15}' | tail -10
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 3 -c 1 -o gpurun_out/prof_frame64_m -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others --no-sharded > gpurun_out/ncu_full_m.log 2>&1
ls -la gpurun_out/prof_frame64_m.ncu-rep
