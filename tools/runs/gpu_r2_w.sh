#!/bin/bash
# round 2, GPU call W: the state of record -- GPU suite, smoke, bench (both arms), ncu launch list + full capture of the frame kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_w.log; cat gpurun_out/pytest_gpu_w.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_w.json').read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'], 'value', d['value'], 'frac %.3f'%d['roofline']['frac'], d['roofline']['phase_ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
for k,v in d['other_configs'].items(): print('   ', k[:90], {a:b for a,b in v.items() if a in ('ms','strategy','voxelise_ms_per_frame','voxelise_and_prefilter_ms_per_frame','error')})
print(d['strand_sharded'].get('one_gpu'), d['cpu_baseline'])
PY
timeout 600 python bench.py --impl reference > gpurun_out/bench_w_reference.json 2>/dev/null; tail -c 600 gpurun_out/bench_w_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_w.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-sharded > gpurun_out/ncu_w.log 2>&1
grep -E "k_frame|k_repair" gpurun_out/launches_w.csv | awk -F'","' '{print $5, $13, $15}' | cut -c1-160 | head -24
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 3 -c 1 -o gpurun_out/prof_frame64_w -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others --no-sharded > gpurun_out/ncu_full_w.log 2>&1
ls -la gpurun_out/prof_frame64_w.ncu-rep
