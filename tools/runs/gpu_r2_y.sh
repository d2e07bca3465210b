#!/bin/bash
# round 2, GPU call Y: frame kernel after the instruction diet (reconvergence before the transform, incremental strand-end
# test, grid constants pinned once per CTA, constant-size bulk copy, multiply-high brick coordinates + dp4a in the copy-out)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu_y.log; cat gpurun_out/pytest_gpu_y.log
B="timeout 300 python bench.py --no-e2e --no-cpu --no-sharded --steps 20 --warmup 3"
$B > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_y.json').read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'], 'value', d['value'], 'frac %.3f'%d['roofline']['frac'], d['roofline']['phase_ms_per_step'])
for k,v in d['other_configs'].items(): print('   ', k[:90], {a:b for a,b in v.items() if a in ('ms','strategy','voxelise_ms_per_frame','voxelise_and_prefilter_ms_per_frame','error')})
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_frame -s 3 -c 1 -o gpurun_out/prof_frame64_y -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others --no-sharded > gpurun_out/ncu_full_y.log 2>&1
ls -la gpurun_out/prof_frame64_y.ncu-rep
