#!/bin/bash
# round 2, re-validation after the last clean-up of the frame plan / comments: GPU suite, smoke, short bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_final2.log; cat gpurun_out/pytest_gpu_final2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py --no-e2e --no-cpu --no-others --no-sharded --steps 20 --warmup 3 > gpurun_out/bench_final2.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final2.json').read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])
PY
