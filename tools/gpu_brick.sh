#!/bin/bash
# BRICK8 strategy: parity of the voxelisation tests, timing of the crowd frame, ncu of its walk
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
for st in auto; do
  timeout 300 python bench.py --strategy $st --no-e2e --no-cpu --no-others --steps 20 --warmup 3 > gpurun_out/bench_$st.json 2> gpurun_out/bench_$st.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$st.json").read().strip().splitlines()[-1])
print("$st", d["ms_per_step"], d["roofline"]["phase_ms_per_step"], d["value"])
PY
done
if [ -n "$NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_walk_uniform -s 3 -c 1 -o gpurun_out/prof_walk64_brick -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-others > gpurun_out/ncu_full.log 2>&1
fi
