#!/bin/bash
# BRICK8 strategy: parity of the voxelisation tests, the other BASELINE configs under each strategy
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
for st in packed8 brick8 auto; do
  timeout 300 python bench.py --strategy $st --no-e2e --no-cpu --steps 10 --warmup 3 > gpurun_out/bench_$st.json 2> gpurun_out/bench_$st.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$st.json").read().strip().splitlines()[-1])
print("$st", d["ms_per_step"], d["roofline"]["phase_ms_per_step"], d["value"])
for k,v in d.get("other_configs",{}).items(): print("   ", k, v.get("ms"), v.get("strategy"), v.get("error"))
PY
done
