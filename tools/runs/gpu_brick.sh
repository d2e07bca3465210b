#!/bin/bash
# BRICK8 strategy: parity of the voxelisation tests (+ smoke), then the crowd frame under the product build and under A/B builds
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_sharding.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for lib in libvkhr_b200 $AB_LIBS; do
  [ -f vkhr_b200/lib/$lib.so ] || continue
  VKHR_B200_LIB=$PWD/vkhr_b200/lib/$lib.so timeout 300 python bench.py --no-e2e --no-cpu --steps 20 --warmup 3 > gpurun_out/bench_$lib.json 2> gpurun_out/bench_$lib.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$lib.json").read().strip().splitlines()[-1])
print("$lib", d["ms_per_step"], d["roofline"]["phase_ms_per_step"], d["value"])
for k,v in d.get("other_configs",{}).items(): print("   ", k, v.get("ms"), v.get("strategy"), v.get("error"))
PY
done
