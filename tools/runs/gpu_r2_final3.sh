#!/bin/bash
# round 2, last check of the shipped library (rebuilt from the validated sources): smoke + the parity cases at size + a short bench
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "full_size or frame or golden or edge" 2>&1 | tail -2
timeout 300 python bench.py --no-e2e --no-cpu --no-others --no-sharded --steps 20 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.4f frac %.3f'%(d['ms_per_step'],d['roofline']['frac']))"
