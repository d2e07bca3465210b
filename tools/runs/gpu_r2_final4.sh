#!/bin/bash
# round 2, check after the host-side clamp of the last instance's copiers: smoke + frame / batch cases
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "frame or batch or golden" 2>&1 | tail -1
