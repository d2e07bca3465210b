#!/bin/bash
# round 2, GPU call AM: compute-sanitizer synccheck (barrier use of the frame kernel's roles) and initcheck on the frame-kernel and batch cases
mkdir -p gpurun_out
K="frame or batch or brick8"
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 9 --log-file gpurun_out/sanitize_sync.log \
    python -m pytest tests/test_parity_gpu.py -m gpu -q -k "($K) and not full_size and not crowd_at_256 and not fingerprints" > gpurun_out/sanitize_sync_pytest.log 2>&1
echo "synccheck rc=$?"; tail -1 gpurun_out/sanitize_sync_pytest.log; grep -E "ERROR SUMMARY" gpurun_out/sanitize_sync.log | tail -1
timeout 300 compute-sanitizer --tool initcheck --error-exitcode 9 --log-file gpurun_out/sanitize_init.log \
    python -m pytest tests/test_parity_gpu.py -m gpu -q -k "frame and not full_size and not crowd_at_256" > gpurun_out/sanitize_init_pytest.log 2>&1
echo "initcheck rc=$?"; tail -1 gpurun_out/sanitize_init_pytest.log; grep -E "ERROR SUMMARY" gpurun_out/sanitize_init.log | tail -1; grep -E "Uninitialized" gpurun_out/sanitize_init.log | sort | uniq -c | head -5
