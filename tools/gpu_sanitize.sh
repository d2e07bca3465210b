#!/bin/bash
# compute-sanitizer memcheck over the GPU suite minus the full-size cases (out-of-bounds brick / tile addressing, misaligned accesses)
mkdir -p gpurun_out
timeout ${SAN_TIMEOUT:-240} compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitize_parity.log \
    python -m pytest tests/test_parity_gpu.py -m gpu -q -k "not fast_division and not full_size and not fingerprints and not crowd_at_256 and not 2pow24 and not cpp_dropin" > gpurun_out/sanitize_parity_pytest.log 2>&1
echo "parity rc=$?"; tail -2 gpurun_out/sanitize_parity_pytest.log; grep "ERROR SUMMARY" gpurun_out/sanitize_parity.log | tail -2
timeout ${SAN_TIMEOUT:-240} compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitize_pf.log \
    python -m pytest tests/test_prefilter_gpu.py -m gpu -q -k "not full_size" > gpurun_out/sanitize_pf_pytest.log 2>&1
echo "prefilter rc=$?"; tail -2 gpurun_out/sanitize_pf_pytest.log; grep "ERROR SUMMARY" gpurun_out/sanitize_pf.log | tail -2
