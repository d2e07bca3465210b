#!/bin/bash
# round 2, GPU call AK: compute-sanitizer over the final kernels (memcheck on the parity suite minus the full-size cases; racecheck on the frame-kernel cases)
mkdir -p gpurun_out
SAN_TIMEOUT=400 bash tools/gpu_sanitize.sh
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/sanitize_race.log \
    python -m pytest tests/test_parity_gpu.py -m gpu -q -k "frame and not full_size and not crowd_at_256" > gpurun_out/sanitize_race_pytest.log 2>&1
echo "racecheck rc=$?"; tail -2 gpurun_out/sanitize_race_pytest.log; grep -E "RACECHECK SUMMARY|ERROR SUMMARY" gpurun_out/sanitize_race.log | tail -2
