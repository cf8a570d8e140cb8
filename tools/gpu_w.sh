#!/bin/bash
# round 2, GPU session W: compute-sanitizer over every loop driver, the communicator and the sort (small frames)
mkdir -p gpurun_out
export MCRAT_B200_DEBUG=1
timeout 200 python tools/sanitize_driver.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -12 gpurun_out/sanitize_plain.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_driver.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|persistent stream|Invalid|error" gpurun_out/sanitize_memcheck.log | head -20
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_driver.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|persistent stream|hazard" gpurun_out/sanitize_racecheck.log | head -20
