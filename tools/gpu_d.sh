#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_api.py -v -m gpu -k "cyclosynchrotron_emission or outside_the_table or dropin_covers" > gpurun_out/pytest_gpu_d1.log 2>&1; grep -E "PASS|FAIL|Error|error|assert" gpurun_out/pytest_gpu_d1.log | cut -c1-220 | head -40; tail -3 gpurun_out/pytest_gpu_d1.log
