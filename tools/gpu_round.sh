#!/bin/bash
# One GPU session of a round: tests, both bench arms, ncu launch list of bench.py, ncu --set full of the top kernels.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 600 gpurun_out/bench_ours.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-pass-roofline > gpurun_out/ncu_launch.log 2>&1
timeout 120 python tools/prof_driver.py 200 10000000 16 > gpurun_out/prof_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"scan_kernel|frame_loop_kernel|event_kernel" -c 4 \
    -f -o gpurun_out/prof_full python tools/prof_driver.py 200 0 16 > gpurun_out/ncu_full.log 2>&1
# steady-state pass over a list larger than L2: the 6th pass launch (the 2nd re-checks every photon after the upload)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pass_kernel" --launch-skip 5 --launch-count 1 \
    -f -o gpurun_out/prof_full_pass python tools/prof_driver.py 2 10000000 16 > gpurun_out/ncu_full_pass.log 2>&1
timeout 120 python tools/prof_driver.py 50 0 16 auto C5 > gpurun_out/prof_plain_c5.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan_kernel" -c 1 \
    -f -o gpurun_out/prof_full_c5 python tools/prof_driver.py 50 0 16 auto C5 > gpurun_out/ncu_full_c5.log 2>&1
timeout 200 python tools/loop_timing.py C2 100000 16 2000 > gpurun_out/loop_timing.log 2>&1
ls -la gpurun_out | tail -12
