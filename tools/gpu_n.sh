#!/bin/bash
# round 2, GPU session N: DRAM traffic of K1 at the bench's size, --set full of the large-list pass / event kernels
mkdir -p gpurun_out
timeout 300 python tools/prof_loop.py 1e7 128 24 0 > gpurun_out/prof_loop_plain.log 2>&1; tail -2 gpurun_out/prof_loop_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pass_local_kernel|event_local_kernel" --launch-skip 30 --launch-count 2 \
  -f -o gpurun_out/prof_r02_loop1e7 python tools/prof_loop.py 1e7 128 24 0 > gpurun_out/ncu_full_loop.log 2>&1; echo "ncu loop rc=$?"; tail -3 gpurun_out/ncu_full_loop.log
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active --clock-control none \
  -k regex:"scan_kernel" -c 2 --csv --log-file gpurun_out/scan1e7_dram.csv python tools/prof_loop.py 1e7 128 2 1 > gpurun_out/ncu_scan1e7.log 2>&1; echo "ncu scan rc=$?"; cat gpurun_out/scan1e7_dram.csv | tail -10
ls -la gpurun_out/*.ncu-rep
