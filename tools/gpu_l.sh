#!/bin/bash
# round 2, GPU session L: whole GPU suite on the current build, persistent-stream defaults
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu --maxfail=12 ) > gpurun_out/pytest_gpu_l.log 2>&1; tail -15 gpurun_out/pytest_gpu_l.log
export MCRAT_B200_DEBUG=1
L=mcrat_b200/csrc/libmcrat_b200.so
run() { echo "== $*"; env "$@" timeout 120 python tools/ab_compare.py $L:persistent_stream $L:persistent_stream C5 10000000 128 200 2>&1 | grep -E "persistent stream|us/iteration" | sed -n '1p;3p'; }
( run A=1
run MCRAT_B200_STREAM_PPT=48
run MCRAT_B200_STREAM_PPT=64
run MCRAT_B200_STREAM_EVT_BLOCKS=43
run MCRAT_B200_STREAM_PASS_BLOCKS=720 ) 2>&1 | tee gpurun_out/dbg_l.log
unset MCRAT_B200_DEBUG
timeout 300 python tools/ab_compare.py $L:streamed $L:auto C5 5000000 64 300 2>&1 | tail -3
timeout 300 python tools/ab_compare.py $L:streamed $L:auto C5 2500000 32 300 2>&1 | tail -3
