#!/bin/bash
# round 2, GPU session K3: timings of the persistent stream with matched shared-memory configurations
mkdir -p gpurun_out
export MCRAT_B200_DEBUG=1
L=mcrat_b200/csrc/libmcrat_b200.so
run() { echo "== $*"; env "$@" timeout 120 python tools/ab_compare.py $L:persistent_stream $L:persistent_stream C5 10000000 128 200 2>&1 | grep -E "persistent stream|us/iteration" | sed -n '1p;3p'; }
( run A=1
run MCRAT_B200_STREAM_CARVEOUT=28 MCRAT_B200_STREAM_PAD_PASS=5120 MCRAT_B200_STREAM_PAD_EVT=3072
run MCRAT_B200_STREAM_EVT_BLOCKS=16
run MCRAT_B200_STREAM_EVT_BLOCKS=64
run MCRAT_B200_STREAM_EVT_BLOCKS=128
run MCRAT_B200_STREAM_PPT=8
run MCRAT_B200_STREAM_PPT=32
run MCRAT_B200_STREAM_PASS_BLOCKS=592
run MCRAT_B200_STREAM_PASS_BLOCKS=740 ) 2>&1 | tee gpurun_out/dbg_k3.log
unset MCRAT_B200_DEBUG
timeout 300 python tools/ab_compare.py $L:streamed $L:persistent_stream C5 10000000 16 300 2>&1 | tail -3
timeout 300 python tools/ab_compare.py $L:streamed $L:persistent_stream C5 10000000 296 300 2>&1 | tail -3
timeout 300 python tools/ab_compare.py $L:streamed $L:persistent_stream C5 5000000 64 300 2>&1 | tail -3
