#!/bin/bash
# round 2, GPU session I: persistent stream with shared event blocks -- tests, A/B, knobs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "persistent or klein or walks or called_off" 2>&1 | tail -4
L=mcrat_b200/csrc/libmcrat_b200.so
( timeout 300 python tools/ab_compare.py $L:streamed $L:persistent_stream C5 10000000 128 300 2>&1 | tail -3 | sed "s/^/default /"
for E in 16 64 128; do
MCRAT_B200_STREAM_EVT_BLOCKS=$E timeout 300 python tools/ab_compare.py $L:persistent_stream $L:persistent_stream C5 10000000 128 300 2>&1 | tail -2 | head -1 | sed "s/^/E=$E /"
done
for G in 592 650 740; do
MCRAT_B200_STREAM_PASS_BLOCKS=$G timeout 300 python tools/ab_compare.py $L:persistent_stream $L:persistent_stream C5 10000000 128 300 2>&1 | tail -2 | head -1 | sed "s/^/passblocks=$G /"
done
for ppt in 8 32; do
MCRAT_B200_STREAM_PPT=$ppt timeout 300 python tools/ab_compare.py $L:persistent_stream $L:persistent_stream C5 10000000 128 300 2>&1 | tail -2 | head -1 | sed "s/^/ppt=$ppt /"
done
timeout 300 python tools/ab_compare.py $L:streamed $L:persistent_stream C5 10000000 16 300 2>&1 | tail -3
timeout 300 python tools/ab_compare.py $L:streamed $L:persistent_stream C5 10000000 296 300 2>&1 | tail -3
timeout 300 python tools/ab_compare.py $L:streamed $L:persistent_stream C5 5000000 64 300 2>&1 | tail -3
timeout 300 python tools/ab_compare.py $L:streamed $L:persistent_stream C2 10000000 128 300 2>&1 | tail -3 ) | tee gpurun_out/ab_i.log
