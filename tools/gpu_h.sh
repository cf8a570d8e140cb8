#!/bin/bash
# round 2, GPU session H: co-residency of the two grids of the persistent stream
mkdir -p gpurun_out
L=mcrat_b200/csrc/libmcrat_b200.so
( timeout 300 python tools/ab_compare.py $L:streamed $L:persistent_stream C5 10000000 128 300 2>&1 | tail -3 | sed "s/^/carve+pad /"
MCRAT_B200_STREAM_NOCARVE=1 MCRAT_B200_STREAM_PAD=0 timeout 300 python tools/ab_compare.py $L:persistent_stream $L:persistent_stream C5 10000000 128 300 2>&1 | tail -2 | sed "s/^/nocarve,nopad /"
MCRAT_B200_STREAM_PAD=0 timeout 300 python tools/ab_compare.py $L:persistent_stream $L:persistent_stream C5 10000000 128 300 2>&1 | tail -2 | sed "s/^/carve,nopad /"
for ppt in 8 32; do
MCRAT_B200_STREAM_PPT=$ppt timeout 300 python tools/ab_compare.py $L:persistent_stream $L:persistent_stream C5 10000000 128 300 2>&1 | tail -2 | sed "s/^/ppt=$ppt /"
done
timeout 300 python tools/ab_compare.py $L:streamed $L:persistent_stream C5 10000000 16 300 2>&1 | tail -3
timeout 300 python tools/ab_compare.py $L:streamed $L:persistent_stream C5 5000000 64 300 2>&1 | tail -3 ) | tee gpurun_out/ab_h.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "persistent or klein or walks or called_off" 2>&1 | tail -4
