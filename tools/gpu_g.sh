#!/bin/bash
# round 2, GPU session G: persistent-stream loop -- parity tests, A/B against the interleaved streamed loop at 1e7 photons
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "persistent or klein or walks" ) > gpurun_out/pytest_gpu_g.log 2>&1; tail -8 gpurun_out/pytest_gpu_g.log
L=mcrat_b200/csrc/libmcrat_b200.so
for S in 128 64 16; do
timeout 300 python tools/ab_compare.py $L:streamed $L:persistent_stream C5 10000000 $S 300 2>&1 | tail -3
done | tee gpurun_out/ab_g.log
for ppt in 8 32; do
MCRAT_B200_STREAM_PPT=$ppt timeout 300 python tools/ab_compare.py $L:persistent_stream $L:persistent_stream C5 10000000 128 300 2>&1 | tail -2 | sed "s/^/ppt=$ppt /"
done | tee -a gpurun_out/ab_g.log
for c in 4 6; do
MCRAT_B200_STREAM_CTAS_PER_SM=$c timeout 300 python tools/ab_compare.py $L:persistent_stream $L:persistent_stream C5 10000000 128 300 2>&1 | tail -2 | sed "s/^/ctas=$c /"
done | tee -a gpurun_out/ab_g.log
timeout 300 python tools/ab_compare.py $L:streamed $L:persistent_stream C5 5000000 64 300 2>&1 | tail -3 | tee -a gpurun_out/ab_g.log
