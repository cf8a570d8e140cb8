#!/bin/bash
# round 2, GPU session R2: photons per thread and item of the persistent stream against the list size
mkdir -p gpurun_out
L=mcrat_b200/csrc/libmcrat_b200.so
( for cfg in "10000000 128" "5000000 64" "2500000 32" "1250000 16"; do
  set -- $cfg
  for ppt in 4 8 12 16 24; do
    MCRAT_B200_STREAM_PPT=$ppt timeout 200 python tools/ab_compare.py $L:persistent_stream $L:persistent_stream C5 $1 $2 400 2>&1 | tail -2 | head -1 | sed "s/^/$1 $2 ppt=$ppt /"
  done
done ) 2>&1 | tee gpurun_out/ab_r2.log
