#!/bin/bash
# round 2, GPU session R3: one rank per event block for short lists?
mkdir -p gpurun_out
L=mcrat_b200/csrc/libmcrat_b200.so
( for cfg in "1250000 16" "2500000 32"; do
  set -- $cfg
  timeout 200 python tools/ab_compare.py $L:persistent $L:persistent C5 $1 $2 400 2>&1 | tail -2 | head -1 | sed "s/^/$1 $2 team /"
  for ppt in 3 4 6 8 12; do
    MCRAT_B200_STREAM_EVT_BLOCKS=$2 MCRAT_B200_STREAM_PPT=$ppt timeout 200 python tools/ab_compare.py $L:persistent_stream $L:persistent_stream C5 $1 $2 400 2>&1 | tail -2 | head -1 | sed "s/^/$1 $2 E=$2 ppt=$ppt /"
  done
done ) 2>&1 | tee gpurun_out/ab_r3.log
