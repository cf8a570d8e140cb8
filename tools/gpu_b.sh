#!/bin/bash
# round 2, GPU session B: FP64 issue micro-benchmark, K1 variants
mkdir -p gpurun_out
./tools/micro/fp64_issue > gpurun_out/fp64_issue.txt 2>&1; cat gpurun_out/fp64_issue.txt
for v in ${VARIANTS:-v7 v8 v9 v10 v11}; do
  for args in "C5 1e5" "C5 1e6" "C2 1e5"; do
    MCRAT_B200_LIB=$PWD/mcrat_b200/csrc/libmcrat_b200_$v.so timeout 300 python tools/scan_bench.py $args 2>&1 | tail -1
  done
done | tee gpurun_out/scan_variants.txt
