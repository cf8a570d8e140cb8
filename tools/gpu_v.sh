#!/bin/bash
# round 2, GPU session V: AUTO after the last tuning -- suite, the four per-GPU shares of the job, bench line
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu --maxfail=12 ) > gpurun_out/pytest_gpu_v.log 2>&1; tail -5 gpurun_out/pytest_gpu_v.log
L=mcrat_b200/csrc/libmcrat_b200.so
( for cfg in "10000000 128" "5000000 64" "2500000 32" "1250000 16" "1000000 16" "300000 16"; do
  set -- $cfg
  timeout 200 python tools/ab_compare.py $L:streamed $L:auto C5 $1 $2 400 2>&1 | tail -3 | sed "s/^/$1 $2 /"
done ) 2>&1 | tee gpurun_out/ab_v.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; echo "bench rc=$?"
python - <<P
import json
d=json.load(open("gpurun_out/bench_v.json"))
print(d["value"], d["ms_per_step"], d["k1_full_scan_ms"], d["loop_us_per_iteration"], d["roofline"]["frac"], d["pass_roofline"]["frac"], d["loop_roofline"]["frac"], d["e2e"]["value"])
P
