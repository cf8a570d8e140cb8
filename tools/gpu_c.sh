#!/bin/bash
# round 2, GPU session C: tests with the interleaved streamed loop, A/B at 1e7, bench line
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu --maxfail=12 > gpurun_out/pytest_gpu_c.log 2>&1; tail -15 gpurun_out/pytest_gpu_c.log
L=mcrat_b200/csrc/libmcrat_b200.so
for S in 128 32 2; do
timeout 600 python tools/ab_compare.py $L:streamed_global $L:streamed C5 10000000 $S 300 2>&1 | tail -3
done | tee gpurun_out/ab_c.log
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -4 gpurun_out/bench_c.err
python - <<P
import json
d=json.load(open("gpurun_out/bench_c.json"))
print(d["value"], d["ms_per_step"], d["k1_full_scan_ms"], d["loop_us_per_iteration"], d["roofline"]["frac"], d["pass_roofline"]["frac"], d["loop_roofline"]["frac"], d["e2e"]["value"])
print(d["s_sweep"])
P
