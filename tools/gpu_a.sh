#!/bin/bash
# round 2, GPU session A: tests, first C5@1e7 bench line, persistent-vs-streamed experiment at 1e7
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu --maxfail=12 > gpurun_out/pytest_gpu_a.log 2>&1; tail -15 gpurun_out/pytest_gpu_a.log
timeout 900 python bench.py --steps 3 --warmup 1 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -3 gpurun_out/bench_a.err; tail -c 3000 gpurun_out/bench_a.json
for mode in streamed persistent; do
timeout 300 python bench.py --steps 2 --warmup 1 --ranks 64 --loop $mode --index --no-cpu-baseline --no-e2e --no-sweep --iters 1000 > gpurun_out/bench_a_$mode.json 2> gpurun_out/bench_a_$mode.err; tail -2 gpurun_out/bench_a_$mode.err
python - <<P
import json
d=json.load(open("gpurun_out/bench_a_$mode.json"))
print("$mode", d["loop_us_per_iteration"], d["pass_roofline"]["ms_per_launch"], d["loop_roofline"]["frac"])
P
done
