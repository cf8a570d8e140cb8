#!/bin/bash
# round 2, GPU session U: last check of the final build -- suite, smoke, bench line
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu --maxfail=12 ) > gpurun_out/pytest_gpu_u.log 2>&1; tail -6 gpurun_out/pytest_gpu_u.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_u.err
python - <<P
import json
d=json.load(open("gpurun_out/bench_u.json"))
print(d["value"], d["ms_per_step"], d["k1_full_scan_ms"], d["loop_us_per_iteration"], d["roofline"]["frac"], d["pass_roofline"]["frac"], d["pass_roofline"]["ms_per_launch"], d["loop_roofline"]["frac"], d["e2e"]["value"])
print(d["cpu_baseline"]["value"] if d["cpu_baseline"] else None, d["gpu_launches"], d["clocks"])
P
