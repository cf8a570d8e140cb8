#!/bin/bash
# round 2, GPU session S: the final build -- suite, smoke, both bench arms, per-config report
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu --maxfail=12 ) > gpurun_out/pytest_gpu_s.log 2>&1; tail -6 gpurun_out/pytest_gpu_s.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_s.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_s_ref.json 2> gpurun_out/bench_s_ref.err; echo "ref rc=$?"
timeout 900 python tools/config_report.py 2000 > gpurun_out/config_report_r02.md 2> gpurun_out/config_report.err; echo "report rc=$?"; tail -20 gpurun_out/config_report_r02.md
python - <<P
import json
d=json.load(open("gpurun_out/bench_s.json"))
print(d["value"], d["ms_per_step"], d["k1_full_scan_ms"], d["loop_us_per_iteration"], d["roofline"]["frac"], d["roofline"]["traffic"], d["pass_roofline"]["frac"], d["pass_roofline"]["traffic"], d["loop_roofline"]["frac"], d["e2e"])
print(d["s_sweep"]); print(d["cpu_baseline"]["value"] if d["cpu_baseline"] else None, d["gpu_launches"], d["clocks"])
r=json.load(open("gpurun_out/bench_s_ref.json")); print(r["value"], r["ms_per_step"])
P
