#!/bin/bash
# round 2, GPU session E: state of HEAD -- full GPU suite, the default bench line (both arms), ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
( time timeout 1500 python -m pytest tests -q -m gpu --maxfail=12 ) > gpurun_out/pytest_gpu_e.log 2>&1; tail -12 gpurun_out/pytest_gpu_e.log
timeout 900 python bench.py > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; echo "bench rc=$?"; tail -4 gpurun_out/bench_e.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_e_ref.json 2> gpurun_out/bench_e_ref.err; echo "ref rc=$?"; tail -2 gpurun_out/bench_e_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_e.csv \
  python bench.py --steps 1 --warmup 1 --iters 100 --no-cpu-baseline --no-sweep --no-e2e > gpurun_out/ncu_e.log 2>&1; echo "ncu rc=$?"
python - <<P
import json
d=json.load(open("gpurun_out/bench_e.json"))
print(d["value"], d["ms_per_step"], d["k1_full_scan_ms"], d["loop_us_per_iteration"], d["roofline"]["frac"], d["pass_roofline"]["frac"], d["loop_roofline"]["frac"], d["e2e"])
print(d["s_sweep"]); print(d["cpu_baseline"]["value"] if d["cpu_baseline"] else None, d["gpu_launches"], d["clocks"])
P
