#!/bin/bash
# round 2, GPU session M: evidence on the current build -- suite, both bench arms, ncu launch list, ncu --set full captures,
# per-config report
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
( time timeout 1500 python -m pytest tests -q -m gpu --maxfail=12 ) > gpurun_out/pytest_gpu_m.log 2>&1; tail -6 gpurun_out/pytest_gpu_m.log
timeout 900 python bench.py > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_m.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_m_ref.json 2> gpurun_out/bench_m_ref.err; echo "ref rc=$?"
# launch list of the bench command (reduced iteration count; ncu serialises kernels, so the persistent pair is called off by its
# handshake and the loop shows as the streamed loop's kernels -- same pass / event code)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_m.csv \
  python bench.py --steps 1 --warmup 1 --iters 100 --no-cpu-baseline --no-sweep --no-e2e > gpurun_out/ncu_m.log 2>&1; echo "ncu list rc=$?"
# --set full: K1 3-D (final build), the pass / event kernels at 1e7 photons (streamed loop), the team kernel (C2)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan_kernel" -c 1 -f -o gpurun_out/prof_r02_scan3d \
  python tools/prof_driver.py 50 0 128 auto C5 > gpurun_out/ncu_full_scan.log 2>&1; echo "ncu scan rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pass_local_kernel|event_local_kernel" --launch-skip 8 --launch-count 2 \
  -f -o gpurun_out/prof_r02_loop1e7 python tools/prof_driver.py 2 10000000 128 streamed C5 > gpurun_out/ncu_full_loop.log 2>&1; echo "ncu loop rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"frame_loop_kernel" -c 1 -f -o gpurun_out/prof_r02_team \
  python tools/prof_driver.py 200 0 16 auto C2 > gpurun_out/ncu_full_team.log 2>&1; echo "ncu team rc=$?"
timeout 900 python tools/config_report.py 2000 > gpurun_out/config_report_r02.md 2> gpurun_out/config_report.err; echo "report rc=$?"; tail -22 gpurun_out/config_report_r02.md
python - <<P
import json
d=json.load(open("gpurun_out/bench_m.json"))
print(d["value"], d["ms_per_step"], d["k1_full_scan_ms"], d["loop_us_per_iteration"], d["roofline"]["frac"], d["pass_roofline"]["frac"], d["loop_roofline"]["frac"], d["e2e"])
print(d["s_sweep"]); print(d["cpu_baseline"]["value"] if d["cpu_baseline"] else None, d["gpu_launches"], d["clocks"]); print(d["comm"])
P
ls -la gpurun_out/*.ncu-rep
