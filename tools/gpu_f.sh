#!/bin/bash
# round 2, GPU session F (2 GPUs): communicator tests, strong-scaling pair N=1 / N=2 of the default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
( time timeout 900 python -m pytest tests/test_comm.py -v -m gpu ) > gpurun_out/pytest_gpu_f.log 2>&1; grep -E "PASS|FAIL|SKIP|Error|error|assert" gpurun_out/pytest_gpu_f.log | cut -c1-220 | head -30; tail -3 gpurun_out/pytest_gpu_f.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 > gpurun_out/bench_f_2gpu.json 2> gpurun_out/bench_f_2gpu.err; echo "bench2 rc=$?"; tail -5 gpurun_out/bench_f_2gpu.err
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_f_1gpu.json 2> gpurun_out/bench_f_1gpu.err; echo "bench1 rc=$?"; tail -3 gpurun_out/bench_f_1gpu.err
python - <<P
import json
for f in ("gpurun_out/bench_f_1gpu.json","gpurun_out/bench_f_2gpu.json"):
    try:
        d=json.load(open(f))
        print(f, d["value"], d["ms_per_step"], d["k1_full_scan_ms"], d["loop_us_per_iteration"], d["roofline"]["frac"], d["pass_roofline"]["frac"], d["loop_roofline"]["frac"], d["e2e"]["value"])
        print(d["comm"])
    except Exception as e: print(f, "ERR", e)
P
