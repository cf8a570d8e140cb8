#!/bin/bash
# round 2, GPU session P (2 GPUs): communicator tests on two ranks, the default bench split over two GPUs
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
( time timeout 900 python -m pytest tests/test_comm.py -v -m gpu ) > gpurun_out/pytest_gpu_p.log 2>&1; grep -E "PASS|FAIL|SKIP|rror" gpurun_out/pytest_gpu_p.log | cut -c1-200 | head -20; tail -3 gpurun_out/pytest_gpu_p.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 > gpurun_out/bench_p_2gpu.json 2> gpurun_out/bench_p_2gpu.err; echo "bench2 rc=$?"; tail -4 gpurun_out/bench_p_2gpu.err
python - <<P
import json
d=json.load(open("gpurun_out/bench_p_2gpu.json"))
print(d["value"], d["ms_per_step"], d["ms_per_step_per_gpu"], d["k1_full_scan_ms"], d["loop_us_per_iteration"], d["roofline"]["frac"], d["loop_roofline"]["frac"], d["e2e"]["value"])
print(d["comm"])
P
