#!/bin/bash
# round 2, GPU session X (4 GPUs): the default bench split over four GPUs
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus 4 --no-cpu-baseline > gpurun_out/bench_x_4gpu.json 2> gpurun_out/bench_x_4gpu.err; echo "bench4 rc=$?"; tail -3 gpurun_out/bench_x_4gpu.err
python - <<P
import json
d=json.load(open("gpurun_out/bench_x_4gpu.json"))
print(d["value"], d["ms_per_step"], d["ms_per_step_per_gpu"], d["k1_full_scan_ms"], d["loop_us_per_iteration"], d["roofline"]["frac"], d["loop_roofline"]["frac"], d["e2e"]["value"])
print(d["comm"])
P
