#!/usr/bin/env python
"""Driver for ncu captures of the large-list loop: the bench's job (C5, 1e7 photons, 128 ranks) located through the
bounding-box index, then a few iterations of the STREAMED loop (pass_local_kernel / event_local_kernel: the same
pass_body / event_body the persistent stream runs, which ncu cannot capture because it serialises kernels).

  python tools/prof_loop.py [photons] [ranks] [iterations] [full_scan: 0|1]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcrat_b200 import HotPath, synth  # noqa: E402

nph = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
ranks = int(sys.argv[2]) if len(sys.argv) > 2 else 128
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 24
full = int(sys.argv[4]) if len(sys.argv) > 4 else 0

cfg, hydro, photons, frame = synth.workload("C5", n_photons=nph, seed=1234)
hp = HotPath(cfg, seed=20261018, num_shards=ranks, scan_index=not full, loop_mode="streamed")
hp.set_hydro(hydro)
hp.set_photons(photons)
st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=1)
print("located + %d iterations:" % iters, st)
if full:
    print("full rescan:", hp.rescan_all())
