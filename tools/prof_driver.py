#!/usr/bin/env python
"""Short driver for ncu captures: one C2 frame slice (full rescan + a few loop iterations),
then a few fused passes over a photon list larger than L2.  Run plainly first, then under ncu
(see profiles/README.md for the exact commands)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcrat_b200 import HotPath, synth  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
big = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
shards = int(sys.argv[3]) if len(sys.argv) > 3 else 16
loop = sys.argv[4] if len(sys.argv) > 4 else "auto"
wl = sys.argv[5] if len(sys.argv) > 5 else "C2"

cfg, hydro, photons, frame = synth.workload(wl)
hp = HotPath(cfg, seed=1, num_shards=shards, loop_mode=loop)
hp.set_hydro(hydro)
hp.set_photons(photons)
st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=1)
print(wl, "slice:", st)
if big > 0:
    hp.set_photons(np.resize(photons, big))
    st = hp.run_frame(st["time_now"], 1.0 / frame["fps"], max_iters=8, switch=0)
    print("big list:", st)
