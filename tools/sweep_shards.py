#!/usr/bin/env python
"""Loop throughput (scatterings/s) of the C2 workload as a function of sub-shards per GPU."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcrat_b200 import HotPath, synth  # noqa: E402

nph = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
wl = sys.argv[3] if len(sys.argv) > 3 else "C2"
cfg, hydro, photons, frame = synth.workload(wl, n_photons=nph)
for S in [int(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "1,4,16,64,148,296,592,1024").split(",")]:
    hp = HotPath(cfg, seed=1, num_shards=S)
    hp.set_hydro(hydro)
    hp.set_photons(photons)
    st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=20, switch=1)
    hp.synchronize()
    t0 = time.perf_counter()
    st = hp.run_frame(st["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=0)
    hp.synchronize()
    dt = time.perf_counter() - t0
    print("S=%5d  shards=%d  iters=%d  scatterings=%d  %.1f us/iter  %.3e scatterings/s" %
          (S, hp.num_shards(), st["iterations"], st["scatterings"], 1e6 * dt / max(st["iterations"], 1),
           st["scatterings"] / dt), flush=True)
    hp.close()
