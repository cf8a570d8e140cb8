#!/usr/bin/env python
"""Small frames through every loop driver, for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool racecheck python tools/sanitize_driver.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcrat_b200 import HotPath, synth  # noqa: E402

for wl, scale, nph, shards, mode in (("C2", 1.0 / 32, 600, 2, "persistent"), ("C5", 1.0 / 16, 500, 1, "persistent"),
                                     ("C2", 1.0 / 32, 600, 200, "persistent"), ("C1", 1.0 / 16, 300, 2, "persistent"),
                                     ("C2", 1.0 / 32, 600, 2, "streamed")):
    cfg, hydro, photons, frame = synth.workload(wl, scale=scale, n_photons=nph, seed=3)
    hp = HotPath(cfg, seed=5, num_shards=shards, loop_mode=mode)
    hp.set_hydro(hydro)
    hp.set_photons(photons)
    st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=25, switch=1)
    got = hp.get_photons()
    print(wl, shards, mode, st["iterations"], st["scatterings"], st["relocations"], st["error"], flush=True)
    hp.close()
