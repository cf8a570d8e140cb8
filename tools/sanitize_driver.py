#!/usr/bin/env python
"""Small frames through every loop driver, for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool racecheck python tools/sanitize_driver.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcrat_b200 import HotPath, synth  # noqa: E402

for wl, scale, nph, shards, mode in (("C2", 1.0 / 32, 600, 2, "persistent"), ("C5", 1.0 / 16, 500, 1, "persistent"),
                                     ("C2", 1.0 / 32, 600, 200, "persistent"), ("C1", 1.0 / 16, 300, 2, "persistent"),
                                     ("C2", 1.0 / 32, 600, 2, "streamed"), ("C5", 1.0 / 16, 2000, 8, "persistent_stream"),
                                     ("C2", 1.0 / 32, 900, 3, "persistent_stream"), ("C2", 1.0 / 32, 600, 2, "cluster")):
    cfg, hydro, photons, frame = synth.workload(wl, scale=scale, n_photons=nph, seed=3)
    if mode == "cluster":  # the cooperative team's protocol through distributed shared memory
        os.environ["MCRAT_B200_CLUSTER_TEAM"] = "1"
        mode = "persistent"
    hp = HotPath(cfg, seed=5, num_shards=shards, loop_mode=mode)
    hp.set_hydro(hydro)
    hp.set_photons(photons)
    st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=25, switch=1)
    got = hp.get_photons()
    print(wl, shards, mode, st["iterations"], st["scatterings"], st["relocations"], st["error"], flush=True)
    hp.close()
    os.environ.pop("MCRAT_B200_CLUSTER_TEAM", None)

# the communicator on one rank (pack, counts, gather, table built through the all-gather path) and the full time order
from mcrat_b200 import Comm, comm_unique_id  # noqa: E402
cfg, hydro, photons, frame = synth.workload("C3", scale=1.0 / 32, n_photons=400, seed=3)
hp = HotPath(cfg, seed=5)
hp.set_hydro(hydro)
comm = Comm(hp, 1, 0, comm_unique_id())
comm.build_thermal_table(calls=200, seed=1)
hp.set_photons(photons)
st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=10, switch=1)
print("comm", comm.reduce_frame_stats(st)["scatterings"], comm.photon_counts()["list_capacity"], comm.gather_photons(root=0)[1], flush=True)
hp.findContainingHydroCell(0)
hp.calcMeanFreePath()
print("sorted", hp.sortedIndexes()[:4], flush=True)
comm.close()
hp.close()
