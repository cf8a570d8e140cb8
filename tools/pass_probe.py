#!/usr/bin/env python
"""Per-launch time of the pass kernel in profile mode, after different loop drivers have run (diagnosis)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcrat_b200 import HotPath, synth

cfg, hydro, photons, frame = synth.workload("C5", n_photons=10_000_000, seed=1234)
for first, iters in (("none", 0), ("streamed", 300), ("persistent_stream", 300), ("persistent_stream", 3000)):
    hp = HotPath(cfg, seed=20261018, num_shards=128, scan_index=True, loop_mode="auto" if first == "none" else first)
    hp.set_hydro(hydro)
    hp.set_photons(photons)
    st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=20, switch=1)
    if iters:
        st = hp.run_frame(st["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=0)
    for rep in range(2):
        hp.set_profile(True)
        hp.kernel_times(reset=True)
        st = hp.run_frame(st["time_now"], 1.0 / frame["fps"], max_iters=24, switch=0)
        kt = hp.kernel_times(reset=True)
        hp.set_profile(False)
        print("after %-18s x %4d, rep %d: pass %.1f us x %d launches, event %.1f us, relocations %d" %
              (first, iters, rep, 1e3 * kt["pass_ms"] / max(kt["pass_launches"], 1), kt["pass_launches"],
               1e3 * kt["event_ms"] / max(kt["event_launches"], 1), st["relocations"]), flush=True)
    hp.close()
