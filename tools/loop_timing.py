#!/usr/bin/env python
"""Where one iteration of the persistent loop spends its time (SM cycles, shard 0): needs the
instrumented build  make -C mcrat_b200/csrc libmcrat_b200_timing.so

  python tools/loop_timing.py [workload] [photons] [shards] [iters]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MCRAT_B200_LIB"] = os.path.join(ROOT, "mcrat_b200", "csrc", "libmcrat_b200_timing.so")
from mcrat_b200 import HotPath, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
nph = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
shards = int(sys.argv[3]) if len(sys.argv) > 3 else 16
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
cfg, hydro, photons, frame = synth.workload(wl, n_photons=nph, seed=5)
hp = HotPath(cfg, seed=99, num_shards=shards, scan_index=True, loop_mode="persistent")
hp.set_hydro(hydro)
if cfg["tau_calculation"] == 2:
    hp.build_thermal_table(calls=200000, seed=3)
hp.set_photons(photons)
st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=50, switch=1)
buf = (C.c_longlong * 32)()
hp.L.mcrat_b200_debug_counters(hp.ctx, buf, 1)
import time
hp.synchronize()
t0 = time.perf_counter()
st = hp.run_frame(st["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=0)
hp.synchronize()
wall = time.perf_counter() - t0
hp.L.mcrat_b200_debug_counters(hp.ctx, buf, 0)
n = max(st["iterations"], 1)
print("%.2f us per iteration end to end (instrumented build)" % (1e6 * wall / n))
names = {8: "scattering lane: loads (idx, temp, comv p), Philox prefetch, warp-wide Box-Muller", 9: "electron direction + rotation",
         10: "boost into the electron frame", 11: "wait (helper / Stokes warps)", 12: "Klein-Nishina accept + polar angle",
         13: "wait (Stokes q,u; alignment)", 14: "azimuth + outgoing photon", 15: "wait", 16: "boosts back",
         17: "wait (angles, Fano)", 18: "write-back"}
print("%s %d photons %d shards: SM-cycle stamps of shard 0's scattering lane, per iteration" % (wl, nph, shards))
for k in sorted(names):
    print("  %-86s %9.0f cycles / iteration  (%.2f us at 1.965 GHz)" % (names[k], buf[k] / n, buf[k] / n / 1965.0))
