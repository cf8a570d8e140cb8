#!/usr/bin/env python
"""Soak of the skipped cell re-checks in verified mode (mcrat_b200_set_recheck_skip(ctx, 2)): every skip decision is
checked against the real re-check on the device, for full-size lists over several hydro frames, dense and optically thin
flows.  Any wrong skip fails the frame with MCRAT_B200_ERR_STATE.

  python tools/skip_soak.py [frames] [iterations per frame]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcrat_b200 import HotPath, synth  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
CASES = [("C2", 100_000, 16, 1.0), ("C2", 100_000, 16, 1e-5), ("C5", 1_000_000, 16, 1.0), ("C5", 300_000, 148, 1e-4),
         ("C1", 10_000, 1, 1.0), ("G3P", 100_000, 16, 1e-4), ("G2S", 100_000, 16, 1e-4), ("G3C", 100_000, 16, 1e-4),
         ("C5", 4_000_000, 4, 1e-3)]
for wl, nph, shards, dilute in CASES:
    cfg, hydro, photons, frame = synth.workload(wl, n_photons=nph, seed=3)
    if dilute != 1.0:
        hydro = dict(hydro)
        hydro["dens"] = np.asarray(hydro["dens"]) * dilute
        hydro["dens_lab"] = np.asarray(hydro["dens_lab"]) * dilute
    hp = HotPath(cfg, seed=11, num_shards=shards, scan_index=True)
    hp.set_recheck_skip(2)
    hp.set_photons(photons)
    t, tot = frame["time_now"], [0, 0, 0]
    t0 = time.perf_counter()
    for f in range(frames):
        hp.set_hydro(hydro)
        n_it = iters if nph <= 1_000_000 else max(iters // 20, 50)
        st = hp.run_frame(t, 1.0 / frame["fps"], max_iters=n_it, switch=1)
        st2 = hp.run_frame(st["time_now"], 1.0 / frame["fps"], max_iters=n_it, switch=0)
        t = st2["time_now"]
        for s in (st, st2):
            tot[0] += s["iterations"]; tot[1] += s["scatterings"]; tot[2] += s["relocations"]
    hp.synchronize()
    print("%-4s %8d photons %4d shards dilution %-6g: %d frames, %d iterations, %d scatterings, %d re-locations, every skip verified  (%.1f s)"
          % (wl, nph, shards, dilute, frames, tot[0], tot[1], tot[2], time.perf_counter() - t0), flush=True)
    hp.close()
