#!/usr/bin/env python
"""Time the fused pass (K4+K2) alone on a photon list larger than L2 (HBM-bound regime)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcrat_b200 import HotPath, synth  # noqa: E402

nbig = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
shards = int(sys.argv[2]) if len(sys.argv) > 2 else 1
distinct = len(sys.argv) > 3 and sys.argv[3] == "distinct"  # nbig different photons (cells spread) instead of 1e5 repeated
wl = sys.argv[4] if len(sys.argv) > 4 else "C2"
cfg, hydro, photons, frame = synth.workload(wl, n_photons=nbig if distinct else None)
hp = HotPath(cfg, seed=7, profile=True, num_shards=shards)
hp.set_hydro(hydro)
hp.set_photons(photons if distinct else np.resize(photons, nbig))
if os.environ.get("MCRAT_RECHECK_SKIP"):
    hp.set_recheck_skip(int(os.environ["MCRAT_RECHECK_SKIP"]))
# the rescan iteration and two more: the pass after a re-location re-checks every photon and sets its skip threshold
st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=3, switch=1)
hp.kernel_times(reset=True)
st = hp.run_frame(st["time_now"], 1.0 / frame["fps"], max_iters=24, switch=0)
kt = hp.kernel_times()
ms = kt["pass_ms"] / kt["pass_launches"]
print("%s lib=%s photons=%d shards=%d pass %.2f us/launch -> %.0f GB/s (100 B/photon-iteration); event %.1f us; other %.1f us" %
      (wl, os.environ.get("MCRAT_B200_LIB", "default"), nbig, hp.num_shards(), 1e3 * ms, nbig * 100.0 / (ms * 1e-3) / 1e9,
       1e3 * kt["event_ms"] / max(kt["event_launches"], 1), 1e3 * kt["other_ms"] / max(kt["other_launches"], 1)))
