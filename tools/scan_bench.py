#!/usr/bin/env python
"""Time the full photon x cell scan (K1) alone: C2 = 1e5 photons x 1,048,576 cells."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcrat_b200 import HotPath, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
cfg, hydro, photons, frame = synth.workload(wl)
hp = HotPath(cfg, seed=7)
hp.set_hydro(hydro)
hp.set_photons(photons)
ms = []
for k in range(5):
    ev, t = hp.rescan_all()
    ms.append(t)
best = min(ms[1:])
peak = hp.measure_fp64_peak()
ipe = 6 if cfg["dimensions"] == 2 else 4
print("lib=%s %s: %.3f ms (%s) %.3e evals/s  %.0f G FP64 instr/s = %.1f%% of %.0f (DFMA rate)" %
      (os.path.basename(os.environ.get("MCRAT_B200_LIB", "default")), wl, best, " ".join("%.2f" % m for m in ms),
       ev / (best * 1e-3), ev * ipe / (best * 1e-3) / 1e9, 100 * ev * ipe / (best * 1e-3) / 1e9 / peak, peak))
