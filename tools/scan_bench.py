#!/usr/bin/env python
"""Time the full photon x cell scan (K1) alone: `scan_bench.py [C2|C5] [photons]` (default 1e5 photons x 1 048 576 cells).
MCRAT_B200_LIB selects an alternative build of the library (A/B of kernel variants)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcrat_b200 import HotPath, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
nph = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100000
cfg, hydro, photons, frame = synth.workload(wl, n_photons=nph)
hp = HotPath(cfg, seed=7)
hp.set_hydro(hydro)
hp.set_photons(photons)
ms = []
for k in range(4):
    ev, t = hp.rescan_all()
    ms.append(t)
best = min(ms[1:])
peak = hp.measure_fp64_peak()
hw = 148 * 64 * 1.965  # G thread-instr/s: 148 SMs x 64 FP64 lanes x 1965 MHz
ipe = 6 if cfg["dimensions"] == synth.THREE else 4
g = ev * ipe / (best * 1e-3) / 1e9
print("lib=%s %s %d photons: %.3f ms (%s) %.3e evals/s  %.0f G FP64 instr/s = %.1f%% of the DFMA probe (%.0f), %.1f%% of %.0f (hardware)" %
      (os.path.basename(os.environ.get("MCRAT_B200_LIB", "default")), wl, nph, best, " ".join("%.2f" % m for m in ms),
       ev / (best * 1e-3), g, 100 * g / peak, peak, 100 * g / hw, hw))
