#!/usr/bin/env python
"""Per-config report of SURVEY 8(d): for each BASELINE config the rescan time and photon-cell evaluations/s (E_scan),
the steady-state photon-iterations/s (E_mfp) and scatterings/s (S) of the frame loop, and the roofline fractions
(FP64 issue rate measured on this GPU; HBM copy rate from MEASURED_PEAKS.json).  Prints a markdown table.

  python tools/config_report.py [iters] > gpurun_out/config_report.md
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcrat_b200 import HotPath, synth  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = peaks.get("hbm_gbs", 6650.0)

CASES = [
    # label, workload, photons, sub-shards, scan_index, iterations
    ("C1 2-D Cartesian spherical outflow, Stokes off, 1e4 photons, 1 rank", "C1", 10_000, 1, False, iters),
    ("C1, 16 sub-shards", "C1", 10_000, 16, False, iters),
    ("C2 2-D cylindrical jet, Stokes on, 1e5 photons, 16 sub-shards", "C2", 100_000, 16, False, iters),
    ("C2, 1 rank", "C2", 100_000, 1, False, iters),
    ("C3 = C2 + hot-electron table (built on the device)", "C3", 100_000, 16, False, max(iters // 4, 100)),
    ("C4 3-D spherical, cyclo-synchrotron switch on, 1e5 photons, 1 rank", "C4", 100_000, 1, False, max(iters // 4, 100)),
    ("C5 3-D spherical, 1e5 photons, 16 sub-shards", "C5", 100_000, 16, False, iters),
    ("C5, 3e5 photons", "C5", 300_000, 16, True, iters),
    ("C5, 1e6 photons", "C5", 1_000_000, 16, True, iters),
    ("C5, 3e6 photons (streamed loop)", "C5", 3_000_000, 16, True, max(iters // 10, 100)),
    ("C5, 1e7 photons (streamed loop)", "C5", 10_000_000, 16, True, max(iters // 10, 100)),
    ("C5, 1e7 photons, 128 ranks (bench default; persistent stream)", "C5", 10_000_000, 128, True, max(iters // 10, 100)),
    ("C5, 5e6 photons, 64 ranks (one of 2 GPUs)", "C5", 5_000_000, 64, True, max(iters // 10, 100)),
    ("C5, 2.5e6 photons, 32 ranks (one of 4 GPUs)", "C5", 2_500_000, 32, True, max(iters // 10, 100)),
    ("C5, 1.25e6 photons, 16 ranks (one of 8 GPUs)", "C5", 1_250_000, 16, True, iters),
    ("C5, 1e7 photons, 1 rank", "C5", 10_000_000, 1, True, max(iters // 10, 100)),
]

fp64 = None
hw = 148 * 64 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e9  # hardware FP64 issue peak, G thread-instr/s
rows = []
for label, wl, nph, shards, index, n_it in CASES:
    cfg, hydro, photons, frame = synth.workload(wl, n_photons=nph, seed=5)
    hp = HotPath(cfg, seed=99, num_shards=shards, scan_index=index)
    if fp64 is None:
        fp64 = hp.measure_fp64_peak()
    hp.set_hydro(hydro)
    t_table = None
    if cfg["tau_calculation"] == 2:
        t0 = time.perf_counter()
        hp.build_thermal_table(calls=200000, seed=3)
        hp.synchronize()
        t_table = time.perf_counter() - t0
    hp.set_photons(photons)
    if cfg["cyclosynch"]:
        hp.set_cs_limits(10 ** 9, 0)
    # rescan: the K1 kernel alone on the first 1e5 photons' worth of the list (full scan of every cell)
    ev, ms = (0, 0.0)
    if nph <= 1_000_000:
        hpk = HotPath(cfg, seed=99, num_shards=1)
        hpk.set_hydro(hydro)
        hpk.set_photons(photons)
        hpk.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=1, switch=1)
        ev, ms = hpk.rescan_all()
        ev, ms = hpk.rescan_all()
        hpk.close()
    st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=20, switch=1)
    hp.synchronize()
    t0 = time.perf_counter()
    st2 = hp.run_frame(st["time_now"], 1.0 / frame["fps"], max_iters=n_it, switch=0)
    hp.synchronize()
    dt = time.perf_counter() - t0
    instr = 6 if cfg["dimensions"] == synth.THREE else 4
    e_scan = ev / (ms * 1e-3) if ms > 0 else 0.0
    e_mfp = st2["photon_slots"] / dt
    rows.append((label, ms, e_scan, (e_scan * instr / 1e9 / hw) if ms > 0 else None, 1e6 * dt / max(st2["iterations"], 1),
                 e_mfp, e_mfp * 100.0 / 1e9 / hbm, st2["scatterings"] / dt, st2["relocations"], hp.launch_count(), t_table))
    hp.close()

print("FP64 pipe: hardware issue peak 148 SMs x 64 lanes x %.0f MHz = %.0f G instr/s (the 'of FP64 peak' column); DFMA probe on this GPU: "
      "%.0f G instr/s; HBM copy rate (MEASURED_PEAKS.json): %.1f GB/s\n" % (peaks.get("sm_max_mhz", 1965.0), hw, fp64, hbm))
print("| config | K1 rescan (ms) | E_scan (evals/s) | of FP64 peak | loop (us / iteration) | E_mfp (photon-iterations/s) | "
      "x 100 B of HBM | S (scatterings/s) | re-locations | launches |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for r in rows:
    print("| %s%s | %s | %s | %s | %.1f | %.3g | %.1f %% | %.3g | %d | %d |" % (
        r[0], (" (table: %.0f ms)" % (1e3 * r[10])) if r[10] else "", ("%.2f" % r[1]) if r[1] else "--",
        ("%.3g" % r[2]) if r[1] else "--", ("%.1f %%" % (100 * r[3])) if r[3] else "--", r[4], r[5], 100 * r[6], r[7], r[8], r[9]))
