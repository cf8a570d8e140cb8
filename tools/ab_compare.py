#!/usr/bin/env python
"""A/B harness: run the same frames through two builds of the library (MCRAT_B200_LIB) and compare
the photon lists bit for bit, plus the loop time per iteration.

  python tools/ab_compare.py libA.so[:loop_mode[:recheck_skip]] libB.so[:loop_mode[:recheck_skip]] [workload] [photons] [shards] [iters] [scale]
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(out, wl, nph, shards, iters, scale):
    import numpy as np
    from mcrat_b200 import HotPath, synth
    cfg, hydro, photons, frame = synth.workload(wl, scale=scale, n_photons=nph, seed=5)
    mode = os.environ.get("MCRAT_LOOP_MODE")
    hp = HotPath(cfg, seed=99, num_shards=shards, scan_index=True, loop_mode=mode or None)
    if os.environ.get("MCRAT_RECHECK_SKIP"):
        hp.set_recheck_skip(int(os.environ["MCRAT_RECHECK_SKIP"]))
    hp.set_hydro(hydro)
    hp.set_photons(photons)
    st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=50, switch=1)
    hp.synchronize()
    t0 = time.perf_counter()
    st2 = hp.run_frame(st["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=0)
    hp.synchronize()
    dt = time.perf_counter() - t0
    np.save(out, hp.get_photons())
    print("%s[%s]: %d iterations %d scatterings  %.2f us/iteration  %.3e scatterings/s  launches %d" %
          (os.path.basename(os.environ.get("MCRAT_B200_LIB", "default")), os.environ.get("MCRAT_LOOP_MODE", "-") + "/" + os.environ.get("MCRAT_RECHECK_SKIP", "-"), st2["iterations"], st2["scatterings"],
           1e6 * dt / max(st2["iterations"], 1), st2["scatterings"] / dt, hp.launch_count()), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), float(sys.argv[7]))
        sys.exit(0)
    import numpy as np
    libs = sys.argv[1:3]
    wl = sys.argv[3] if len(sys.argv) > 3 else "C2"
    nph = sys.argv[4] if len(sys.argv) > 4 else "100000"
    shards = sys.argv[5] if len(sys.argv) > 5 else "16"
    iters = sys.argv[6] if len(sys.argv) > 6 else "2000"
    scale = sys.argv[7] if len(sys.argv) > 7 else "1.0"
    outs = []
    for k, lib in enumerate(libs):
        out = "/tmp/ab_%d.npy" % k
        lib, _, mode = lib.partition(":")
        mode, _, skip = mode.partition(":")
        env = dict(os.environ, MCRAT_B200_LIB=os.path.abspath(lib))
        if mode:
            env["MCRAT_LOOP_MODE"] = mode
        if skip:
            env["MCRAT_RECHECK_SKIP"] = skip  # 0 / 1 / 2 (mcrat_b200_set_recheck_skip)
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", out, wl, nph, shards, iters, scale], env=env)
        outs.append(np.load(out))
    a, b = outs
    same = all(np.array_equal(a[f], b[f], equal_nan=True) for f in a.dtype.names if a.dtype[f].kind == "f") and \
        all(np.array_equal(a[f], b[f]) for f in a.dtype.names if a.dtype[f].kind != "f")
    print("A/B %s %s photons %s shards: photon lists %s" % (wl, nph, shards, "BIT-IDENTICAL" if same else "DIFFER"))
    if not same:
        for f in a.dtype.names:
            if not np.array_equal(a[f], b[f], equal_nan=(a.dtype[f].kind == "f")):
                bad = np.nonzero(~((a[f] == b[f]) | ((a[f] != a[f]) & (b[f] != b[f]))))[0]
                print("  field %s differs at %d slots, first %s" % (f, bad.size, bad[:5]))
        sys.exit(1)
