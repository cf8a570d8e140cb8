#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ FROM THE REFERENCE'S OWN SOURCES.

Needs oracle/_ref (built by oracle/build_ref.py from /root/reference/Src, present only in
the build container).  For every configuration a small synthetic frame is replayed through the
reference's functions (findContainingHydroCell / calcMeanFreePath / photonEvent /
updatePhotonPosition driven by the loop of Src/mcrat.c:754-851) with a seeded ranlxs0 stream;
inputs, the uniform stream the reference consumed, and the resulting photon list are stored.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from mcrat_b200 import hotxs, synth  # noqa: E402
from oracle import api, configs  # noqa: E402

HYDRO_KEYS = api.HYDRO_FIELDS + ["r0_domain", "r1_domain", "r2_domain"]

# name -> (reference configuration, workload, grid scale, photons, iterations)
CASES = {
    "c1_2d_cart": ("c1_2d_cart", "C1", 1.0 / 32, 96, 80),
    "c2_2d_cyl_stokes": ("c2_2d_cyl_stokes", "C2", 1.0 / 64, 96, 80),
    "c3_2d_cyl_table": ("c3_2d_cyl_table", "C3", 1.0 / 64, 96, 80),
    "c5_3d_sph": ("c5_3d_sph", "C5", 1.0 / 16, 96, 80),
}


def table():
    path = os.path.join(HERE, "thermal_table.npy")
    if not os.path.exists(path):
        np.save(path, hotxs.build_table(order=32))
    return np.load(path)


def main():
    tab = table()
    for name, (refname, wl, scale, nph, iters) in CASES.items():
        cfg, hydro, photons, frame = synth.workload(wl, scale=scale, n_photons=nph, seed=99)
        ref = api.RefLib(refname)
        if configs.CONFIGS[refname]["tau_calculation"] == configs.TABLE:
            ref.set_table(tab)
        for seed in range(1, 100):
            ref.set_hydro(hydro)
            ref.set_photons(photons)
            rng, tee = ref.new_rng(seed=seed, tee=2_000_000)
            st = ref.run_frame(rng, frame["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=1)
            u = ref.tee_values(rng, tee)
            if not np.any(u == 0.0):
                break
        out = ref.photons()
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            photons_in=photons, photons_out=out, uniforms=u, seed=seed, iters=iters,
                            time_now=frame["time_now"], dt=1.0 / frame["fps"], stats=np.array(sorted(st.items()), dtype=object),
                            num_elements=hydro["num_elements"], fps=hydro["fps"],
                            **{k: np.asarray(hydro[k]) for k in HYDRO_KEYS})
        print(name, hydro["num_elements"], "cells", nph, "photons", st, "uniforms", u.size)


def _index_chunk(args):
    refname, hydro, ph = args
    ref = api.RefLib(refname)
    ref.set_hydro(hydro)
    ref.set_photons(ph)
    rng, _ = ref.new_rng(seed=1)
    ref.find_containing_hydro_cell(1, rng)
    return ref.photons()["nearest_block_index"].copy()


def index_goldens():
    """Full BASELINE-size cell indices (1e5 photons x 1 048 576 cells) from the reference's own findContainingHydroCell
    (switch = 1): tests/golden/index_full_<cfg>.npz holds the 1e5 int32 indices and a hash of the grid geometry."""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    for name in helpers.INDEX_GOLDEN:
        cfg, hydro, ph, frame, refname, geo = helpers.index_golden_inputs(name)
        h = {k: hydro[k] for k in HYDRO_KEYS + ["num_elements", "fps"]}
        nproc = os.cpu_count() or 1
        chunks = np.array_split(np.arange(ph.size), nproc * 4)
        with mp.get_context("fork").Pool(nproc) as pool:
            parts = pool.map(_index_chunk, [(refname, h, ph[c]) for c in chunks])
        idx = np.concatenate(parts).astype(np.int32)
        np.savez_compressed(os.path.join(HERE, "index_full_%s.npz" % name), idx=idx, geometry_sha256=geo)
        print("index_full_%s: %d photons, %d outside (-1), max index %d" % (name, idx.size, int((idx < 0).sum()), idx.max()))


if __name__ == "__main__":
    if "--index" in sys.argv:
        index_goldens()
    else:
        main()
        index_goldens()
