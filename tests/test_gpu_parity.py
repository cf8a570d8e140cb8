"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle."""
import numpy as np
import pytest

from mcrat_b200 import HotPath, synth
from mcrat_b200.lib import RNG_REPLAY
from oracle import api, configs

from helpers import compare_photons

pytestmark = pytest.mark.gpu

CASES = [
    # workload, reference configuration, grid scale, photons, iterations
    ("C1", "c1_2d_cart", 1.0 / 8, 400, 300),
    ("C2", "c2_2d_cyl_stokes", 1.0 / 16, 400, 300),
    ("C5", "c5_3d_sph", 1.0 / 8, 400, 300),
]


def _oracle_frame(cfg, hydro, photons, frame, rng, iters, switch=1):
    o = api.Oracle(cfg)
    o.set_hydro(hydro)
    o.set_photons(photons)
    st = o.run_frame(rng, frame["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=switch)
    return o, st


@pytest.mark.parametrize("scan_index", [False, True])
@pytest.mark.parametrize("wl,refname,scale,nph,iters", CASES)
def test_frame_philox_parity(wl, refname, scale, nph, iters, scan_index):
    """Fused production path (Philox streams) vs the oracle drawing from the same keyed streams,
    with the full photon x cell scan and with the bounding-box index."""
    cfg, hydro, photons, frame = synth.workload(wl, scale=scale, n_photons=nph, seed=11)
    hp = HotPath(cfg, seed=2024, shard=3, scan_index=scan_index)
    hp.set_hydro(hydro)
    hp.set_photons(photons)
    st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=1)
    got = hp.get_photons()
    rng = api.OracleRng("philox", seed=2024, shard=3)
    o, ost = _oracle_frame(cfg, hydro, photons, frame, rng, iters)
    for k in ("iterations", "scatterings", "relocations", "photon_slots"):
        assert st[k] == ost[k], (k, st, ost)
    assert abs(st["time_now"] - ost["time_now"]) <= 1e-12 * abs(ost["time_now"])
    errs = compare_photons(got, o.photons(), label=wl, stokes_tol=1e-9, hydro=hydro)
    print(wl, "max rel errors", {k: "%.1e" % v for k, v in errs.items()})


@pytest.mark.parametrize("wl,refname,scale,nph,iters", CASES)
def test_frame_replay_parity(wl, refname, scale, nph, iters):
    """Replay harness: the uniform stream the reference consumed, fed to the GPU in reference order."""
    cfg, hydro, photons, frame = synth.workload(wl, scale=scale, n_photons=nph, seed=5)
    for seed in range(1, 50):
        src = api.OracleRng("ranlxs0", seed=seed)
        tee = src.tee(4_000_000)
        o, ost = _oracle_frame(cfg, hydro, photons, frame, src, iters)
        u = src.tee_values()
        if not np.any(u == 0.0):
            break
    hp = HotPath(cfg, rng_mode=RNG_REPLAY)
    hp.set_hydro(hydro)
    hp.set_photons(photons)
    hp.set_replay_uniforms(u)
    st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=1)
    got = hp.get_photons()
    assert hp.replay_consumed() == u.size, (hp.replay_consumed(), u.size)
    for k in ("iterations", "scatterings", "relocations"):
        assert st[k] == ost[k], (k, st, ost)
    errs = compare_photons(got, o.photons(), label=wl, stokes_tol=1e-9, hydro=hydro)
    print(wl, "max rel errors", {k: "%.1e" % v for k, v in errs.items()})


@pytest.mark.parametrize("wl,nph,shards,iters", [("C2", 1500, 6, 120), ("C5", 1000, 7, 100), ("C1", 640, 64, 40)])
def test_sub_shards_equal_independent_ranks(wl, nph, shards, iters):
    """S sub-shards in one context == S stand-alone oracle ranks, each with its slot range and its
    own Philox shard key (the reference's rank decomposition: shard-local arg-min, no exchange)."""
    scale = {"C1": 1.0 / 8, "C2": 1.0 / 16, "C5": 1.0 / 8}[wl]
    cfg, hydro, photons, frame = synth.workload(wl, scale=scale, n_photons=nph, seed=31)
    hp = HotPath(cfg, seed=77, shard=100, num_shards=shards)
    hp.set_hydro(hydro)
    hp.set_photons(photons)
    assert hp.num_shards() == shards or nph % shards
    st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=1)
    got = hp.get_photons()
    total_scatt = 0
    for s in range(hp.num_shards()):
        ss = hp.shard_stats(s)
        sl = slice(ss["first_slot"], ss["first_slot"] + ss["num_slots"])
        rng = api.OracleRng("philox", seed=77, shard=100 + s)
        o, ost = _oracle_frame(cfg, hydro, photons[sl], frame, rng, iters)
        assert ss["iterations"] == ost["iterations"] and ss["scatterings"] == ost["scatterings"], (s, ss, ost)
        assert abs(ss["time_now"] - ost["time_now"]) <= 1e-12 * ost["time_now"]
        compare_photons(got[sl], o.photons(), label="%s shard %d" % (wl, s), stokes_tol=1e-9, hydro=hydro)
        total_scatt += ost["scatterings"]
    assert st["scatterings"] == total_scatt
