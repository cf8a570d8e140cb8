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
    # the other coordinate systems of Src/geometry.c: 2.5-D cylindrical, 2-D spherical, 3-D Cartesian, 3-D polar
    ("G25", "g_25d_cyl", 1.0 / 8, 400, 300),
    ("G2S", "g_2d_sph", 1.0 / 8, 400, 300),
    ("G3C", "g_3d_cart", 1.0 / 4, 400, 300),
    ("G3P", "g_3d_polar", 1.0 / 4, 400, 300),
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
    errs = compare_photons(got, o.photons(), label=wl, hydro=hydro)
    print(wl, "max rel errors", {k: "%.1e" % v for k, v in errs.items()})


LARGE_CASES = [
    # workload, grid scale (cells), photons, sub-shards, iterations per shard: tens of seconds of oracle time each
    ("C2", 1.0 / 4, 24000, 4, 1500),     # 65 536 cells, 6 000 scatterings in all
    ("C5", 1.0 / 2, 20000, 4, 1200),     # 131 072 cells, 3-D spherical
    ("C3", 1.0 / 8, 12000, 3, 600),      # hot electrons + table, a third of the candidates rejected
]


@pytest.mark.parametrize("wl,scale,nph,shards,iters", LARGE_CASES)
def test_longer_frames_on_larger_grids_hold_the_same_tolerances(wl, scale, nph, shards, iters):
    """The per-field tolerances of tests/helpers.py were measured on 400-photon, 300-iteration frames.  Here: tens of
    thousands of photons, thousands of scatterings per rank, every sub-shard against its own oracle rank -- errors do
    not accumulate beyond them (a photon's error is reset by nothing, but each scattering re-derives its momenta from
    freshly drawn angles, and positions are sums of exact pushes)."""
    cfg, hydro, photons, frame = synth.workload(wl, scale=scale, n_photons=nph, seed=97)
    hp = HotPath(cfg, seed=4711, shard=20, num_shards=shards, scan_index=True)
    hp.set_hydro(hydro)
    table = None
    if wl == "C3":
        table, _ = hp.build_thermal_table(calls=20000, seed=3)
    hp.set_photons(photons)
    st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=1)
    got = hp.get_photons()
    total = 0
    for s in range(hp.num_shards()):
        ss = hp.shard_stats(s)
        sl = slice(ss["first_slot"], ss["first_slot"] + ss["num_slots"])
        o = api.Oracle(cfg)
        o.set_hydro(hydro)
        if table is not None:
            o.set_table(table)
        o.set_photons(photons[sl])
        ost = o.run_frame(api.OracleRng("philox", seed=4711, shard=20 + s), frame["time_now"], 1.0 / frame["fps"],
                          max_iters=iters, switch=1)
        assert ss["iterations"] == ost["iterations"] and ss["scatterings"] == ost["scatterings"], (s, ss, ost)
        compare_photons(got[sl], o.photons(), label="%s large shard %d" % (wl, s), hydro=hydro)
        total += ost["scatterings"]
    assert st["scatterings"] == total and total > 0.5 * shards * iters


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
    errs = compare_photons(got, o.photons(), label=wl, hydro=hydro)
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
        compare_photons(got[sl], o.photons(), label="%s shard %d" % (wl, s), hydro=hydro)
        total_scatt += ost["scatterings"]
    assert st["scatterings"] == total_scatt


def _thin(hydro, factor):
    """The same flow, `factor` times more dilute: free paths and therefore event time steps grow by 1/factor, so
    that many photons change cell in every loop iteration (the regime the persistent loop hands back to K1b / K1c)."""
    h = dict(hydro)
    h["dens"] = np.asarray(hydro["dens"]) * factor
    h["dens_lab"] = np.asarray(hydro["dens_lab"]) * factor
    return h


LOOP_CASES = [
    # workload, grid scale, photons, sub-shards, iterations, dilution, scan_index
    ("C2", 1.0 / 16, 3000, 1, 150, 1.0, False),
    ("C2", 1.0 / 16, 3000, 5, 150, 1.0, True),
    ("C5", 1.0 / 8, 4000, 40, 60, 1.0, False),
    ("C1", 1.0 / 8, 2000, 3, 120, 1.0, False),
    ("C2", 1.0 / 4, 3000, 2, 30, 5e-6, False),     # optically thin: ~100 re-locations per shard and iteration
    ("C5", 1.0 / 8, 20000, 320, 30, 1.0, True),     # more shards than resident wide blocks: four-warp blocks
]


@pytest.mark.parametrize("wl,scale,nph,shards,iters,dilute,scan_index", LOOP_CASES)
def test_persistent_loop_is_bit_identical_to_streamed_loop_and_matches_oracle(wl, scale, nph, shards, iters, dilute, scan_index):
    """One cooperative launch per frame (frame_loop_kernel) vs the interleaved two-stream loop vs four grid-wide launches
    per iteration: same photons bit for bit, same counters; and sub-shard 0 against the oracle.  Covers the hand-back to the streamed loop when a shard
    re-locates more than RELOC_HEAVY photons per iteration, frames that end inside the call, and max_iters stops."""
    cfg, hydro, photons, frame = synth.workload(wl, scale=scale, n_photons=nph, seed=41)
    if dilute != 1.0:
        hydro = _thin(hydro, dilute)
    out = {}
    for mode in ("streamed_global", "streamed", "persistent", "persistent_stream"):
        hp = HotPath(cfg, seed=5150, shard=7, num_shards=shards, scan_index=scan_index, loop_mode=mode)
        hp.set_hydro(hydro)
        hp.set_photons(photons)
        st1 = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=1)
        st2 = hp.run_frame(st1["time_now"], 1.0 / frame["fps"] - (st1["time_now"] - frame["time_now"]), max_iters=iters // 2,
                           switch=0)
        out[mode] = (st1, st2, hp.get_photons(), [hp.shard_stats(s) for s in range(hp.num_shards())], hp.launch_count())
        hp.close()
    a = out["streamed_global"]
    # two-stream interleaved halves; one cooperative launch; resident event blocks beside a stream of pass items (the
    # loop of lists larger than L2, forced onto these small ones; with more sub-shards than SMs it is the streamed loop)
    for mode in ("streamed", "persistent", "persistent_stream"):
        b = out[mode]
        for k in ("iterations", "scatterings", "relocations", "photon_slots", "not_found", "time_now", "last_time_step",
                  "last_scattered_index"):
            assert a[0][k] == b[0][k] and a[1][k] == b[1][k], (mode, k, a[0], b[0], a[1], b[1])
        for f in a[2].dtype.names:
            assert np.array_equal(a[2][f], b[2][f], equal_nan=(a[2].dtype[f].kind == "f")), "%s: field %s differs" % (mode, f)
        for sa, sb in zip(a[3], b[3]):
            for k in ("iterations", "scatterings", "relocations", "time_now"):
                assert sa[k] == sb[k], (mode, k, sa, sb)
    b = out["persistent"]
    if dilute == 1.0:
        assert b[4] < a[4]  # far fewer launches
    else:
        assert a[0]["relocations"] > 64 * iters / 4, "the thin case must actually exercise heavy re-location: %s" % a[0]
    # sub-shard 0 against the oracle (two consecutive calls, the second continuing the frame)
    ss = b[3][0]
    sl = slice(ss["first_slot"], ss["first_slot"] + ss["num_slots"])
    o = api.Oracle(cfg)
    o.set_hydro(hydro)
    o.set_photons(photons[sl])
    rng = api.OracleRng("philox", seed=5150, shard=7)
    o1 = o.run_frame(rng, frame["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=1)
    o2 = o.run_frame(rng, o1["time_now"], 1.0 / frame["fps"] - (o1["time_now"] - frame["time_now"]), max_iters=iters // 2, switch=0)
    assert ss["scatterings"] == o1["scatterings"] + o2["scatterings"]
    compare_photons(b[2][sl], o.photons(), label="%s persistent shard 0" % wl, hydro=hydro)


def test_persistent_loop_runs_a_frame_to_its_end():
    """No iteration cap: every shard stops when its own clock reaches the next hydro frame (Src/mcrat.c:834-846)."""
    cfg, hydro, photons, frame = synth.workload("C2", scale=1.0 / 16, n_photons=600, seed=43)
    hydro = _thin(hydro, 3e-4)  # a few hundred scatterings per shard and frame instead of millions
    res = {}
    for mode in ("streamed", "persistent", "persistent_stream"):
        hp = HotPath(cfg, seed=99, num_shards=4, loop_mode=mode)
        hp.set_hydro(hydro)
        hp.set_photons(photons)
        st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=-1, switch=1)
        res[mode] = (st, hp.get_photons(), [hp.shard_stats(s) for s in range(4)])
        hp.close()
    st, ph, shards = res["persistent"]
    assert st["scatterings"] > 50
    for s in shards:
        assert abs(s["time_now"] - (frame["time_now"] + 1.0 / frame["fps"])) <= 1e-9 * s["time_now"], s
    assert res["streamed"][0]["scatterings"] == st["scatterings"] and res["streamed"][0]["iterations"] == st["iterations"]
    assert res["persistent_stream"][0]["scatterings"] == st["scatterings"] and res["persistent_stream"][0]["iterations"] == st["iterations"]
    for f in ph.dtype.names:
        assert np.array_equal(ph[f], res["streamed"][1][f], equal_nan=(ph.dtype[f].kind == "f")), f
        assert np.array_equal(ph[f], res["persistent_stream"][1][f], equal_nan=(ph.dtype[f].kind == "f")), f


def test_refused_cooperative_launch_falls_back_to_the_streamed_loop(monkeypatch):
    """If the device cannot hold the persistent grid (MPS share, no cooperative launch) the frame still runs --
    streamed, same photons -- and the context stays on the streamed loop."""
    cfg, hydro, photons, frame = synth.workload("C2", scale=1.0 / 16, n_photons=1200, seed=47)
    out = []
    for refuse in (False, True):
        if refuse:
            monkeypatch.setenv("MCRAT_B200_REFUSE_COOPERATIVE", "1")
        hp = HotPath(cfg, seed=3, num_shards=3, loop_mode="persistent")
        hp.set_hydro(hydro)
        hp.set_photons(photons)
        st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=90, switch=1)
        out.append((st, hp.get_photons(), hp.launch_count()))
        hp.close()
    (sa, pa, la), (sb, pb, lb) = out
    assert sa["iterations"] == sb["iterations"] == 90 and sa["scatterings"] == sb["scatterings"]
    assert lb > la + 200  # four launches per iteration instead of one per frame
    for f in pa.dtype.names:
        assert np.array_equal(pa[f], pb[f], equal_nan=(pa.dtype[f].kind == "f")), f


def test_division_by_c_is_exact():
    """The free path's division by C_LIGHT (Src/mclib.c:684) runs as an FMA-corrected multiplication by the reciprocal;
    it has to return the correctly rounded quotient, bit for bit, for every input (4 x 5e7 doubles here, all binades the
    free path can reach plus both ends of the double range)."""
    cfg, hydro, photons, frame = synth.workload("C1", scale=0.25, n_photons=64)
    hp = HotPath(cfg, seed=3)
    assert hp.selftest_div_by_c(50_000_000, seed=11) == 0


SKIP_CASES = [
    # workload, grid scale, photons, sub-shards, iterations, dilution, loop
    ("C1", 1.0 / 8, 3000, 2, 400, 1.0, "persistent"),     # 2-D Cartesian
    ("C2", 1.0 / 8, 4000, 3, 400, 1.0, "streamed"),       # 2-D cylindrical
    ("C2", 1.0 / 4, 3000, 2, 60, 5e-6, "persistent"),     # optically thin: long steps, photons cross cells all the time
    ("C5", 1.0 / 8, 4000, 4, 300, 1.0, "persistent"),     # 3-D spherical
    ("C5", 1.0 / 8, 4000, 4, 60, 1e-5, "streamed"),
    ("G2S", 1.0 / 8, 3000, 2, 300, 1.0, "persistent"),    # 2-D spherical
    ("G3C", 1.0 / 8, 3000, 2, 300, 1.0, "streamed"),      # 3-D Cartesian
    ("G3P", 1.0 / 8, 3000, 2, 300, 1.0, "persistent"),    # 3-D polar
    ("G3P", 1.0 / 8, 3000, 2, 60, 1e-5, "persistent"),
]


@pytest.mark.parametrize("wl,scale,nph,shards,iters,dilute,loop", SKIP_CASES)
def test_skipped_cell_rechecks_change_nothing(wl, scale, nph, shards, iters, dilute, loop):
    """The pass skips the containment re-check of a photon that provably cannot have left its cell
    (mcrat_b200_set_recheck_skip).  Mode 2 makes every skipped re-check anyway and fails the frame if one would not
    have succeeded; modes 0 (the reference's behaviour: re-check everything) and 1 must give the same photons bit for
    bit.  Several frames' worth of calls, so thresholds survive run_frame boundaries and a new hydro frame."""
    cfg, hydro, photons, frame = synth.workload(wl, scale=scale, n_photons=nph, seed=77)
    if dilute != 1.0:
        hydro = _thin(hydro, dilute)
    out = {}
    for mode in (0, 1, 2):
        hp = HotPath(cfg, seed=808, num_shards=shards, loop_mode=loop)
        hp.set_recheck_skip(mode)
        hp.set_hydro(hydro)
        hp.set_photons(photons)
        t, stats = frame["time_now"], []
        for call, sw in enumerate((1, 0, 0, 1, 0)):
            if sw == 1 and call > 0:
                hp.set_hydro(hydro)  # "new" hydro frame: thresholds of the old cells must not survive
            st = hp.run_frame(t, 1.0 / frame["fps"], max_iters=iters, switch=sw)  # raises on MCRAT_B200_ERR_STATE
            t = st["time_now"]
            stats.append((st["iterations"], st["scatterings"], st["relocations"]))
        out[mode] = (hp.get_photons(), stats)
    for mode in (1, 2):
        assert out[mode][1] == out[0][1]
        for f in out[0][0].dtype.names:
            assert np.array_equal(out[0][0][f], out[mode][0][f], equal_nan=out[0][0].dtype[f].kind == "f"), (mode, f)
    if dilute != 1.0:
        assert sum(s[2] for s in out[0][1]) > 100  # photons did change cells


def test_warp_wide_maxwell_juttner_sampling_equals_the_sequential_loop(monkeypatch):
    """Hot electrons (T >= 1e7 K) are drawn by a rejection loop (Src/electron.c:207-226) that the event evaluates 64
    trials at a time across a warp.  With MCRAT_B200_MJ_ROUNDS = 0 every electron is drawn by the sequential loop, with 1
    the sequential loop takes over after 64 rejected trials (which happens for most cold-ish cells): all three must give
    the same photons bit for bit, and the default build must match the oracle's sequential loop."""
    cfg, hydro, photons, frame = synth.workload("C3", scale=1.0 / 8, n_photons=4000, seed=13)
    out = {}
    for rounds in ("0", "1", None):
        if rounds is None:
            monkeypatch.delenv("MCRAT_B200_MJ_ROUNDS", raising=False)
        else:
            monkeypatch.setenv("MCRAT_B200_MJ_ROUNDS", rounds)
        hp = HotPath(cfg, seed=4242, num_shards=4)
        hp.set_hydro(hydro)
        hp.build_thermal_table(calls=20000, seed=3)
        hp.set_photons(photons)
        st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=150, switch=1)
        out[rounds] = (hp.get_photons(), st["scatterings"])
    assert out["0"][1] == out["1"][1] == out[None][1] > 300
    for r in ("0", "1"):
        for f in out[None][0].dtype.names:
            assert np.array_equal(out[None][0][f], out[r][0][f], equal_nan=out[None][0].dtype[f].kind == "f"), (r, f)


@pytest.mark.parametrize("wl,scale,nph,shards,iters", [("C2", 1.0 / 16, 6000, 6, 200), ("C5", 1.0 / 8, 8000, 16, 120),
                                                     ("C3", 1.0 / 8, 6000, 2, 120)])
def test_cluster_team_kernel_equals_the_cooperative_one(monkeypatch, wl, scale, nph, shards, iters):
    """The team of a sub-shard as one thread-block cluster (minima and state through distributed shared memory, remote
    mbarrier arrivals; MCRAT_B200_CLUSTER_TEAM=1) against the default cooperative team kernel: same photons, same counters."""
    cfg, hydro, photons, frame = synth.workload(wl, scale=scale, n_photons=nph, seed=61)
    out = {}
    for cluster in (False, True):
        if cluster:
            monkeypatch.setenv("MCRAT_B200_CLUSTER_TEAM", "1")
        hp = HotPath(cfg, seed=8, shard=1, num_shards=shards, loop_mode="persistent")
        hp.set_hydro(hydro)
        if wl == "C3":
            hp.build_thermal_table(calls=20000, seed=3)
        hp.set_photons(photons)
        st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=1)
        st2 = hp.run_frame(st["time_now"], 1.0 / frame["fps"], max_iters=iters // 2, switch=0)
        out[cluster] = (st, st2, hp.get_photons())
        hp.close()
        monkeypatch.delenv("MCRAT_B200_CLUSTER_TEAM", raising=False)
    for k in ("iterations", "scatterings", "relocations", "photon_slots", "time_now"):
        assert out[False][0][k] == out[True][0][k] and out[False][1][k] == out[True][1][k], k
    for f in out[False][2].dtype.names:
        assert np.array_equal(out[False][2][f], out[True][2][f], equal_nan=(out[False][2].dtype[f].kind == "f")), f


def test_persistent_stream_pair_is_called_off_when_its_grids_cannot_meet(monkeypatch):
    """The loop of lists larger than L2 runs as two co-resident grids.  Where they cannot be co-resident (ncu serialises
    kernels; a device shared with another tenant) the start-up handshake calls the launch off before a photon is touched
    and the frame runs through the streamed loop: same photons, no error."""
    cfg, hydro, photons, frame = synth.workload("C5", scale=1.0 / 8, n_photons=4000, seed=41)
    out = {}
    for mode, serialise in (("streamed", False), ("persistent_stream", False), ("persistent_stream", True)):
        if serialise:
            monkeypatch.setenv("MCRAT_B200_STREAM_SERIALIZE", "1")
        hp = HotPath(cfg, seed=31, shard=3, num_shards=8, loop_mode=mode)
        hp.set_hydro(hydro)
        hp.set_photons(photons)
        st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=80, switch=1)
        st2 = hp.run_frame(st["time_now"], 1.0 / frame["fps"], max_iters=40, switch=0)  # stays on the streamed loop
        out[(mode, serialise)] = (st, st2, hp.get_photons(), hp.launch_count())
        hp.close()
        monkeypatch.delenv("MCRAT_B200_STREAM_SERIALIZE", raising=False)
    a = out[("streamed", False)]
    for key in (("persistent_stream", False), ("persistent_stream", True)):
        b = out[key]
        assert a[0]["scatterings"] == b[0]["scatterings"] and a[1]["scatterings"] == b[1]["scatterings"], key
        for f in a[2].dtype.names:
            assert np.array_equal(a[2][f], b[2][f], equal_nan=(a[2].dtype[f].kind == "f")), (key, f)
    assert out[("persistent_stream", False)][3] < out[("persistent_stream", True)][3]  # the called-off run paid per-iteration launches


@pytest.mark.parametrize("nph,shards", [(6000, 2), (9000, 1), (40000, 80)])
def test_klein_nishina_rejections_streamed_equals_persistent(nph, shards):
    """C3 (x = h nu / m c^2 up to 10): a third of the candidates is rejected by the Klein-Nishina test and the event
    moves on to the next entry of the time order (Src/mclib.c:1128-1339).  The streamed loop finds it from the pass
    blocks' minima plus the slices of the blocks whose minimum is used up, the persistent loop reads the shard's times:
    same photons bit for bit, and fewer scatterings than shard-iterations would give without rejections' extra pushes."""
    cfg, hydro, photons, frame = synth.workload("C3", scale=1.0 / 8, n_photons=nph, seed=19)
    out = {}
    for mode in ("streamed", "persistent", "persistent_stream", "streamed_global"):
        hp = HotPath(cfg, seed=777, num_shards=shards, loop_mode=mode)
        hp.set_hydro(hydro)
        hp.build_thermal_table(calls=20000, seed=3)
        hp.set_photons(photons)
        st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=120, switch=1)
        st2 = hp.run_frame(st["time_now"], 1.0 / frame["fps"], max_iters=80, switch=0)
        out[mode] = (hp.get_photons(), st["scatterings"] + st2["scatterings"], st["iterations"] + st2["iterations"])
    a = out["streamed_global"]
    for mode in ("streamed", "persistent", "persistent_stream"):
        b = out[mode]
        assert a[1] == b[1] and a[2] == b[2] == 200
        for f in a[0].dtype.names:
            assert np.array_equal(a[0][f], b[0][f], equal_nan=a[0].dtype[f].kind == "f"), (mode, f)


@pytest.mark.parametrize("shards", [1, 3])
def test_walks_longer_than_the_push_list_match_the_oracle(shards):
    """photonEvent's walk over rejected candidates is unbounded in the reference (Src/mclib.c:1128-1339); the device
    records 16 pushes per event and, when the list is full, applies them in place and goes on.  Photons of x = h nu / m c^2
    ~ 50 are rejected 25 times in a row on average (sigma_KN / sigma_T ~ 0.04), so nearly every event overflows the list:
    streamed loop == persistent loop bit for bit, and shard 0 == the oracle."""
    cfg, hydro, photons, frame = synth.workload("C2", scale=1.0 / 16, n_photons=900 * shards, seed=23)
    ph = photons.copy()
    f = 50.0 * synth.M_EL * synth.C_LIGHT / ph["comv_p0"]
    for k in ("p0", "p1", "p2", "p3", "comv_p0", "comv_p1", "comv_p2", "comv_p3"):
        ph[k] *= f
    out = {}
    for mode in ("streamed", "persistent", "persistent_stream"):
        hp = HotPath(cfg, seed=616, shard=2, num_shards=shards, loop_mode=mode)
        hp.set_hydro(hydro)
        hp.set_photons(ph)
        st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=40, switch=1)
        out[mode] = (st, hp.get_photons(), hp.shard_stats(0))
        hp.close()
    (sa, pa, _), (sb, pb, ss) = out["streamed"], out["persistent"]
    assert sa["iterations"] == sb["iterations"] == 40 and sa["scatterings"] == sb["scatterings"]
    assert out["persistent_stream"][0]["scatterings"] == sa["scatterings"]
    for fld in pa.dtype.names:
        assert np.array_equal(pa[fld], pb[fld], equal_nan=pa.dtype[fld].kind == "f"), fld
        assert np.array_equal(pa[fld], out["persistent_stream"][1][fld], equal_nan=pa.dtype[fld].kind == "f"), fld
    sl = slice(ss["first_slot"], ss["first_slot"] + ss["num_slots"])
    o = api.Oracle(cfg)
    o.set_hydro(hydro)
    o.set_photons(ph[sl])
    rng = api.OracleRng("philox", seed=616, shard=2)
    ost = o.run_frame(rng, frame["time_now"], 1.0 / frame["fps"], max_iters=40, switch=1)
    assert ss["scatterings"] == ost["scatterings"] and ss["iterations"] == ost["iterations"]
    # every iteration draws (electron + 1) uniforms per candidate: far more than 17 candidates per event on average
    assert rng.ndraws > 40 * 17 * 4, rng.ndraws
    compare_photons(pb[sl], o.photons(), label="long walks", hydro=hydro)


def test_a_new_list_layout_never_replays_a_stream():
    """The Philox counters hold (slot, iteration): when set_photons lays the list out anew (injection, list growth) every
    sub-shard continues from the largest iteration number reached so far, so no (key, counter) pair is ever used twice.
    Checked against oracle ranks told to continue at that iteration."""
    cfg, hydro, photons, frame = synth.workload("C2", scale=1.0 / 16, n_photons=2400, seed=29)
    hp = HotPath(cfg, seed=31337, shard=40, num_shards=4)
    hp.set_hydro(hydro)
    hp.set_photons(photons[:2000])
    st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=25, switch=1)
    assert st["iterations"] == 25
    first = hp.get_photons()
    hp.set_photons(photons)  # 2400 photons: other slot ranges
    assert hp.num_shards() == 4
    st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=10, switch=1)
    got = hp.get_photons()
    for s in range(4):
        ss = hp.shard_stats(s)
        sl = slice(ss["first_slot"], ss["first_slot"] + ss["num_slots"])
        o = api.Oracle(cfg)
        o.set_hydro(hydro)
        o.set_photons(photons[sl])
        o.set_iter(25)
        rng = api.OracleRng("philox", seed=31337, shard=40 + s)
        ost = o.run_frame(rng, frame["time_now"], 1.0 / frame["fps"], max_iters=10, switch=1)
        assert ss["scatterings"] == ost["scatterings"], (s, ss, ost)
        compare_photons(got[sl], o.photons(), label="relayout shard %d" % s, hydro=hydro)
    # and the draws of the second frame are not those of the first one (shards >= 1 used to restart at iteration 0)
    assert not np.array_equal(first["time_to_scatter"][600:1100], got["time_to_scatter"][600:1100])
