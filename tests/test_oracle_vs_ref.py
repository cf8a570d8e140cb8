"""The oracle restatement (oracle/mcrat_oracle.c) pinned against the reference's own sources
(oracle/_ref, built from /root/reference/Src) and against the committed golden fixtures."""
import os

import numpy as np
import pytest

from mcrat_b200 import synth
from oracle import api, configs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FRAME_CASES = [
    ("C1", "c1_2d_cart", 1.0 / 16, 200, 150),
    ("C2", "c2_2d_cyl_stokes", 1.0 / 32, 200, 150),
    ("C3", "c3_2d_cyl_table", 1.0 / 32, 150, 100),
    ("C5", "c5_3d_sph", 1.0 / 8, 200, 150),
    # the remaining coordinate systems of Src/geometry.c
    ("G25", "g_25d_cyl", 1.0 / 8, 200, 150),
    ("G2S", "g_2d_sph", 1.0 / 8, 200, 150),
    ("G3C", "g_3d_cart", 1.0 / 4, 200, 150),
    ("G3P", "g_3d_polar", 1.0 / 4, 200, 150),
]


def _need_ref(name):
    if not api.ref_available(name):
        pytest.skip("oracle/_ref/%s not built (reference sources absent)" % name)


def _bitwise(a, b, skip=()):
    bad = []
    for f in api.PHOTON_DTYPE.names:
        if f in skip:
            continue
        if not np.array_equal(a[f], b[f], equal_nan=(f != "type")):
            bad.append(f)
    return bad


def _table():
    return np.load(os.path.join(GOLDEN, "thermal_table.npy"))


@pytest.mark.parametrize("wl,refname,scale,nph,iters", FRAME_CASES)
def test_frame_replay_bit_identical(wl, refname, scale, nph, iters):
    """Same inputs + the uniform stream the reference drew => bit-identical photon lists."""
    _need_ref(refname)
    cfg, hydro, photons, frame = synth.workload(wl, scale=scale, n_photons=nph, seed=21)
    ref = api.RefLib(refname)
    table = _table() if configs.CONFIGS[refname]["tau_calculation"] == configs.TABLE else None
    if table is not None:
        ref.set_table(table)
    ref.set_hydro(hydro)
    ref.set_photons(photons)
    rng, tee = ref.new_rng(seed=42, tee=4_000_000)
    st = ref.run_frame(rng, frame["time_now"], 1.0 / frame["fps"], max_iters=iters)
    u = ref.tee_values(rng, tee)
    o = api.Oracle(configs.CONFIGS[refname], table=table)
    o.set_hydro(hydro)
    o.set_photons(photons)
    orng = api.OracleRng("replay", buf=u)
    ost = o.run_frame(orng, frame["time_now"], 1.0 / frame["fps"], max_iters=iters)
    assert orng.replay_pos == u.size
    for k in ("iterations", "scatterings", "relocations", "time_now", "last_time_step"):
        assert st[k] == ost[k], (k, st, ost)
    assert _bitwise(ref.photons(), o.photons()) == []
    assert np.array_equal(ref.sorted_indexes()[:1], o.sorted_indexes()[:1])


@pytest.mark.parametrize("name", ["c1_2d_cart", "c2_2d_cyl_stokes", "c3_2d_cyl_table", "c5_3d_sph"])
def test_oracle_reproduces_golden(name):
    """Golden vectors were produced by the reference's sources (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=True)
    hydro = {k: g[k] for k in api.HYDRO_FIELDS}
    hydro.update(num_elements=int(g["num_elements"]), fps=float(g["fps"]), r0_domain=tuple(g["r0_domain"]),
                 r1_domain=tuple(g["r1_domain"]), r2_domain=tuple(g["r2_domain"]))
    cfg = configs.CONFIGS[name]
    o = api.Oracle(cfg, table=_table() if cfg["tau_calculation"] == configs.TABLE else None)
    o.set_hydro(hydro)
    o.set_photons(g["photons_in"])
    rng = api.OracleRng("replay", buf=g["uniforms"])
    st = o.run_frame(rng, float(g["time_now"]), float(g["dt"]), max_iters=int(g["iters"]))
    assert rng.replay_pos == g["uniforms"].size
    assert st["scatterings"] == dict(g["stats"])["scatterings"]
    assert _bitwise(o.photons(), g["photons_out"]) == []


GEOM_CONFIGS = ["c1_2d_cart", "c2_2d_cyl_stokes", "c5_3d_sph", "g_25d_cyl", "g_2d_sph", "g_3d_cart", "g_3d_polar"]


@pytest.mark.parametrize("refname", GEOM_CONFIGS)
def test_geometry_and_boost_units(refname):
    _need_ref(refname)
    ref = api.RefLib(refname)
    o = api.Oracle(configs.CONFIGS[refname])
    rs = np.random.default_rng(1)
    for _ in range(200):
        x, y, z = rs.normal(size=3) * 1e12
        assert np.array_equal(ref.coord_to_hydro(x, y, z), o.coord_to_hydro(x, y, z), equal_nan=True)
        v = rs.uniform(-0.5, 0.5, 3)
        pos = rs.uniform(0.1, 3.0, 3)
        assert np.array_equal(ref.hydro_vector_to_cartesian(*v, *pos), o.hydro_vector_to_cartesian(*v, *pos))
        b = rs.normal(size=3)
        b *= rs.uniform(0, 0.9999) / np.linalg.norm(b)
        n = rs.normal(size=3)
        n /= np.linalg.norm(n)
        e = 10 ** rs.uniform(-20, -16)
        p = np.array([e, *(e * n)])
        for obj in "pe":
            assert np.array_equal(ref.lorentz_boost(b, p, obj), o.lorentz_boost(b, p, obj))
    # zero boost: input renormalised in place, Src/mclib.c:388-391
    p = np.array([1.0, 0.6, 0.0, 0.7])
    assert np.array_equal(ref.lorentz_boost([0, 0, 0], p), o.lorentz_boost([0, 0, 0], p))
    # boosted photons stay null (zeroNorm, Src/mclib.c:409-434)
    out = o.lorentz_boost([0.3, -0.2, 0.9], [2.0, 2.0 / 3 ** 0.5, 2.0 / 3 ** 0.5, 2.0 / 3 ** 0.5])
    assert abs(out[0] - np.linalg.norm(out[1:])) <= 4e-16 * out[0]


@pytest.mark.parametrize("refname,temp", [("c2_2d_cyl_stokes", 3e5), ("c2_2d_cyl_stokes", 5e8), ("c1_2d_cart", 1e9)])
def test_single_scatter_units(refname, temp):
    """Electron sampling + polarised Klein-Nishina scatter, draw for draw."""
    _need_ref(refname)
    ref = api.RefLib(refname)
    o = api.Oracle(configs.CONFIGS[refname])
    rs = np.random.default_rng(2)
    for trial in range(60):
        n = rs.normal(size=3)
        n /= np.linalg.norm(n)
        e = 10 ** rs.uniform(-19, -16.5)
        ph = np.array([e, *(e * n)])
        q, uu = rs.uniform(-0.5, 0.5, 2)
        s = np.array([1.0, q, uu, 0.0]) if trial % 3 else np.array([1.0, 0.0, 0.0, 0.0])
        rng, tee = ref.new_rng(seed=1000 + trial, tee=100000)
        el = ref.single_thermal_electron(temp, ph, rng)
        ok, p_ref, s_ref = ref.single_scatter(el, ph, s, rng)
        u = ref.tee_values(rng, tee)
        orng = api.OracleRng("replay", buf=u)
        el_o = o.single_thermal_electron(temp, ph, orng)
        ok_o, p_o, s_o = o.single_scatter(el_o, ph, s, orng)
        assert orng.replay_pos == u.size
        assert ok == ok_o
        # the three Gaussian deviates of Src/electron.c:233 sit in one expression (unspecified order)
        assert np.allclose(el, el_o, rtol=4e-16, atol=0)
        assert np.allclose(p_ref, p_o, rtol=1e-13, atol=0) and np.allclose(s_ref, s_o, rtol=0, atol=1e-13)
        if ok:
            # scattered photon stays on the light cone; I == 1 after normalisation (:430-433)
            assert abs(p_o[0] - np.linalg.norm(p_o[1:])) <= 1e-15 * p_o[0]
            if configs.CONFIGS[refname]["stokes"]:
                assert s_o[0] == 1.0 and s_o[1] ** 2 + s_o[2] ** 2 <= 1 + 1e-12


def test_unpolarised_scatter_polarisation_degree():
    """Unpolarised in => Pi = sin^2 / (1 + cos^2 + (x0 - x)(1 - cos)) (Fano matrix, :411-416).

    The electron drifts at beta = 1e-7 in a generic direction: exactly at rest (or with the photon
    along z) the reference's findXY divides by a zero cross product (Src/mcrat_scattering.c:51)."""
    cfg = dict(configs.CONFIGS["c2_2d_cyl_stokes"])
    o = api.Oracle(cfg)
    rs = np.random.default_rng(4)
    me_c = 9.1093879e-28 * 2.99792458e10
    checked = 0
    for trial in range(60):
        x0 = 10 ** rs.uniform(-3, 0.5)
        n = rs.normal(size=3)
        n /= np.linalg.norm(n)
        ph = np.array([x0 * me_c, *(x0 * me_c * n)])
        b = rs.normal(size=3)
        b *= 1e-7 / np.linalg.norm(b)
        el = np.array([me_c, *(me_c * b)])
        rng = api.OracleRng("ranlxs0", seed=77 + trial)
        ok, p, s = o.single_scatter(el, ph, [1.0, 0.0, 0.0, 0.0], rng)
        if not ok:
            continue
        x1 = p[0] / me_c
        cth = float(np.dot(p[1:], n) / p[0])
        want = (1 - cth ** 2) / (1 + cth ** 2 + (x0 - x1) * (1 - cth))
        assert abs(np.hypot(s[1], s[2]) - want) < 1e-5
        assert abs(s[3]) < 1e-12 and s[0] == 1.0
        # Compton formula, :322
        assert abs(x1 - x0 / (1 + x0 * (1 - cth))) < 1e-5 * x0
        checked += 1
    assert checked > 30


def test_table_interpolation_matches_reference():
    _need_ref("c3_2d_cyl_table")
    ref = api.RefLib("c3_2d_cyl_table")
    tab = _table()
    ref.set_table(tab)
    o = api.Oracle(configs.CONFIGS["c3_2d_cyl_table"], table=tab)
    rng, _ = ref.new_rng(seed=1)
    orng = api.OracleRng("ranlxs0", seed=1)
    rs = np.random.default_rng(3)
    me_c = 9.1093879e-28 * 2.99792458e10
    for _ in range(300):
        e = 10 ** rs.uniform(-11.9, 5.9) * me_c
        temp = 10 ** rs.uniform(-3.9, 3.9) * 5.9298e9
        assert ref.thermal_cross_section(e, temp, rng) == o.thermal_cross_section(e, temp, orng)
    # below the table in temperature: closed-form early returns (Src/hot_x_section.c:336-339), no draws
    assert ref.thermal_cross_section(1e-3 * me_c, 1e4, rng) == o.thermal_cross_section(1e-3 * me_c, 1e4, orng)
    assert orng.ndraws == 0


def test_deterministic_table_agrees_with_reference_monte_carlo():
    """The quadrature table (mcrat_b200/hotxs.py) vs the reference's 5e5-sample MC integral."""
    _need_ref("c3_2d_cyl_table")
    ref = api.RefLib("c3_2d_cyl_table")
    from mcrat_b200 import hotxs
    rng, _ = ref.new_rng(seed=9)
    for x, theta in ((1e-3, 1e-2), (0.1, 0.3), (3.0, 2.0), (1e-5, 30.0)):
        mc = ref.L.ref_calculateTotalThermalCrossSection(api.C.c_double(x), api.C.c_double(theta), rng)
        assert abs(hotxs.hot_cross_section(x, theta) / mc - 1) < 2e-2, (x, theta, mc)


@pytest.mark.parametrize("refname", ["c4_3d_sph_cs", "c4b_3d_sph_cs_tote", "g_2d_cyl_cs"])
def test_cyclosynchrotron_absorb_and_emit(refname):
    _need_ref(refname)
    c = configs.CONFIGS[refname]
    if c["dimensions"] == configs.THREE:
        cfg, hydro, photons, frame = synth.workload("C4", scale=1.0 / 16, n_photons=300, seed=8)
    else:
        cfg, hydro, photons, frame = synth.workload("C2", scale=1.0 / 32, n_photons=300, seed=8)
        synth.toroidal_b_field(hydro, r_ref=2e12)
    hydro["scatt_frame_number"], hydro["inj_frame_number"] = 3, 2
    ref = api.RefLib(refname)
    o = api.Oracle(c)
    for eng in (ref, o):
        eng.set_hydro(hydro)
        eng.set_photons(photons)
    rng, tee = ref.new_rng(seed=3, tee=4_000_000)
    ref.find_containing_hydro_cell(1, rng)
    orng = api.OracleRng("ranlxs0", seed=3)
    o.find_containing_hydro_cell(1, orng)
    assert _bitwise(ref.photons(), o.photons()) == []
    # pool emission into the shell while the list is full (the driver's situation at the first
    # emission: addToPhotonList doubles the capacity, Src/photons.c:117-129).  Same stream => same
    # photons; the scratch fields of freshly malloc'ed records are not compared.
    r_inj = 1e12 if c["dimensions"] == configs.THREE else 2e12
    args = dict(r_inj=r_inj - 2.99792458e10 / 5, ph_weight=1e48, max_photons=2000, theta_min=0.0, theta_max=0.2)
    n_ref = ref.photon_emit_cyclosynch(rng, **args)
    n_o = o.photon_emit_cyclosynch(orng, **args)
    assert n_ref == n_o and n_ref > 0
    skip = ("time_to_scatter", "total_optical_depth")
    assert _bitwise(ref.photons(), o.photons(), skip=skip) == []
    assert ref.L.ref_list_num_null(ref.list) == o.list.num_null_photons
    # locate the pool photons, make a third of the list cold enough to be absorbed, then phAbsCyclosynch
    ref.find_containing_hydro_cell(0, rng)
    o.find_containing_hydro_cell(0, orng)
    ph = ref.photons()
    assert _bitwise(ph, o.photons(), skip=("time_to_scatter",)) == []
    ph = ph[ph["type"] != b"N"]  # setPhotonList assumes a list without null slots (Src/photons.c:93-106)
    ph["time_to_scatter"] = 0.0
    ph["comv_p0"][::3] *= 1e-12
    ph["type"][1::7] = b"k"
    for eng in (ref, o):
        eng.set_photons(ph)
    assert ref.ph_abs_cyclosynch() == o.ph_abs_cyclosynch()
    assert _bitwise(ref.photons(), o.photons()) == []


def _cs_list_for_rebin(wl, seed=77, n_photons=20000):
    """A list in the middle of a cyclo-synchrotron run: injected, comptonised, unabsorbed, pool and null photons, the
    rebinnable ones inside a narrow cone (total bins = 0.1 max_photons x n_theta x n_phi must stay <= max_photons)."""
    cfg, hydro, photons, frame = synth.workload(wl, scale=1.0 / 32, n_photons=n_photons, seed=seed)
    rng = np.random.default_rng(5)
    n = photons.size
    kinds = rng.choice(np.frombuffer(b"ikcpN", dtype="S1"), n, p=[0.15, 0.35, 0.2, 0.1, 0.2])
    photons["num_scatt"] = rng.integers(0, 40, n)
    photons["weight"] = 10 ** rng.uniform(48, 50, n)
    f = 10 ** rng.uniform(-2, 2, n)
    for k in ("p0", "p1", "p2", "p3"):
        photons[k] *= f
    ang = rng.uniform(0, 2 * np.pi, n)
    photons["s1"], photons["s2"] = 0.3 * np.cos(ang), 0.3 * np.sin(ang)
    rho, phi = np.hypot(photons["r0"], photons["r1"]), np.arctan2(photons["r1"], photons["r0"]) % (2 * np.pi)
    photons["r0"], photons["r1"] = rho * np.cos(phi / 20.0), rho * np.sin(phi / 20.0)
    rr = np.sqrt(photons["r0"] ** 2 + photons["r1"] ** 2 + photons["r2"] ** 2)
    th = np.degrees(np.arccos(photons["r2"] / rr))
    inside = th < th.min() + 1.7
    kinds = np.where(~inside & ((kinds == b"k") | (kinds == b"c")), b"i", kinds)
    photons["type"] = kinds
    null = kinds == b"N"
    for k in photons.dtype.names:
        if k != "type":
            photons[k][null] = 0
    photons["nearest_block_index"][null] = -1
    return hydro, photons


@pytest.mark.parametrize("refname,wl", [("c4_3d_sph_cs", "C4"), ("g_2d_cyl_cs", "C2")])
def test_cyclosynchrotron_rebin_bit_identical(refname, wl):
    """rebinCyclosynchCompPhotons (Src/mc_cyclosynch.c:600-710): the restatement against the reference's own function."""
    _need_ref(refname)
    hydro, photons = _cs_list_for_rebin(wl)
    ref = api.RefLib(refname)
    ref.set_hydro(hydro)
    ref.set_photons(photons)
    want_rc = ref.rebin_cyclosynch_comp_photons(3000)
    o = api.Oracle(configs.CONFIGS[refname])
    o.set_hydro(hydro)
    o.set_photons(photons)
    got_rc = o.rebin_cyclosynch_comp_photons(3000)
    assert want_rc[0] >= 0 and got_rc == want_rc
    assert (ref.photons()["type"] == b"k").sum() == want_rc[2] > 50
    assert _bitwise(ref.photons(), o.photons(), skip=("time_to_scatter", "total_optical_depth")) == []
    # no energy bins at all (0.1 x max_photons < 1): both refuse (:352-355)
    ref.set_photons(photons)
    o.set_photons(photons)
    assert ref.rebin_cyclosynch_comp_photons(5)[0] == -1 and o.rebin_cyclosynch_comp_photons(5)[0] == -1


@pytest.mark.parametrize("name", ["c2", "c5"])
def test_oracle_first_match_indices_at_full_grid_size(name):
    """A strided sample of the full-size index golden (the reference's own findContainingHydroCell at 1 048 576 cells)
    through the restatement: bit-exact, -1 included."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import index_golden_inputs
    g = np.load(os.path.join(GOLDEN, "index_full_%s.npz" % name))
    cfg, hydro, ph, frame, refname, geo = index_golden_inputs(name)
    if geo != str(g["geometry_sha256"]):
        pytest.skip("this machine's numpy builds the grid with other last bits than the golden's")
    pick = np.arange(0, ph.size, 400)
    o = api.Oracle(configs.CONFIGS[refname])
    o.set_hydro(hydro)
    o.set_photons(ph[pick])
    o.find_containing_hydro_cell(1, api.OracleRng("ranlxs0", seed=1))
    assert np.array_equal(o.photons()["nearest_block_index"], g["idx"][pick])
