"""End-to-end statistical consistency (BASELINE.json north_star, validation level 2).

Kernel-level parity is a replay of one uniform stream (test_gpu_parity.py).  Here the two sides
draw from *different* generators -- the reference's own sources on RANLXS0 streams, ours on keyed
Philox streams -- so only distributions can agree: binned spectrum, scattering counts, light
curve (detector arrival time) and polarisation degree per polar-angle bin.  Every statistic is a
two-sample test at the stated significance ALPHA; seeds are fixed, so the outcome is deterministic.

Both arms start from the same photon list, are cut into the same independent ranks of
PER_RANK photons (the reference's MPI decomposition, Src/mcrat.c:139-164) and stop every rank after
the same number of while-loop iterations, i.e. they sample the same stopped process.
"""
import numpy as np
import pytest
from scipy import stats

from mcrat_b200 import synth
from oracle import api

ALPHA = 1e-3          # per statistic; 5 statistics x 2 workloads => family-wise < 1.2e-2 under H0
C_LIGHT = 2.99792458e10

# workload, reference build, grid scale, photons, photons per rank, loop iterations per rank
CASES = [
    ("C2", "c2_2d_cyl_stokes", 1.0 / 16, 6000, 100, 1500),
    ("C5", "c5_3d_sph", 1.0 / 8, 4000, 100, 1500),
]


def _reference_arm(refname, cfg, hydro, photons, frame, per, iters, seed0=1):
    """The reference's own sources (oracle/_ref) rank by rank; the oracle port on ranlxs0 if the
    reference build did not travel (it does: oracle/_ref is not gpurun-ignored)."""
    use_ref = api.ref_available(refname)
    eng = api.RefLib(refname) if use_ref else api.Oracle(cfg)
    eng.set_hydro(hydro)
    out = photons.copy()
    scatt = 0
    for s in range(photons.size // per):
        sl = slice(s * per, (s + 1) * per)
        eng.set_photons(photons[sl])
        rng = eng.new_rng(seed=seed0 + s)[0] if use_ref else api.OracleRng("ranlxs0", seed=seed0 + s)
        st = eng.run_frame(rng, frame["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=1)
        out[sl] = eng.photons()
        scatt += st["scatterings"]
    return out, scatt


def _oracle_philox_arm(cfg, hydro, photons, frame, per, iters, seed, shard0):
    eng = api.Oracle(cfg)
    eng.set_hydro(hydro)
    out = photons.copy()
    scatt = 0
    for s in range(photons.size // per):
        sl = slice(s * per, (s + 1) * per)
        eng.set_photons(photons[sl])
        st = eng.run_frame(api.OracleRng("philox", seed=seed, shard=shard0 + s), frame["time_now"], 1.0 / frame["fps"],
                           max_iters=iters, switch=1)
        out[sl] = eng.photons()
        scatt += st["scatterings"]
    return out, scatt


def _observables(ph0, ph):
    """What ProcessMCRaT bins: energy, scattering count, detector arrival time, polarisation."""
    r = np.sqrt(ph["r0"] ** 2 + ph["r1"] ** 2 + ph["r2"] ** 2)
    pdir = np.stack([ph["p1"], ph["p2"], ph["p3"]]) / ph["p0"]
    theta_p = np.arccos(np.clip(pdir[2], -1, 1))
    # arrival-time offset of the photon at a distant detector along its own direction of flight:
    # t_det = t - r.n/c (common t drops out in a two-sample comparison)
    t_det = -(ph["r0"] * pdir[0] + ph["r1"] * pdir[1] + ph["r2"] * pdir[2]) / C_LIGHT
    return dict(
        log_e=np.log10(ph["p0"]),
        d_log_e=np.log10(ph["p0"] / ph0["p0"]),
        log_comv_e=np.log10(ph["comv_p0"]),
        num_scatt=ph["num_scatt"].astype(np.float64),
        t_det=t_det + r / C_LIGHT,   # lag behind the photon's own light cone: small, scattering-made
        theta_p=theta_p,
        q=ph["s1"] / ph["s0"], u=ph["s2"] / ph["s0"],
        pol=np.sqrt(ph["s1"] ** 2 + ph["s2"] ** 2 + ph["s3"] ** 2) / ph["s0"],
    )


def _chi2_binned(a, b, nbins=24):
    """Two-sample chi-square on a common binning (bins with < 10 counts merged into their neighbour)."""
    lo, hi = min(a.min(), b.min()), max(a.max(), b.max())
    edges = np.linspace(lo, hi, nbins + 1)
    ha, _ = np.histogram(a, edges)
    hb, _ = np.histogram(b, edges)
    A, B, ca, cb = [], [], 0, 0
    for x, y in zip(ha, hb):
        ca += x
        cb += y
        if ca + cb >= 20:
            A.append(ca)
            B.append(cb)
            ca = cb = 0
    if A:
        A[-1] += ca
        B[-1] += cb
    A, B = np.array(A, float), np.array(B, float)
    if A.size < 2:
        return 1.0
    k1, k2 = np.sqrt(B.sum() / A.sum()), np.sqrt(A.sum() / B.sum())
    chi2 = np.sum((k1 * A - k2 * B) ** 2 / (A + B))
    return float(stats.chi2.sf(chi2, A.size - 1))


def _assert_consistent(ph0, ours, ref, scatt_ours, scatt_ref, stokes, label):
    a, b = _observables(ph0, ours), _observables(ph0, ref)
    report = {}
    names = ["d_log_e", "log_comv_e", "num_scatt", "t_det"]
    for nm in names:
        report["ks:" + nm] = float(stats.ks_2samp(a[nm], b[nm]).pvalue)
    report["chi2:spectrum"] = _chi2_binned(a["log_e"], b["log_e"])
    if stokes:
        # polarisation degree and Stokes q, u of the scattered photons, overall and per polar-angle bin
        ma, mb = ours["num_scatt"] > 0, ref["num_scatt"] > 0
        for nm in ("pol", "q", "u"):
            report["ks:" + nm] = float(stats.ks_2samp(a[nm][ma], b[nm][mb]).pvalue)
        edges = np.quantile(np.concatenate([a["theta_p"][ma], b["theta_p"][mb]]), [0, 1 / 3, 2 / 3, 1])
        for k in range(3):
            sa = ma & (a["theta_p"] >= edges[k]) & (a["theta_p"] <= edges[k + 1])
            sb = mb & (b["theta_p"] >= edges[k]) & (b["theta_p"] <= edges[k + 1])
            report["ks:pol|theta_bin%d" % k] = float(stats.ks_2samp(a["pol"][sa], b["pol"][sb]).pvalue)
    # total scattering counts: Poisson-like totals of independent ranks; compare through the per-photon mean
    na, nb = a["num_scatt"], b["num_scatt"]
    z = (na.mean() - nb.mean()) / np.sqrt(na.var(ddof=1) / na.size + nb.var(ddof=1) / nb.size)
    report["z:mean_num_scatt"] = float(2 * stats.norm.sf(abs(z)))
    print(label, "scatterings ours/ref = %d/%d" % (scatt_ours, scatt_ref),
          {k: "%.3g" % v for k, v in report.items()})
    bad = {k: v for k, v in report.items() if not v > ALPHA}
    assert not bad, "%s: distributions differ at alpha=%g: %s" % (label, ALPHA, bad)
    assert scatt_ours > 0.2 * ph0.size and scatt_ref > 0.2 * ph0.size, "test too weak: almost nothing scattered"


@pytest.mark.parametrize("wl,refname,scale,nph,per,iters", CASES)
def test_oracle_on_philox_streams_is_statistically_the_reference(wl, refname, scale, nph, per, iters):
    """CPU: the restatement drawing from keyed Philox streams (what the GPU consumes) vs the
    reference's own sources on RANLXS0 -- the generator swap changes no distribution."""
    nph = nph // 3
    cfg, hydro, photons, frame = synth.workload(wl, scale=scale, n_photons=nph, seed=17)
    ref, sref = _reference_arm(refname, cfg, hydro, photons, frame, per, iters)
    ours, sours = _oracle_philox_arm(cfg, hydro, photons, frame, per, iters, seed=99, shard0=0)
    _assert_consistent(photons, ours, ref, sours, sref, cfg["stokes"], wl + " oracle/philox vs reference/ranlxs0")


@pytest.mark.gpu
@pytest.mark.parametrize("wl,refname,scale,nph,per,iters", CASES)
def test_gpu_spectra_lightcurve_polarisation_consistent_with_reference(wl, refname, scale, nph, per, iters):
    """GPU production path (Philox, sub-shards) vs the reference's own sources on RANLXS0 streams."""
    from mcrat_b200 import HotPath
    cfg, hydro, photons, frame = synth.workload(wl, scale=scale, n_photons=nph, seed=23)
    hp = HotPath(cfg, seed=4242, shard=0, num_shards=nph // per)
    hp.set_hydro(hydro)
    hp.set_photons(photons)
    assert hp.num_shards() == nph // per
    st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=1)
    ours = hp.get_photons()
    ref, sref = _reference_arm(refname, cfg, hydro, photons, frame, per, iters)
    _assert_consistent(photons, ours, ref, st["scatterings"], sref, cfg["stokes"], wl + " GPU/philox vs reference/ranlxs0")
