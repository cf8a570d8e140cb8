"""Shared comparison helpers for the parity tests."""
import numpy as np

# Parity bars (BASELINE.json north_star): cell indices bit-exact; optical depth,
# time-to-scatter, 4-momenta and Stokes parameters within 1e-12 relative when both sides
# consume the same uniform stream.  "Relative" is taken against the natural scale of each
# quantity: |r| for positions, p0 for 4-momentum components, 1 for the normalised Stokes vector.
TOL = 1e-12


M_P, THOM_X_SECT = 1.6726231e-24, 6.65246e-25  # Src/mclib.c:5


def compare_photons(got, want, tol=TOL, stokes_tol=None, check_tts=True, label="", hydro=None):
    """Assert photon lists agree: integers exactly, floating point to `tol` relative.

    The optical depth tau' = n_lab sigma_T sigma_hat (1 - beta cos(theta)) (Src/optical_depth.c:46-58)
    cancels by up to 2 Gamma^2 for photons moving with the flow, which amplifies last-ulp
    differences between CUDA's and glibc's sin/cos/atan2 by the same factor.  tau' is therefore
    held to `tol` against its un-cancelled scale n_lab sigma_T (pass `hydro`), and
    time_to_scatter = -ln(xi)/(tau' c) to the matching relative bound tol * n_lab sigma_T / tau'.
    """
    stokes_tol = tol if stokes_tol is None else stokes_tol
    assert got.size == want.size, (label, got.size, want.size)
    assert np.array_equal(got["type"], want["type"]), label + ": photon types differ"
    assert np.array_equal(got["nearest_block_index"], want["nearest_block_index"]), \
        label + ": cell indices differ at %s" % np.nonzero(got["nearest_block_index"] != want["nearest_block_index"])[0][:10]
    assert np.array_equal(got["num_scatt"], want["num_scatt"]), label + ": num_scatt differs"
    assert np.array_equal(got["recalc_properties"], want["recalc_properties"]), label + ": recalc_properties differs"
    assert np.array_equal(got["weight"], want["weight"]), label + ": weights differ"
    errs = {}
    rnorm = np.sqrt(want["r0"] ** 2 + want["r1"] ** 2 + want["r2"] ** 2)
    for f in ("r0", "r1", "r2"):
        errs[f] = _rel(got[f], want[f], rnorm)
    live = want["nearest_block_index"] != -1
    # the lab momentum of a scattered photon is the boost of its comoving momentum back to the lab,
    # p0 = Gamma (p0' + beta.p') (Src/mclib.c:1262-1265): for photons scattered against the flow it
    # cancels by 2 Gamma^2, so it is held to `tol` against Gamma * p0' (pass `hydro`)
    p_scale = np.abs(want["p0"])
    if hydro is not None:
        gidx0 = np.where(live, want["nearest_block_index"], 0)
        p_scale = np.where(live, np.maximum(p_scale, np.asarray(hydro["gamma"])[gidx0] * np.abs(want["comv_p0"])), p_scale)
    for f in ("p0", "p1", "p2", "p3"):
        errs[f] = _rel(got[f], want[f], p_scale)
    # comoving momentum = Lorentz boost of the lab momentum, p0' = Gamma (p0 - beta.p)
    # (Src/mclib.c:302-407): for photons moving with the flow it cancels by the same 2 Gamma^2, so it
    # is held to `tol` against its un-cancelled scale Gamma * p0 (pass `hydro`)
    comv_scale = np.abs(want["comv_p0"])
    if hydro is not None:
        gidx = np.where(live, want["nearest_block_index"], 0)
        comv_scale = np.where(live, np.maximum(comv_scale, np.asarray(hydro["gamma"])[gidx] * np.abs(want["p0"])),
                              comv_scale)
    for f in ("comv_p0", "comv_p1", "comv_p2", "comv_p3"):
        errs[f] = _rel(got[f], want[f], comv_scale)
    tau = np.abs(want["total_optical_depth"])
    tau_scale = tau.copy()
    if hydro is not None:
        idx = np.where(live, want["nearest_block_index"], 0)
        tau_scale = np.where(live, np.maximum(tau, np.asarray(hydro["dens_lab"])[idx] / M_P * THOM_X_SECT), tau)
    errs["total_optical_depth"] = _rel(got["total_optical_depth"][live], want["total_optical_depth"][live],
                                       tau_scale[live])
    if check_tts:
        cond = np.where(live & (tau > 0), tau_scale / np.where(tau > 0, tau, 1.0), 1.0)
        errs["time_to_scatter"] = _rel(got["time_to_scatter"], want["time_to_scatter"],
                                       np.abs(want["time_to_scatter"]) * cond)
    serrs = {f: _rel(got[f], want[f], 1.0) for f in ("s0", "s1", "s2", "s3")}
    bad = {k: v for k, v in errs.items() if not v <= tol}
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
