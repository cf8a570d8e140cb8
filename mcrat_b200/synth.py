"""Synthetic outflow grids and photon lists of the BASELINE.json shapes.

No hydro simulation files exist offline, so benchmark and test inputs are
built directly as the cell SoA the hot path scans (``struct hydro_dataframe``,
Src/mcrat.h:194-244) with fields from the same analytic outflow models the
reference ships (Src/analytic_outflows.c:70-236), and as ``struct photon``
records (Src/mcrat.h:142-171) placed in a shell of the flow.  numpy only.
"""
import numpy as np

# Src/mclib.c:4-5
A_RAD = 7.56e-15
C_LIGHT = 2.99792458e10
PL_CONST = 6.6260755e-27
K_B = 1.380658e-16
M_P = 1.6726231e-24
THOM_X_SECT = 6.65246e-25
M_EL = 9.1093879e-28

CARTESIAN, SPHERICAL, CYLINDRICAL, POLAR = 0, 1, 2, 3
TWO, TWO_POINT_FIVE, THREE = 0, 1, 2

PHOTON_DTYPE = np.dtype([
    ("type", "S1"),
    ("p0", "f8"), ("p1", "f8"), ("p2", "f8"), ("p3", "f8"),
    ("comv_p0", "f8"), ("comv_p1", "f8"), ("comv_p2", "f8"), ("comv_p3", "f8"),
    ("r0", "f8"), ("r1", "f8"), ("r2", "f8"),
    ("s0", "f8"), ("s1", "f8"), ("s2", "f8"), ("s3", "f8"),
    ("num_scatt", "f8"),
    ("recalc_properties", "i4"),
    ("weight", "f8"),
    ("nearest_block_index", "i4"),
    ("time_to_scatter", "f8"),
    ("total_optical_depth", "f8"),
], align=True)


def _edges_to_cells(edges):
    return 0.5 * (edges[1:] + edges[:-1]), (edges[1:] - edges[:-1])


def _block_major(nx, ny, bs=8):
    """Permutation of a (ny, nx) row-major cell list into FLASH block order
    (8x8-cell blocks, Src/mclib_flash.c:10-12): blocks row-major, cells row-major inside a block."""
    idx = np.arange(nx * ny).reshape(ny, nx)
    idx = idx.reshape(ny // bs, bs, nx // bs, bs).transpose(0, 2, 1, 3)
    return idx.reshape(-1)


def make_grid(dimensions, geometry, shape, extent, log_r=False, flash_blocks=False, fps=5.0):
    """Cell centres / sizes in hydro coordinates plus domain limits.

    shape: (n0, n1[, n2]); extent: ((lo0, hi0), (lo1, hi1)[, (lo2, hi2)]).
    Cell order: first coordinate fastest (PLUTO order, Src/mclib_pluto.c:1163-1169) unless
    ``flash_blocks`` (2-D only).
    """
    nd = 3 if dimensions == THREE else 2
    axes = []
    for d in range(nd):
        lo, hi = extent[d]
        if d == 0 and log_r:
            e = np.logspace(np.log10(lo), np.log10(hi), shape[d] + 1)
        else:
            e = np.linspace(lo, hi, shape[d] + 1)
        axes.append(_edges_to_cells(e))
    if nd == 2:
        c1, c0 = np.meshgrid(axes[1][0], axes[0][0], indexing="ij")
        s1, s0 = np.meshgrid(axes[1][1], axes[0][1], indexing="ij")
        r0, r1, z0, z1 = c0.ravel(), c1.ravel(), s0.ravel(), s1.ravel()
        if flash_blocks:
            p = _block_major(shape[0], shape[1])
            r0, r1, z0, z1 = r0[p], r1[p], z0[p], z1[p]
        n = r0.size
        r2 = np.zeros(n)
        z2 = np.zeros(n)
    else:
        c2, c1, c0 = np.meshgrid(axes[2][0], axes[1][0], axes[0][0], indexing="ij")
        s2, s1, s0 = np.meshgrid(axes[2][1], axes[1][1], axes[0][1], indexing="ij")
        r0, r1, r2 = c0.ravel(), c1.ravel(), c2.ravel()
        z0, z1, z2 = s0.ravel(), s1.ravel(), s2.ravel()
        n = r0.size
    h = dict(num_elements=n, r0=r0.copy(), r1=r1.copy(), r2=r2.copy(), r0_size=z0.copy(), r1_size=z1.copy(),
             r2_size=z2.copy(), fps=fps, dimensions=dimensions, geometry=geometry,
             r0_domain=tuple(extent[0]), r1_domain=tuple(extent[1]),
             r2_domain=tuple(extent[2]) if nd == 3 else (0.0, 0.0))
    h["_edges"] = [np.concatenate([a[0] - 0.5 * a[1], [a[0][-1] + 0.5 * a[1][-1]]]) for a in axes]
    h["_shape"] = tuple(shape[:nd])
    h["_inv_perm"] = None
    if nd == 2 and flash_blocks:
        inv = np.empty(n, dtype=np.int64)
        inv[p] = np.arange(n)
        h["_inv_perm"] = inv
    h["r"], h["theta"] = hydro_to_spherical(dimensions, geometry, h["r0"], h["r1"], h["r2"])
    for f in ("v0", "v1", "v2", "dens", "dens_lab", "pres", "temp", "gamma", "B0", "B1", "B2"):
        h[f] = np.zeros(n)
    return h


def hydro_to_spherical(dimensions, geometry, r0, r1, r2):
    """Src/geometry.c:66-106."""
    if dimensions != THREE:
        if geometry in (CARTESIAN, CYLINDRICAL):
            return np.sqrt(r0 * r0 + r1 * r1), np.arctan2(r0, r1)
        return r0.copy(), r1.copy()
    if geometry == CARTESIAN:
        r = np.sqrt(r0 * r0 + r1 * r1 + r2 * r2)
        return r, np.arccos(r2 / r)
    if geometry == SPHERICAL:
        return r0.copy(), r1.copy()
    r = np.sqrt(r0 * r0 + r2 * r2)
    return r, np.arccos(r2 / r)


def _radial_velocity(h, vel):
    """Radial flow of speed ``vel`` expressed in the grid's own unit vectors
    (Src/analytic_outflows.c:99-140)."""
    dims, g = h["dimensions"], h["geometry"]
    if dims != THREE:
        if g in (CARTESIAN, CYLINDRICAL):
            r = np.sqrt(h["r0"] ** 2 + h["r1"] ** 2)
            h["v0"] = vel * h["r0"] / r
            h["v1"] = vel * h["r1"] / r
        else:
            h["v0"] = vel.copy()
            h["v1"] = np.zeros_like(vel)
        h["v2"] = np.zeros_like(vel)
    else:
        if g == CARTESIAN:
            r = np.sqrt(h["r0"] ** 2 + h["r1"] ** 2 + h["r2"] ** 2)
            h["v0"], h["v1"], h["v2"] = vel * h["r0"] / r, vel * h["r1"] / r, vel * h["r2"] / r
        elif g == SPHERICAL:
            h["v0"], h["v1"], h["v2"] = vel.copy(), np.zeros_like(vel), np.zeros_like(vel)
        else:
            r = np.sqrt(h["r0"] ** 2 + h["r2"] ** 2)
            h["v0"], h["v1"], h["v2"] = vel * h["r0"] / r, np.zeros_like(vel), vel * h["r2"] / r


def spherical_outflow(h, gamma_infinity=100.0, lumi=1e54, r00=1e8):
    """Adiabatic fireball, Src/analytic_outflows.c:70-145."""
    r = h["r"]
    coast = r >= r00 * gamma_infinity
    gamma = np.where(coast, gamma_infinity, r / r00)
    pres = np.where(coast,
                    lumi * r00 ** (2.0 / 3.0) * r ** (-8.0 / 3.0) / (12.0 * np.pi * C_LIGHT * gamma_infinity ** (4.0 / 3.0)),
                    lumi * r00 ** 2 / (12.0 * np.pi * C_LIGHT * r ** 4))
    h["gamma"] = gamma
    h["pres"] = pres
    h["dens"] = lumi / (4 * np.pi * r ** 2 * C_LIGHT ** 3 * gamma_infinity * gamma)
    h["dens_lab"] = h["dens"] * gamma
    h["temp"] = (3 * pres / A_RAD) ** 0.25
    _radial_velocity(h, np.sqrt(1 - gamma ** -2.0))
    return h


def structured_jet(h, gamma_0=100.0, lumi=1e52, r00=1e8, theta_j=1e-2, p=4.0):
    """Lundman, Pe'er & Ryde (2014) structured jet, Src/analytic_outflows.c:147-236."""
    r, th = h["r"], h["theta"]
    T_0 = (lumi / (4 * np.pi * r00 * r00 * A_RAD * C_LIGHT)) ** 0.25
    eta = gamma_0 / np.sqrt(1 + (th / theta_j) ** (2 * p))
    eta = np.where(th >= theta_j * (gamma_0 / 2) ** (1.0 / p), 2.0, eta)
    r_sat = eta * r00
    coast = r >= r_sat
    gamma = np.where(coast, eta, r / r_sat)
    gamma = np.maximum(gamma, 1.0 + 1e-9)
    temp = np.where(coast, T_0 * (r_sat / r) ** (2.0 / 3.0) / eta, T_0)
    vel = np.sqrt(1 - gamma ** -2.0)
    h["gamma"] = gamma
    h["temp"] = temp
    h["dens"] = M_P * lumi / (4 * np.pi * M_P * C_LIGHT ** 3 * eta * vel * gamma * r * r)
    h["dens_lab"] = h["dens"] * gamma
    h["pres"] = A_RAD * temp ** 4 / 3
    _radial_velocity(h, vel)
    return h


def toroidal_b_field(h, sigma=0.1, r_ref=1e11):
    """Toroidal field with B^2/8pi = sigma * rho c^2 at r_ref, falling as 1/r."""
    i = np.argmin(np.abs(h["r"] - r_ref))
    b_ref = np.sqrt(8 * np.pi * sigma * h["dens"][i] * C_LIGHT ** 2)
    b = b_ref * r_ref / h["r"]
    if h["dimensions"] == THREE and h["geometry"] == SPHERICAL:
        h["B0"], h["B1"], h["B2"] = np.zeros_like(b), np.zeros_like(b), b
    else:
        h["B0"], h["B1"], h["B2"] = np.zeros_like(b), np.zeros_like(b), b
    return h


def hydro_vector_to_cartesian(dimensions, geometry, v0, v1, v2, x0, x1, x2):
    """Src/geometry.c:189-253 (vectorised)."""
    if dimensions == TWO:
        if geometry in (CARTESIAN, CYLINDRICAL):
            return v0 * np.cos(x2), v0 * np.sin(x2), v1
        v2 = 0 * v2
    if dimensions == TWO_POINT_FIVE and geometry in (CARTESIAN, CYLINDRICAL):
        return v0 * np.cos(x2) - v2 * np.sin(x2), v0 * np.sin(x2) + v2 * np.cos(x2), v1
    if dimensions == THREE and geometry == CARTESIAN:
        return v0, v1, v2
    if dimensions == THREE and geometry == POLAR:
        return v0 * np.cos(x1) - v1 * np.sin(x1), v0 * np.sin(x1) + v1 * np.cos(x1), v2
    return (v0 * np.sin(x1) * np.cos(x2) + v1 * np.cos(x1) * np.cos(x2) - v2 * np.sin(x2),
            v0 * np.sin(x1) * np.sin(x2) + v1 * np.cos(x1) * np.sin(x2) + v2 * np.cos(x2),
            v0 * np.cos(x1) - v1 * np.sin(x1))


def mcrat_to_hydro(dimensions, geometry, x, y, z):
    """Src/geometry.c:15-64 (vectorised)."""
    if dimensions != THREE:
        if geometry in (CARTESIAN, CYLINDRICAL):
            return np.sqrt(x * x + y * y), z, -np.ones_like(x)
        r = np.sqrt(x * x + y * y + z * z)
        return r, np.arccos(z / r), -np.ones_like(x)
    if geometry == CARTESIAN:
        return x, y, z
    phi = np.fmod(np.arctan2(y, x) * 180.0 / np.pi + 360.0, 360.0) * np.pi / 180
    if geometry == SPHERICAL:
        r = np.sqrt(x * x + y * y + z * z)
        return r, np.arccos(z / r), phi
    return np.sqrt(x * x + y * y), phi, z


def _boost(beta, p):
    """Lorentz boost of 4-vectors p (n,4) by velocities beta (n,3): frame moving with +beta."""
    b2 = np.sum(beta * beta, axis=1)
    g = 1.0 / np.sqrt(1 - b2)
    bp = np.sum(beta * p[:, 1:], axis=1)
    out = np.empty_like(p)
    out[:, 0] = g * (p[:, 0] - bp)
    fac = np.where(b2 > 0, (g - 1) * bp / np.where(b2 > 0, b2, 1.0), 0.0) - g * p[:, 0]
    out[:, 1:] = p[:, 1:] + fac[:, None] * beta
    n = np.sqrt(np.sum(out[:, 1:] ** 2, axis=1))
    out[:, 1:] *= (out[:, 0] / n)[:, None]
    return out


def locate_cells(h, hc0, hc1, hc2):
    """Containing cell on the structured grids built by make_grid (for building inputs only)."""
    if "_edges" not in h:
        return locate_cells_bruteforce(h, hc0, hc1, hc2)
    e = h["_edges"]
    shape = h["_shape"]
    coords = (hc0, hc1, hc2)[:len(shape)]
    idx = []
    ok = np.ones(hc0.size, dtype=bool)
    for d, c in enumerate(coords):
        i = np.searchsorted(e[d], c, side="right") - 1
        ok &= (i >= 0) & (i < shape[d]) & (c <= e[d][-1])
        idx.append(np.clip(i, 0, shape[d] - 1))
    flat = idx[0] + shape[0] * idx[1] if len(shape) == 2 else idx[0] + shape[0] * (idx[1] + shape[1] * idx[2])
    if h.get("_inv_perm") is not None:
        flat = h["_inv_perm"][flat]
    return np.where(ok, flat, -1)


def locate_cells_bruteforce(h, hc0, hc1, hc2, chunk=256):
    """First-match containing cell (numpy; for building inputs only)."""
    nd3 = h["dimensions"] == THREE
    out = np.full(hc0.size, -1, dtype=np.int64)
    for s in range(0, hc0.size, chunk):
        e = min(s + chunk, hc0.size)
        m = (2 * np.abs(hc0[s:e, None] - h["r0"][None, :]) - h["r0_size"][None, :] <= 0)
        m &= (2 * np.abs(hc1[s:e, None] - h["r1"][None, :]) - h["r1_size"][None, :] <= 0)
        if nd3:
            m &= (2 * np.abs(hc2[s:e, None] - h["r2"][None, :]) - h["r2_size"][None, :] <= 0)
        any_ = m.any(axis=1)
        out[s:e] = np.where(any_, m.argmax(axis=1), -1)
    return out


def make_photons(h, n, r_range, theta_range, seed=1234, weight=1e50, phi_range=(0.0, 2 * np.pi)):
    """Black-body photons, isotropic in the local fluid frame, placed uniformly in a shell.

    Plays the role of ``photonInjection`` (Src/mclib.c:9-300, out of scope for the GPU
    path) for synthetic inputs: each photon gets the comoving black-body energy of
    the cell it sits in and is boosted to the lab frame with that cell's velocity.
    """
    rng = np.random.default_rng(seed)
    dims, g = h["dimensions"], h["geometry"]
    u = rng.random(n)
    r = (r_range[0] ** 3 + u * (r_range[1] ** 3 - r_range[0] ** 3)) ** (1.0 / 3.0)
    cmin, cmax = np.cos(theta_range[0]), np.cos(theta_range[1])
    th = np.arccos(cmin + rng.random(n) * (cmax - cmin))
    ph = phi_range[0] + rng.random(n) * (phi_range[1] - phi_range[0])
    x, y, z = r * np.sin(th) * np.cos(ph), r * np.sin(th) * np.sin(ph), r * np.cos(th)
    hc0, hc1, hc2 = mcrat_to_hydro(dims, g, x, y, z)
    cell = locate_cells(h, hc0, hc1, hc2)
    ok = cell >= 0
    c = np.where(ok, cell, 0)
    temp = h["temp"][c]
    # Planck photon-number spectrum (Bjorkman & Wood 2001 series sampling)
    zeta = np.cumsum(1.0 / np.arange(1, 200) ** 3)
    m = 1 + np.searchsorted(zeta / zeta[-1], rng.random(n))
    xx = -np.log(rng.random(n) * rng.random(n) * rng.random(n)) / m
    e_comv = xx * K_B * temp / C_LIGHT  # p0 = E/c
    mu = 2 * rng.random(n) - 1
    az = 2 * np.pi * rng.random(n)
    st = np.sqrt(1 - mu * mu)
    p_comv = np.stack([e_comv, e_comv * st * np.cos(az), e_comv * st * np.sin(az), e_comv * mu], axis=1)
    phi_pos = np.arctan2(y, x)
    x2 = h["r2"][c] if dims == THREE else phi_pos
    bx, by, bz = hydro_vector_to_cartesian(dims, g, h["v0"][c], h["v1"][c], h["v2"][c], h["r0"][c], h["r1"][c], x2)
    beta = np.stack([bx, by, bz], axis=1)
    p_lab = _boost(-beta, p_comv)
    out = np.zeros(n, dtype=PHOTON_DTYPE)
    out["type"] = b"i"
    out["p0"], out["p1"], out["p2"], out["p3"] = p_lab.T
    out["comv_p0"], out["comv_p1"], out["comv_p2"], out["comv_p3"] = p_comv.T
    out["r0"], out["r1"], out["r2"] = x, y, z
    out["s0"] = 1.0
    out["weight"] = weight
    out["nearest_block_index"] = 0
    out["recalc_properties"] = 1
    return out


def locate_photons(h, photons):
    """Containing cell of every photon on the structured grids of make_grid (-1 outside the domain).  Input
    preparation only: lets a bounded CPU sample start from located photons without a full photon x cell scan."""
    hc0, hc1, hc2 = mcrat_to_hydro(h["dimensions"], h["geometry"], photons["r0"], photons["r1"], photons["r2"])
    return locate_cells(h, hc0, hc1, hc2).astype(np.int32)


# ---------------------------------------------------------------------------------------
# named workloads (BASELINE.json configs)
# ---------------------------------------------------------------------------------------
def workload(name, scale=1.0, n_photons=None, seed=1234):
    """Return (config dict, hydro dict, photons, frame dict) for a named BASELINE config.

    ``scale`` < 1 shrinks the grid (cells per axis) for tests; photon counts are given
    explicitly by ``n_photons`` or default to the BASELINE value.
    """
    if name == "C1":
        cfg = dict(dimensions=TWO, geometry=CARTESIAN, stokes=0, tau_calculation=1, cyclosynch=0,
                   b_field_calc=1, epsilon_b=0.5)
        nx, ny = max(8, int(256 * scale) // 8 * 8), max(8, int(1280 * scale) // 8 * 8)
        h = make_grid(TWO, CARTESIAN, (nx, ny), ((0.0, 5e12), (0.0, 2.5e13)), flash_blocks=True)
        spherical_outflow(h)
        n = n_photons or 10000
        ph = make_photons(h, n, (1e12 - 0.5 * C_LIGHT / 5, 1e12 + 0.5 * C_LIGHT / 5), (1e-4, np.deg2rad(2.0)), seed)
        frame = dict(fps=5.0, time_now=1e12 / C_LIGHT)
    elif name in ("C2", "C3"):
        cfg = dict(dimensions=TWO, geometry=CYLINDRICAL, stokes=1, tau_calculation=2 if name == "C3" else 1,
                   cyclosynch=0, b_field_calc=1, epsilon_b=0.5)
        nx = ny = max(8, int(1024 * scale) // 8 * 8)
        h = make_grid(TWO, CYLINDRICAL, (nx, ny), ((0.0, 2.5e11), (1.0e12, 3.0e12)), flash_blocks=True)
        structured_jet(h)
        if name == "C3":
            # hot-electron regime: theta = kT/m_e c^2 log-uniform in [1e-3, 1]
            rng = np.random.default_rng(seed + 7)
            theta = 10 ** rng.uniform(-3, 0, h["num_elements"])
            h["temp"] = theta * M_EL * C_LIGHT ** 2 / K_B
        n = n_photons or 100000
        ph = make_photons(h, n, (2e12 - 0.5 * C_LIGHT / 5, 2e12 + 0.5 * C_LIGHT / 5), (1e-5, np.deg2rad(6.0)), seed)
        if name == "C3":
            rng = np.random.default_rng(seed + 11)
            f = 10 ** rng.uniform(-6, 1, n) * M_EL * C_LIGHT / ph["p0"]
            for k in ("p0", "p1", "p2", "p3", "comv_p0", "comv_p1", "comv_p2", "comv_p3"):
                ph[k] *= f
        frame = dict(fps=5.0, time_now=2e12 / C_LIGHT)
    elif name in ("C4", "C5"):
        cfg = dict(dimensions=THREE, geometry=SPHERICAL, stokes=1, tau_calculation=1,
                   cyclosynch=1 if name == "C4" else 0, b_field_calc=2, epsilon_b=0.5)
        nr, nt, npp = (max(4, int(256 * scale)), max(4, int(64 * scale)), max(4, int(64 * scale)))
        h = make_grid(THREE, SPHERICAL, (nr, nt, npp), ((1e11, 1e13), (0.0, np.pi / 8), (0.0, 2 * np.pi)), log_r=True)
        structured_jet(h, theta_j=0.1)
        toroidal_b_field(h)
        n = n_photons or 100000
        ph = make_photons(h, n, (1e12 - 0.5 * C_LIGHT / 5, 1e12 + 0.5 * C_LIGHT / 5), (1e-3, np.deg2rad(10.0)), seed)
        frame = dict(fps=5.0, time_now=1e12 / C_LIGHT)
    elif name in ("G25", "G2S", "G3C", "G3P"):
        # geometry coverage (Src/geometry.c:15-253): the other coordinate systems the reference supports
        stokes = 0 if name == "G3C" else 1
        if name == "G25":    # 2.5-D cylindrical (PLUTO): (R, z) grid with an azimuthal velocity component
            dims, geom = TWO_POINT_FIVE, CYLINDRICAL
            n0 = n1 = max(8, int(512 * scale))
            h = make_grid(dims, geom, (n0, n1), ((0.0, 2.5e11), (1.0e12, 3.0e12)))
        elif name == "G2S":  # 2-D spherical (PLUTO-CHOMBO): (r, theta)
            dims, geom = TWO, SPHERICAL
            n0, n1 = max(8, int(512 * scale)), max(8, int(256 * scale))
            h = make_grid(dims, geom, (n0, n1), ((1.0e12, 3.0e12), (0.0, np.pi / 16)))
        elif name == "G3C":  # 3-D Cartesian (PLUTO): (x, y, z) box around the axis
            dims, geom = THREE, CARTESIAN
            n0 = n1 = max(4, int(64 * scale))
            n2 = max(4, int(256 * scale))
            h = make_grid(dims, geom, (n0, n1, n2), ((-2.5e11, 2.5e11), (-2.5e11, 2.5e11), (1.0e12, 3.0e12)))
        else:                # 3-D polar (PLUTO): (R, phi, z)
            dims, geom = THREE, POLAR
            n0, n1, n2 = max(4, int(128 * scale)), max(4, int(32 * scale)), max(4, int(256 * scale))
            h = make_grid(dims, geom, (n0, n1, n2), ((0.0, 2.5e11), (0.0, 2 * np.pi), (1.0e12, 3.0e12)))
        structured_jet(h, theta_j=0.05)
        if name == "G25":  # part of the speed goes into rotation about the axis
            vel = np.sqrt(h["v0"] ** 2 + h["v1"] ** 2)
            h["v2"] = 0.02 * vel
            h["v0"] *= np.sqrt(1 - 0.02 ** 2)
            h["v1"] *= np.sqrt(1 - 0.02 ** 2)
        cfg = dict(dimensions=dims, geometry=geom, stokes=stokes, tau_calculation=1, cyclosynch=0, b_field_calc=1, epsilon_b=0.5)
        n = n_photons or 10000
        ph = make_photons(h, n, (2e12 - 0.5 * C_LIGHT / 5, 2e12 + 0.5 * C_LIGHT / 5), (1e-3, np.deg2rad(5.0)), seed)
        frame = dict(fps=5.0, time_now=2e12 / C_LIGHT)
    else:
        raise ValueError(name)
    h["fps"] = frame["fps"]
    return cfg, h, ph, frame
