"""The geometric bound behind the skipped cell re-checks (safe_path / margin_length / margin_angle in
mcrat_b200/csrc/pass_kernels.cuh), restated in numpy and attacked with random displacements: a photon inside a cell
that moves by less than the bound, in any direction, must still be inside the cell and the domain when its new
position is put through the reference's coordinate transform (Src/geometry.c:15-64) and its cell test
(Src/geometry.c:394-417).  CPU only; the device side is checked against the real re-check by
tests/test_gpu_parity.py::test_skipped_cell_rechecks_change_nothing (verified mode)."""
import numpy as np
import pytest

PI = np.pi


def to_hydro(kind, p):
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    if kind in ("2d_cart", "2d_cyl"):
        return np.stack([np.sqrt(x * x + y * y), z, np.zeros_like(x)], 1)
    if kind == "2d_sph":
        r = np.sqrt(x * x + y * y + z * z)
        return np.stack([r, np.arccos(z / r), np.zeros_like(x)], 1)
    if kind == "3d_cart":
        return p.copy()
    if kind == "3d_sph":
        r = np.sqrt(x * x + y * y + z * z)
        return np.stack([r, np.arccos(z / r), np.fmod(np.arctan2(y, x) * 180.0 / PI + 360.0, 360.0) * PI / 180], 1)
    if kind == "3d_polar":
        return np.stack([np.sqrt(x * x + y * y), np.fmod(np.arctan2(y, x) * 180.0 / PI + 360.0, 360.0) * PI / 180, z], 1)
    raise ValueError(kind)


def from_hydro(kind, h, phi):
    if kind in ("2d_cart", "2d_cyl"):
        return np.stack([h[:, 0] * np.cos(phi), h[:, 0] * np.sin(phi), h[:, 1]], 1)
    if kind == "2d_sph":
        return np.stack([h[:, 0] * np.sin(h[:, 1]) * np.cos(phi), h[:, 0] * np.sin(h[:, 1]) * np.sin(phi), h[:, 0] * np.cos(h[:, 1])], 1)
    if kind == "3d_cart":
        return h.copy()
    if kind == "3d_sph":
        return np.stack([h[:, 0] * np.sin(h[:, 1]) * np.cos(h[:, 2]), h[:, 0] * np.sin(h[:, 1]) * np.sin(h[:, 2]), h[:, 0] * np.cos(h[:, 1])], 1)
    if kind == "3d_polar":
        return np.stack([h[:, 0] * np.cos(h[:, 1]), h[:, 0] * np.sin(h[:, 1]), h[:, 2]], 1)


def margin_length(h, c, hs, lo, hi):
    m = np.minimum(hs - np.abs(h - c), np.minimum(h - lo, hi - h))
    return np.where(m > 1e-7 * np.maximum(np.abs(h), np.abs(c)), m, 0.0)


def margin_angle(h, c, hs, lo, hi, full, lever):
    m = np.minimum(hs - np.abs(h - c), np.minimum(h - lo, hi - h))
    m = np.minimum(m, np.minimum(h, full - h))
    return np.where(m > 1e-6, 0.8 * lever * np.minimum(m, 1.0), 0.0)


def safe_distance(kind, h, c, hs, dom):
    """0.5 * min over the coordinates, exactly as safe_path() forms it."""
    args = lambda k: (h[:, k], c[:, k], hs[:, k], dom[k][0], dom[k][1])
    if kind in ("2d_cart", "2d_cyl"):
        d = np.minimum(margin_length(*args(0)), margin_length(*args(1)))
    elif kind == "2d_sph":
        d = np.minimum(margin_length(*args(0)), margin_angle(*args(1), PI, h[:, 0]))
    elif kind == "3d_cart":
        d = np.minimum(np.minimum(margin_length(*args(0)), margin_length(*args(1))), margin_length(*args(2)))
    elif kind == "3d_sph":
        d = np.minimum(np.minimum(margin_length(*args(0)), margin_angle(*args(1), PI, h[:, 0])),
                       margin_angle(*args(2), 2 * PI, h[:, 0] * np.sin(h[:, 1])))
    elif kind == "3d_polar":
        d = np.minimum(np.minimum(margin_length(*args(0)), margin_angle(*args(1), 2 * PI, h[:, 0])), margin_length(*args(2)))
    return 0.5 * d


DOMAINS = {
    "2d_cart": [(0.0, 5e12), (0.0, 2.5e13), (0.0, 1.0)],
    "2d_cyl": [(0.0, 2.5e11), (1.0e12, 3.0e12), (0.0, 1.0)],
    "2d_sph": [(1.0e12, 3.0e12), (0.0, PI / 16), (0.0, 1.0)],
    "3d_cart": [(-2.5e11, 2.5e11), (-2.5e11, 2.5e11), (1.0e12, 3.0e12)],
    "3d_sph": [(1e11, 1e13), (0.0, PI / 8), (0.0, 2 * PI)],
    "3d_polar": [(0.0, 2.5e11), (0.0, 2 * PI), (1.0e12, 3.0e12)],
}


@pytest.mark.parametrize("kind", sorted(DOMAINS))
@pytest.mark.parametrize("cells_per_axis", [8, 64, 1024])
def test_a_push_shorter_than_the_bound_cannot_leave_the_cell(kind, cells_per_axis):
    rng = np.random.default_rng(hash((kind, cells_per_axis)) % 2 ** 32)
    n, nd, dom = 200_000, (2 if kind.startswith("2d") else 3), DOMAINS[kind]
    # random cells of a uniform grid over the domain, random points inside them (also hugging faces, the axis, the wrap)
    c, hs, h = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
    for k in range(nd):
        lo, hi = dom[k]
        size = (hi - lo) / cells_per_axis
        idx = rng.integers(0, cells_per_axis, n)
        idx[: n // 10] = 0                      # first cells: the axis / theta = 0 / phi = 0
        idx[n // 10: n // 5] = cells_per_axis - 1  # last cells: outer edge / phi -> 2 pi
        c[:, k] = lo + (idx + 0.5) * size
        hs[:, k] = 0.5 * size
        u = rng.uniform(-1, 1, n)
        u[::7] = np.sign(u[::7]) * (1 - 10.0 ** rng.uniform(-12, -1, u[::7].size))  # close to a face
        h[:, k] = c[:, k] + u * hs[:, k]
    inside = np.ones(n, bool)
    for k in range(nd):
        inside &= (np.abs(h[:, k] - c[:, k]) <= hs[:, k]) & (h[:, k] > dom[k][0]) & (h[:, k] < dom[k][1])
    p0 = from_hydro(kind, h, rng.uniform(0, 2 * PI, n))
    h0 = to_hydro(kind, p0)  # what the device sees (rounded transform of the rounded position)
    for k in range(nd):
        inside &= (np.abs(h0[:, k] - c[:, k]) <= hs[:, k]) & (h0[:, k] > dom[k][0]) & (h0[:, k] < dom[k][1])
    d = safe_distance(kind, h0, c, hs, dom)
    d = np.where(inside & np.isfinite(d), d, 0.0)
    assert (d > 0).mean() > 0.5  # the bound is not vacuous
    worst = -np.inf
    for trial in range(6):
        v = rng.normal(size=(n, 3))
        if trial == 0:   # straight at the axis / origin
            v = -p0.copy()
            v[:, 2] = 0 if kind != "2d_sph" and kind != "3d_sph" else v[:, 2]
        v /= np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-300)
        p1 = p0 + v * (d * (1 - 1e-9))[:, None]  # the longest path the skip allows (path counter rounds pushes up)
        h1 = to_hydro(kind, p1)
        ok = np.ones(n, bool)
        for k in range(nd):
            ok &= (np.abs(h1[:, k] - c[:, k]) <= hs[:, k]) & (h1[:, k] > dom[k][0]) & (h1[:, k] < dom[k][1])
            with np.errstate(invalid="ignore", divide="ignore"):
                worst = max(worst, np.nanmax(np.where(d > 0, (np.abs(h1[:, k] - c[:, k]) - hs[:, k]) / hs[:, k], -1)))
        bad = (d > 0) & ~ok
        assert not bad.any(), (kind, trial, int(bad.sum()), h0[bad][:3], h1[bad][:3], c[bad][:3], hs[bad][:3], d[bad][:3])
    assert worst < 0  # strictly inside every face
