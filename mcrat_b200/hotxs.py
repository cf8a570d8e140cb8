"""Thermal (hot) Klein-Nishina cross-section table, host-side builder.

The reference tabulates log10(sigma_hot / sigma_T) on a 221 x 81 grid in
(log10 h nu'/m_e c^2, log10 kT/m_e c^2) by plain Monte Carlo integration with 5e5 samples per
point (Src/hot_x_section.c:82-206, 324-357; grid constants Src/hot_x_section.h:2-10).  This
module evaluates the same double integral

    sigma_hot/sigma_T = 1/2 * int_1^{1+12 theta} dgamma int_-1^1 dmu  f_MJ(gamma; theta)
                               (1 - mu beta) sigma_KN(x gamma (1 - mu beta)) / sigma_T

deterministically (Gauss-Legendre) so that tests and benchmarks have a table in seconds.  The
file format written by :func:`write_table` is the reference's ``thermal_hot_x_section.dat``
(Src/hot_x_section.c:116-131), so tables interoperate.
"""
import numpy as np
from scipy import special

LOG_PH_E_MIN, LOG_PH_E_MAX, N_PH_E = -12.0, 6.0, 220
LOG_T_MIN, LOG_T_MAX, N_T = -4.0, 4.0, 80


def kn_cross_section(e):
    """sigma_KN/sigma_T, Src/mcrat_scattering.c:597-623."""
    e = np.asarray(e, dtype=np.float64)
    big = e >= 1e-3
    es = np.where(big, e, 1.0)
    full = 0.75 * (2.0 / (es * es) + (1.0 / (2.0 * es) - (1.0 + es) / (es ** 3)) * np.log1p(2.0 * es) +
                   (1.0 + es) / ((1.0 + 2.0 * es) ** 2))
    return np.where(big, full, 1.0 - 2.0 * e)


def maxwell_juttner(gamma, theta):
    """Src/electron.c:538-560 singleMaxwellJuttner."""
    if theta > 1e-2:
        norm = special.kve(2, 1.0 / theta)  # K_2(1/theta) e^{1/theta}
    else:
        norm = np.sqrt(np.pi * theta / 2.0)
    return gamma * np.sqrt(gamma * gamma - 1.0) / (theta * norm) * np.exp(-(gamma - 1.0) / theta)


def hot_cross_section(x, theta, order=64):
    xg, wg = np.polynomial.legendre.leggauss(order)
    g_lo, g_hi = 1.0, 1.0 + 12.0 * theta
    gam = 0.5 * (g_hi - g_lo) * xg + 0.5 * (g_hi + g_lo)
    wgam = 0.5 * (g_hi - g_lo) * wg
    mu, wmu = xg, wg
    beta = np.sqrt(gam * gam - 1.0) / gam
    fac = 1.0 - mu[None, :] * beta[:, None]
    e = x * gam[:, None] * fac
    integrand = maxwell_juttner(gam, theta)[:, None] * kn_cross_section(e) * fac
    return 0.5 * float(wgam @ integrand @ wmu)


def build_table(order=48):
    """thermal_table[i][j] = log10 sigma_hot/sigma_T, i over photon energy, j over temperature."""
    dph = (LOG_PH_E_MAX - LOG_PH_E_MIN) / N_PH_E
    dt = (LOG_T_MAX - LOG_T_MIN) / N_T
    tab = np.zeros((N_PH_E + 1, N_T + 1))
    for j in range(N_T + 1):
        theta = 10.0 ** (LOG_T_MIN + j * dt)
        for i in range(N_PH_E + 1):
            tab[i, j] = np.log10(hot_cross_section(10.0 ** (LOG_PH_E_MIN + i * dph), theta, order))
    return tab


def write_table(path, tab):
    """Src/hot_x_section.c:116-131 file layout (three header lines, a dashed line, then rows)."""
    dph = (LOG_PH_E_MAX - LOG_PH_E_MIN) / N_PH_E
    dt = (LOG_T_MAX - LOG_T_MIN) / N_T
    with open(path, "w") as f:
        f.write("The comoving photon energy and the temperatures are normalized by the electron rest mass\n")
        f.write("The calculated hot cross sections are normalized by the thompson cross section.\n")
        f.write("Photon index\tTheta Index\tlog10(Comoving Photon Energy)\tlog10(Theta)\tlog10(Hot Cross Section)\n")
        f.write("------------------------------------------------\n")
        for i in range(N_PH_E + 1):
            for j in range(N_T + 1):
                f.write("%d\t%d\t%g\t%g\t%15.10g\n" % (i, j, LOG_PH_E_MIN + i * dph, LOG_T_MIN + j * dt, tab[i, j]))


def read_table(path):
    """Src/hot_x_section.c:208-305: skip to the dashed line, then `i j logx logtheta value` rows."""
    tab = np.zeros((N_PH_E + 1, N_T + 1))
    with open(path) as f:
        for line in f:
            s = line.strip()
            if s and set(s) == {"-"}:
                break
        for line in f:
            p = line.split()
            if len(p) >= 5:
                tab[int(p[0]), int(p[1])] = float(p[4])
    return tab
