// device_math.cuh -- FP64 device functions of the MCRaT hot path (sm_100a).
//
// Everything here is evaluated in the same operation order as the reference's C code
// (file:line cited per function) and the library is compiled with -fmad=false, so the
// only source of difference from the CPU path is the last-ulp behaviour of CUDA's
// libdevice transcendental functions versus glibc's.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mcrat {

// Src/mclib.c:4-5
__device__ constexpr double A_RAD = 7.56e-15, C_LIGHT = 2.99792458e10, PL_CONST = 6.6260755e-27,
                            FINE_STRUCT = 7.29735308e-3, CHARGE_EL = 4.8032068e-10;
__device__ constexpr double K_B = 1.380658e-16, M_P = 1.6726231e-24, THOM_X_SECT = 6.65246e-25,
                            M_EL = 9.1093879e-28, R_EL = 2.817941499892705e-13;
__device__ constexpr double PI = 3.14159265358979323846;

enum { G_CARTESIAN = 0, G_SPHERICAL = 1, G_CYLINDRICAL = 2, G_POLAR = 3 };
enum { D_TWO = 0, D_TWO_POINT_FIVE = 1, D_THREE = 2 };
enum { B_INTERNAL_E = 0, B_TOTAL_E = 1, B_SIMULATION = 2 };
enum { TAU_DIRECT = 1, TAU_TABLE = 2 };

// hot cross-section table extents, Src/hot_x_section.h:2-10
constexpr int N_PH_E = 220, N_T = 80;
constexpr double LOG_PH_E_MIN = -12.0, LOG_PH_E_MAX = 6.0, LOG_T_MIN = -4.0, LOG_T_MAX = 4.0;

// ----------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Counter-based: no RNG state in memory.
// ----------------------------------------------------------------------------------------
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two uniforms strictly inside (0,1): ((52-bit integer) + 0.5) * 2^-52.  x + 0.5 with x < 2^52 needs 53 significand
// bits, so the sum and the scaling are exact: the values are the odd multiples of 2^-53, from 2^-53 to 1 - 2^-53
// (with 53 random bits the + 0.5 would be rounded away for x >= 2^52 and x = 2^53 - 1 would give exactly 1.0)
__host__ __device__ inline void philox_doubles(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1, double &a, double &b)
{
    uint32_t r[4];
    philox4x32_10(c0, c1, c2, c3, k0, k1, r);
    uint64_t x = (((uint64_t)r[0] << 32) | r[1]) >> 12;
    uint64_t y = (((uint64_t)r[2] << 32) | r[3]) >> 12;
    a = ((double)x + 0.5) * (1.0 / 4503599627370496.0);
    b = ((double)y + 0.5) * (1.0 / 4503599627370496.0);
}

// stream 0: the free-path draw of photon `slot` in while-loop iteration `iter`
__device__ inline double philox_mfp_uniform(uint32_t k0, uint32_t k1, uint64_t iter, uint32_t slot)
{
    double a, b;
    philox_doubles(slot, (uint32_t)iter, (uint32_t)(iter >> 32), 0u, k0, k1, a, b);
    return a;
}

// Sequential uniform source of one scattering event: Philox stream 1 keyed by the iteration
// number, or (parity harness) the replayed stream of the reference's gsl_rng.
struct EventRng {
    int replay;
    uint32_t k0, k1;
    uint64_t iter;
    uint64_t draw;
    const double *buf;
    unsigned long long pos, n;
    int exhausted;
    // first `npre` uniforms of this event's Philox stream, evaluated lane-parallel up front
    // (same counters, same values; only the latency moves off the scattering lane)
    const double *pre;
    int npre;

    __device__ double uniform()
    {
        if (!replay && draw < (uint64_t)npre) return pre[draw++];
        if (replay) {
            if (pos >= n) {
                exhausted = 1;
                return 0.5;
            }
            return buf[pos++];
        }
        double a, b;
        philox_doubles((uint32_t)(draw >> 1), (uint32_t)iter, (uint32_t)(iter >> 32), 1u, k0, k1, a, b);
        double u = (draw & 1ull) ? b : a;
        draw++;
        return u;
    }
    __device__ double uniform_pos()
    {
        double x;
        do {
            x = uniform();
        } while (x == 0 && !exhausted);
        return x;
    }
};

// gsl_ran_gaussian (polar Box-Muller)
__device__ inline double ran_gaussian(EventRng &r, double sigma)
{
    double x, y, r2;
    do {
        x = -1 + 2 * r.uniform_pos();
        y = -1 + 2 * r.uniform_pos();
        r2 = x * x + y * y;
    } while ((r2 > 1.0 || r2 == 0) && !r.exhausted);
    return sigma * y * sqrt(-2.0 * log(r2) / r2);
}

// ----------------------------------------------------------------------------------------
// tiny BLAS with the reference CBLAS evaluation order (gsl cblas)
// ----------------------------------------------------------------------------------------
__device__ inline double dnrm2_3(const double *X)
{
    double scale = 0.0, ssq = 1.0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double x = X[i];
        if (x != 0.0) {
            const double ax = fabs(x);
            if (scale < ax) {
                ssq = 1.0 + ssq * (scale / ax) * (scale / ax);
                scale = ax;
            } else {
                ssq += (ax / scale) * (ax / scale);
            }
        }
    }
    return scale * sqrt(ssq);
}

__device__ inline double ddot3(const double *x, const double *y)
{
    double r = 0.0;
    r += x[0] * y[0];
    r += x[1] * y[1];
    r += x[2] * y[2];
    return r;
}

template <int N>
__device__ inline void dgemv(const double *A, const double *x, double *y)
{
#pragma unroll
    for (int i = 0; i < N; i++) {
        double temp = 0.0;
#pragma unroll
        for (int j = 0; j < N; j++) temp += x[j] * A[N * i + j];
        y[i] = 0.0 + 1.0 * temp;
    }
}

// ----------------------------------------------------------------------------------------
// geometry (Src/geometry.c)
// ----------------------------------------------------------------------------------------
// Src/geometry.c:15-64 mcratCoordinateToHydroCoordinate
__device__ inline void coord_to_hydro(int dims, int g, double x, double y, double z, double &r0, double &r1,
                                      double &r2)
{
    r0 = -1; r1 = -1; r2 = -1;
    if (dims != D_THREE) {
        if (g == G_CARTESIAN || g == G_CYLINDRICAL) {
            r0 = sqrt(x * x + y * y);
            r1 = z;
        } else if (g == G_SPHERICAL) {
            r0 = sqrt(x * x + y * y + z * z);
            r1 = acos(z / r0);
        }
    } else {
        if (g == G_CARTESIAN) {
            r0 = x; r1 = y; r2 = z;
        } else if (g == G_SPHERICAL) {
            r0 = sqrt(x * x + y * y + z * z);
            r1 = acos(z / r0);
            r2 = fmod(atan2(y, x) * 180.0 / PI + 360.0, 360.0) * PI / 180;
        } else if (g == G_POLAR) {
            r0 = sqrt(x * x + y * y);
            r1 = fmod(atan2(y, x) * 180.0 / PI + 360.0, 360.0) * PI / 180;
            r2 = z;
        }
    }
}

// Src/geometry.c:108-154 hydroCoordinateToMcratCoordinate
__device__ inline void hydro_coord_to_mcrat(int dims, int g, double *out, double h0, double h1, double h2)
{
    double x = 0, y = 0, z = 0;
    const bool planar = (g == G_CARTESIAN || g == G_CYLINDRICAL);
    if (dims != D_THREE && planar) {
        x = h0 * cos(h2);
        y = h0 * sin(h2);
        z = h1;
    } else if (g == G_SPHERICAL) {
        x = h0 * sin(h1) * cos(h2);
        y = h0 * sin(h1) * sin(h2);
        z = h0 * cos(h1);
    } else if (dims == D_THREE && g == G_CARTESIAN) {
        x = h0; y = h1; z = h2;
    } else if (dims == D_THREE && g == G_POLAR) {
        x = h0 * cos(h1);
        y = h0 * sin(h1);
        z = h2;
    }
    out[0] = x; out[1] = y; out[2] = z;
}

// Src/geometry.c:189-253 hydroVectorToCartesian
__device__ inline void hydro_vector_to_cartesian(int dims, int g, double *out, double v0, double v1, double v2,
                                                 double x0, double x1, double x2)
{
    double t0 = 0, t1 = 0, t2 = 0;
    (void)x0;
    const bool planar = (g == G_CARTESIAN || g == G_CYLINDRICAL);
    if (dims == D_THREE && g == G_CARTESIAN) {
        t0 = v0; t1 = v1; t2 = v2;
    } else if (dims == D_THREE && g == G_POLAR) {
        double s1, c1;
        sincos(x1, &s1, &c1);
        t0 = v0 * c1 - v1 * s1;
        t1 = v0 * s1 + v1 * c1;
        t2 = v2;
    } else if (dims == D_TWO && planar) {
        double s2, c2;
        sincos(x2, &s2, &c2);
        t0 = v0 * c2;
        t1 = v0 * s2;
        t2 = v1;
    } else if (dims == D_TWO_POINT_FIVE && planar) {
        double s2, c2;
        sincos(x2, &s2, &c2);
        t0 = v0 * c2 - v2 * s2;
        t1 = v0 * s2 + v2 * c2;
        t2 = v1;
    } else if (g == G_SPHERICAL) {
        if (dims == D_TWO) v2 = 0;
        double s1, c1, s2, c2;
        sincos(x1, &s1, &c1);
        sincos(x2, &s2, &c2);
        t0 = v0 * s1 * c2 + v1 * c1 * c2 - v2 * s2;
        t1 = v0 * s1 * s2 + v1 * c1 * s2 + v2 * c2;
        t2 = v0 * c1 - v1 * s1;
    }
    out[0] = t0; out[1] = t1; out[2] = t2;
}

// ----------------------------------------------------------------------------------------
// Lorentz boost (Src/mclib.c:302-434)
// ----------------------------------------------------------------------------------------
// Src/mclib.c:409-434 zeroNorm
__device__ inline void zero_norm(double *p)
{
    double nrm = dnrm2_3(p + 1);
    if (p[0] != nrm) {
        p[1] = (p[1] / nrm) * p[0];
        p[2] = (p[2] / nrm) * p[0];
        p[3] = (p[3] / nrm) * p[0];
    }
}

// lorentzBoost in two halves: the matrix depends on the velocity only, so the event kernel builds
// it on a helper warp while the scattering lane is still busy (Src/mclib.c:311-347 / :348-405)
struct BoostMat {
    double L[16];
    int moving; // beta > 0
};

__device__ inline void boost_matrix(const double *b, BoostMat &M)
{
    double beta = dnrm2_3(b);
    M.moving = (beta > 0) ? 1 : 0;
    if (beta > 0) {
        double gamma = 1.0 / sqrt(1 - beta * beta);
        double *L = M.L;
        double bb = beta * beta;
        L[0] = gamma;
        L[1] = -1 * b[0] * gamma;
        L[2] = -1 * b[1] * gamma;
        L[3] = -1 * b[2] * gamma;
        L[5] = 1 + ((gamma - 1) * (b[0] * b[0]) / bb);
        L[6] = ((gamma - 1) * (b[0] * b[1] / bb));
        L[7] = ((gamma - 1) * (b[0] * b[2] / bb));
        L[10] = 1 + ((gamma - 1) * (b[1] * b[1]) / bb);
        L[11] = ((gamma - 1) * (b[1] * b[2]) / bb);
        L[15] = 1 + ((gamma - 1) * (b[2] * b[2]) / bb);
        L[4] = L[1]; L[8] = L[2]; L[12] = L[3];
        L[9] = L[6]; L[13] = L[7]; L[14] = L[11];
    }
}

__device__ inline void boost_apply(const BoostMat &M, double *p_in, double *result, bool photon)
{
    if (M.moving) {
        double pp[4];
        dgemv<4>(M.L, p_in, pp);
        if (photon) zero_norm(pp);
        result[0] = pp[0]; result[1] = pp[1]; result[2] = pp[2]; result[3] = pp[3];
    } else {
        if (photon) zero_norm(p_in);
        result[0] = p_in[0]; result[1] = p_in[1]; result[2] = p_in[2]; result[3] = p_in[3];
    }
}

// Src/mclib.c:302-407 lorentzBoost.  In the zero-velocity branch the reference renormalises
// its *input* in place (Src/mclib.c:390); p_in is therefore mutable here too.
__device__ inline void lorentz_boost(const double *b, double *p_in, double *result, bool photon)
{
    double beta = dnrm2_3(b);
    if (beta > 0) {
        double gamma = 1.0 / sqrt(1 - beta * beta);
        double L[16];
        double bb = beta * beta;
        L[0] = gamma;
        L[1] = -1 * b[0] * gamma;
        L[2] = -1 * b[1] * gamma;
        L[3] = -1 * b[2] * gamma;
        L[5] = 1 + ((gamma - 1) * (b[0] * b[0]) / bb);
        L[6] = ((gamma - 1) * (b[0] * b[1] / bb));
        L[7] = ((gamma - 1) * (b[0] * b[2] / bb));
        L[10] = 1 + ((gamma - 1) * (b[1] * b[1]) / bb);
        L[11] = ((gamma - 1) * (b[1] * b[2]) / bb);
        L[15] = 1 + ((gamma - 1) * (b[2] * b[2]) / bb);
        L[4] = L[1]; L[8] = L[2]; L[12] = L[3];
        L[9] = L[6]; L[13] = L[7]; L[14] = L[11];
        double pp[4];
        dgemv<4>(L, p_in, pp);
        if (photon) zero_norm(pp);
        result[0] = pp[0]; result[1] = pp[1]; result[2] = pp[2]; result[3] = pp[3];
    } else {
        if (photon) zero_norm(p_in);
        result[0] = p_in[0]; result[1] = p_in[1]; result[2] = p_in[2]; result[3] = p_in[3];
    }
}

// ----------------------------------------------------------------------------------------
// modified Bessel function K_2 (stands in for gsl_sf_bessel_Kn(2, x); same algorithm as the
// oracle's mc_bessel_Kn: ascending series for x <= 2, Temme/Steed continued fraction above)
// ----------------------------------------------------------------------------------------
__device__ inline double bessel_K2(double x)
{
    const double EULER = 0.57721566490153286060651209008240243;
    double k0, k1;
    if (x <= 2.0) {
        double q = 0.25 * x * x;
        double lh = log(0.5 * x);
        double term = 1.0, i0 = 1.0, s0 = 0.0, h = 0.0;
        double term1 = 1.0, i1 = 1.0;
        double psi_sum = (-EULER) + (1.0 - EULER);
        double s1 = psi_sum;
        for (int k = 1; k < 60; ++k) {
            term *= q / ((double)k * (double)k);
            h += 1.0 / (double)k;
            i0 += term;
            s0 += term * h;
            term1 *= q / ((double)k * (double)(k + 1));
            i1 += term1;
            psi_sum += 1.0 / (double)k + 1.0 / (double)(k + 1);
            s1 += term1 * psi_sum;
            if (term < 1e-18 * i0 && term1 < 1e-18 * i1) break;
        }
        k0 = -(lh + EULER) * i0 + s0;
        k1 = 1.0 / x + lh * (0.5 * x) * i1 - 0.25 * x * s1;
    } else {
        const double a1 = 0.25;
        double b = 2.0 * (1.0 + x);
        double d = 1.0 / b;
        double h = d, delh = d;
        double q1 = 0.0, q2 = 1.0;
        double q = a1, c = a1, a = -a1;
        double s = 1.0 + q * delh;
        for (int i = 2; i <= 100000; ++i) {
            a -= 2.0 * (double)(i - 1);
            c = -a * c / (double)i;
            double qnew = (q1 - b * q2) / a;
            q1 = q2;
            q2 = qnew;
            q += c * qnew;
            b += 2.0;
            d = 1.0 / (b + a * d);
            delh = (b * d - 1.0) * delh;
            h += delh;
            double dels = q * delh;
            s += dels;
            if (fabs(dels / s) < 1e-17) break;
        }
        h = a1 * h;
        double rk0 = sqrt(PI / (2.0 * x)) * exp(-x) / s;
        k0 = rk0;
        k1 = rk0 * (x + 0.5 - h) / x;
    }
    return k0 + (2.0 * 1.0 / x) * k1;
}

// ----------------------------------------------------------------------------------------
// cross sections / optical depth
// ----------------------------------------------------------------------------------------
// Src/mc_cyclosynch.c:48-52 calcDimlessTheta
__device__ inline double calc_dimless_theta(double temp) { return K_B * temp / (M_EL * C_LIGHT * C_LIGHT); }

// Src/mcrat_scattering.c:597-623 kleinNishinaCrossSection
__device__ inline double kn_cross_section(double e)
{
    if (e >= 1e-3) {
        return (3. / 4.) * (2. / (e * e) + (1. / (2. * e) - (1. + e) / (e * e * e)) * log(1. + 2. * e) +
                            (1. + e) / ((1. + 2. * e) * (1. + 2. * e)));
    }
    return (1. - 2. * e);
}

struct HotTable {
    const double *xa; // N_PH_E+1
    const double *ya; // N_T+1
    const double *za; // za[j*(N_PH_E+1)+i]
};

// The table (221 x 81 doubles + the two grids, 145 KB) is immutable while a kernel runs: every access goes through the
// read-only data cache (ld.global.nc, SASS LDG.E.CONSTANT), which the acquire loads of the persistent loop do not invalidate.
__device__ inline int interp_bsearch(const double *xa, double x, int lo, int hi)
{
    int ilo = lo, ihi = hi;
    while (ihi > ilo + 1) {
        int i = (ihi + ilo) / 2;
        if (__ldg(xa + i) > x)
            ihi = i;
        else
            ilo = i;
    }
    return ilo;
}

// gsl_interp2d bilinear (Src/hot_x_section.c:556); returns false outside the grid (GSL_EDOM)
__device__ inline bool bilinear_eval(const HotTable &t, double x, double y, double &z)
{
    const int nx = N_PH_E + 1, ny = N_T + 1;
    if (x < __ldg(t.xa) || x > __ldg(t.xa + nx - 1)) return false;
    if (y < __ldg(t.ya) || y > __ldg(t.ya + ny - 1)) return false;
    int xi = interp_bsearch(t.xa, x, 0, nx - 1);
    int yi = interp_bsearch(t.ya, y, 0, ny - 1);
    double xmin = __ldg(t.xa + xi), xmax = __ldg(t.xa + xi + 1);
    double ymin = __ldg(t.ya + yi), ymax = __ldg(t.ya + yi + 1);
    double zminmin = __ldg(t.za + yi * nx + xi);
    double zminmax = __ldg(t.za + (yi + 1) * nx + xi);
    double zmaxmin = __ldg(t.za + yi * nx + xi + 1);
    double zmaxmax = __ldg(t.za + (yi + 1) * nx + xi + 1);
    double dx = xmax - xmin, dy = ymax - ymin;
    double tt = (x - xmin) / dx;
    double u = (y - ymin) / dy;
    z = (1. - tt) * (1. - u) * zminmin + tt * (1. - u) * zmaxmin + (1. - tt) * u * zminmax + tt * u * zmaxmax;
    return true;
}

// Src/electron.c:538-560 singleMaxwellJuttner with its normalisation (K_2(1/theta) e^{1/theta}, or the asymptotic form
// below theta = 1e-2) formed once per integral: the same value in every call of the reference
__device__ inline double maxwell_juttner_norm(double theta)
{
    if (theta > 1.e-2) return bessel_K2(1. / theta) * exp(1. / theta);
    return sqrt(PI * theta / 2.);
}
__device__ inline double maxwell_juttner_pdf(double gamma, double theta, double normalization)
{
    return ((gamma * sqrt(gamma * gamma - 1.) / (theta * normalization)) * exp(-(gamma - 1.) / theta));
}

// Src/hot_x_section.c:369-400 boostedCrossSection
__device__ inline double boosted_cross_section(double norm_ph_comv, double mu, double gamma)
{
    double beta = sqrt(gamma * gamma - 1.) / gamma;
    double norm_ph_e = norm_ph_comv * gamma * (1. - mu * beta);
    return kn_cross_section(norm_ph_e) * (1. - mu * beta);
}

// Where a hot cross-section lookup falls outside the table the reference integrates the cross section on the spot by
// plain Monte Carlo with its gsl_rng (calculateTotalThermalCrossSection, Src/hot_x_section.c:324-357, reached from
// interpolateThermalHotCrossSection :563-599): 500 000 samples of (gamma, mu) over [1, 1 + 12 theta] x [-1, 1],
// gsl_monte_plain's running mean.  The device does the same integral, sample for sample, from a Philox stream of its own
// keyed by the photon and the iteration: counter (sample, iteration_lo, slot, 3 + 8 * iteration_hi), the two doubles of a
// block = the sample's (gamma, mu) draws.  One thread, sequentially, like the reference (a catastrophic path there as
// here: ~0.4 s per lookup; photons outside the table in the same launch integrate in parallel).  The replay harness
// (uniforms of a recorded gsl_rng stream) cannot supply the 10^6 draws in stream order from parallel threads: there the
// lookup is reported as MCRAT_B200_ERR_TABLE, as in round 1.
struct FallbackRng {
    uint32_t k0, k1, slot;
    uint64_t iter;
    int replay;
};

__device__ __noinline__ double total_thermal_cross_section_mc(double ph_comv, double theta, uint32_t k0, uint32_t k1, uint32_t slot,
                                                              uint64_t iter)
{
    const double normalization = maxwell_juttner_norm(theta);
    const double xl0 = 1, xu0 = 1. + 12 * theta, xl1 = -1, xu1 = 1;
    const double vol = (xu0 - xl0) * (xu1 - xl1);
    const uint32_t c1 = (uint32_t)iter, c3 = 3u + ((uint32_t)(iter >> 32) << 3);
    double m = 0; // gsl_monte_plain_integrate: m += (f - m) / (n + 1)
    for (uint32_t n = 0; n < 500000u; ++n) {
        double u1, u2;
        philox_doubles(n, c1, slot, c3, k0, k1, u1, u2);
        const double gamma = xl0 + u1 * (xu0 - xl0);
        const double mu = xl1 + u2 * (xu1 - xl1);
        const double fval = maxwell_juttner_pdf(gamma, theta, normalization) * boosted_cross_section(ph_comv, mu, gamma);
        const double dd = fval - m;
        m += dd / (n + 1.0);
    }
    return 0.5 * (vol * m);
}

// Src/optical_depth.c:132-149 getThermalCrossSection + Src/hot_x_section.c:545-605.
// Outside the table: the two closed-form early returns of calculateTotalThermalCrossSection (theta below the table,
// :336-339), else the Monte Carlo integral above (`fr` given and not in replay mode) or `*table_err`.
__device__ inline double thermal_cross_section(int tau_calc, const HotTable &t, double comv_e, double temp,
                                               int *table_err, const FallbackRng *fr = nullptr)
{
    if (tau_calc != TAU_TABLE) return 1;
    double ne = comv_e / (M_EL * C_LIGHT);
    double theta = calc_dimless_theta(temp);
    double le = log10(ne), lt = log10(theta);
    double res;
    if (!bilinear_eval(t, le, lt, res)) {
        double ph_comv = pow(10.0, le);
        double th = pow(10.0, lt);
        double direct;
        if (th < pow(10.0, LOG_T_MIN) && ph_comv < pow(10.0, LOG_PH_E_MIN)) {
            direct = 1;
        } else if (th < pow(10.0, LOG_T_MIN)) {
            direct = kn_cross_section(ph_comv);
        } else if (fr && !fr->replay) {
            direct = total_thermal_cross_section_mc(ph_comv, th, fr->k0, fr->k1, fr->slot, fr->iter);
        } else {
            if (table_err) *table_err = 1;
            direct = kn_cross_section(ph_comv);
        }
        res = log10(direct);
    }
    return pow(10.0, res);
}

// cell fields a photon needs once it is located
struct CellState {
    double v0, v1, v2, r0, r1, r2, gamma, dens_lab, temp;
};

// Src/optical_depth.c:7-115 calculateOpticalDepth (NONTHERMAL_E_DIST == OFF)
__device__ inline double optical_depth(int dims, int g, int tau_calc, const HotTable &t, const CellState &c,
                                       double ph_r0, double ph_r1, double p1, double p2, double p3, double comv_p0,
                                       int *table_err, const FallbackRng *fr = nullptr)
{
    double fb[3];
    if (dims == D_THREE) {
        hydro_vector_to_cartesian(dims, g, fb, c.v0, c.v1, c.v2, c.r0, c.r1, c.r2);
    } else if (dims == D_TWO_POINT_FIVE) {
        double ph_phi = atan2(ph_r1, ph_r0);
        hydro_vector_to_cartesian(dims, g, fb, c.v0, c.v1, c.v2, c.r0, c.r1, ph_phi);
    } else {
        double ph_phi = atan2(ph_r1, ph_r0);
        hydro_vector_to_cartesian(dims, g, fb, c.v0, c.v1, 0, c.r0, c.r1, ph_phi);
    }
    double fl_v_norm = sqrt(fb[0] * fb[0] + fb[1] * fb[1] + fb[2] * fb[2]);
    double ph_v_norm = sqrt(p1 * p1 + p2 * p2 + p3 * p3);
    double n_cosangle = ((fb[0] * p1) + (fb[1] * p2) + (fb[2] * p3)) / (fl_v_norm * ph_v_norm);
    double beta = sqrt(1.0 - 1.0 / (c.gamma * c.gamma));
    double fluid_factor = (1.0 - beta * n_cosangle);
    double n_lab = c.dens_lab / M_P;
    double sig = thermal_cross_section(tau_calc, t, comv_p0, c.temp, table_err, fr);
    return (n_lab) * (THOM_X_SECT * sig) * fluid_factor;
}

// ----------------------------------------------------------------------------------------
// Stokes-plane rotations (Src/mcrat_scattering.c:10-149)
// ----------------------------------------------------------------------------------------
// Src/mcrat_scattering.c:10-39 mullerMatrixRotation.  The 4x4 dgemv of the reference reduces to
// the two middle rows (rows 0 and 3 are unit rows; the zero products add exactly 0 for finite s)
__device__ inline void muller_rotation(double theta, double *s)
{
    double sn, cs;
    sincos(2 * theta, &sn, &cs);
    double r1 = 0.0;
    r1 += s[1] * cs;
    r1 += s[2] * (-1 * sn);
    double r2 = 0.0;
    r2 += s[1] * sn;
    r2 += s[2] * cs;
    s[1] = r1;
    s[2] = r2;
}

// Src/mcrat_scattering.c:41-65 findXY
__device__ inline void find_xy(const double *v, const double *a, double *x, double *y)
{
    y[0] = (v[1] * a[2] - v[2] * a[1]);
    y[1] = -1 * (v[0] * a[2] - v[2] * a[0]);
    y[2] = (v[0] * a[1] - v[1] * a[0]);
    double norm = 1.0 / sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
    y[0] *= norm; y[1] *= norm; y[2] *= norm;
    x[0] = y[1] * v[2] - y[2] * v[1];
    x[1] = -1 * (y[0] * v[2] - y[2] * v[0]);
    x[2] = y[0] * v[1] - y[1] * v[0];
    norm = 1.0 / sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    x[0] *= norm; x[1] *= norm; x[2] *= norm;
}

// Src/mcrat_scattering.c:67-101 findPhi
__device__ inline double find_phi(const double *x_old, const double *y_old, const double *y_new)
{
    double dot = ddot3(x_old, y_new);
    double factor = (dot > 0) ? 1.0 : ((dot < 0) ? -1.0 : 0.0);
    dot = ddot3(y_old, y_new);
    if ((dot < -1) || (dot > 1)) dot = round(dot);
    return -1 * factor * acos(dot);
}

// One half of stokesRotation: the angle between the polarisation bases (k, a) and (k, b),
// Src/mcrat_scattering.c:115-123 / :133-141.  The angle depends on momenta only -- never on the
// Stokes vector -- which is what lets the event kernel evaluate all of them on a second warp.
__device__ inline double stokes_angle(const double *k, const double *a, const double *b)
{
    double x[3], y[3], xn[3], yn[3];
    find_xy(k, a, x, y);
    find_xy(k, b, xn, yn);
    return find_phi(x, y, yn);
}

// general form: first basis from (k1, a), second from (k2, b) -- the scattering-plane rotation of
// singleScatter pairs two different wave vectors (Src/mcrat_scattering.c:402-405)
__device__ inline double stokes_angle4(const double *k1, const double *a, const double *k2, const double *b)
{
    double x[3], y[3], xn[3], yn[3];
    find_xy(k1, a, x, y);
    find_xy(k2, b, xn, yn);
    return find_phi(x, y, yn);
}

// mullerMatrixRotation with sin(2 theta), cos(2 theta) already evaluated (same products, same order)
__device__ inline void muller_rotation_sc(double sn, double cs, double *s)
{
    double r1 = 0.0;
    r1 += s[1] * cs;
    r1 += s[2] * (-1 * sn);
    double r2 = 0.0;
    r2 += s[1] * sn;
    r2 += s[2] * cs;
    s[1] = r1;
    s[2] = r2;
}

// Src/mcrat_scattering.c:103-149 stokesRotation
__device__ inline void stokes_rotation(const double *v, const double *v_ph, const double *v_ph_boosted, double *s)
{
    const double z_hat[3] = {0, 0, 1};
    double x[3], y[3], xn[3], yn[3];
    find_xy(v_ph, z_hat, x, y);
    find_xy(v_ph, v, xn, yn);
    double phi = find_phi(x, y, yn);
    muller_rotation(phi, s);
    find_xy(v_ph_boosted, v, x, y);
    find_xy(v_ph_boosted, z_hat, xn, yn);
    phi = find_phi(x, y, yn);
    muller_rotation(phi, s);
}

// ----------------------------------------------------------------------------------------
// Klein-Nishina scatter (Src/mcrat_scattering.c:151-595)
// ----------------------------------------------------------------------------------------
// Src/mcrat_scattering.c:509-595 kleinNishinaScatter.  The reference's pow(m, -1/-2/-3) are written
// as reciprocals of products (<= 2 ulp apart; pow is several hundred instructions in FP64).
struct KnTheta {
    double er, st, ct, imu2, f_theta; // st, ct = sin / cos of the polar angle
};

// first half: accept / reject and the polar angle (Src/mcrat_scattering.c:519-553)
__device__ inline int kn_accept_theta(double &theta, double p0, KnTheta &k, EventRng &rng)
{
    double er = p0 / (M_EL * C_LIGHT);
    double kn = kn_cross_section(er);
    double rand_num = rng.uniform();
    if (!(rand_num <= kn)) return 0;
    double cty = 1, ct = 0, fct = 0;
    while (cty > fct && !rng.exhausted) {
        cty = rng.uniform() * 2;
        ct = rng.uniform() * 2 - 1;
        double m1 = (1 + er * (1 - ct));
        fct = (1.0 / (m1 * m1)) * (er * (1 - ct) + (1 / m1) + ct * ct);
    }
    theta = acos(ct);
    double st, ctheta;
    sincos(theta, &st, &ctheta);
    double mu = 1 + er * (1 - ctheta);
    double imu2 = 1.0 / (mu * mu);
    k.er = er;
    k.st = st;
    k.ct = ctheta;
    k.imu2 = imu2;
    k.f_theta = ((1.0 / mu) + (1.0 / (mu * mu * mu)) - imu2 * st * st) * st;
    return 1;
}

// second half: the azimuth, which is the only place the Stokes vector (q, u) enters the
// kinematics (Src/mcrat_scattering.c:555-585)
__device__ inline double kn_phi(int stokes, const KnTheta &k, double q, double u, EventRng &rng)
{
    const double st = k.st, imu2 = k.imu2, f_theta = k.f_theta;
    double phi_y = 1, f_phi = 0, phi_dum = 0;
    if (!stokes || (u == 0 && q == 0)) {
        phi_dum = rng.uniform() * 2 * PI;
    } else {
        double phi_max = fabs(atan2(-u, q)) / 2.0;
        double s2m, c2m;
        sincos(2 * phi_max, &s2m, &c2m);
        double norm = (f_theta + imu2 * st * st * st * (q * c2m - u * s2m));
        while (phi_y > f_phi && !rng.exhausted) {
            phi_y = rng.uniform();
            phi_dum = rng.uniform() * 2 * PI;
            double s2, c2;
            sincos(2 * phi_dum, &s2, &c2);
            f_phi = (f_theta + imu2 * st * st * st * (q * c2 - u * s2)) / norm;
        }
    }
    return phi_dum;
}

__device__ inline int kn_scatter(int stokes, double &theta, double &phi, double p0, double q, double u, EventRng &rng)
{
    KnTheta k;
    if (!kn_accept_theta(theta, p0, k, rng)) return 0;
    phi = kn_phi(stokes, k, q, u, rng);
    return 1;
}

// Src/mcrat_scattering.c:151-485 singleScatter in one piece.  The kernels use the staged form below
// (scatter_stage_*), which is this function cut at its data dependencies; it stays here as the readable statement
// of the whole scatter in the reference's order.
__device__ inline int single_scatter(int stokes, double *el_comov, double *ph_comov, double *s, EventRng &rng)
{
    const double z_axis[3] = {0, 0, 1};
    double el_v[3], neg_el_v[3], php[4], elp[4];
    double rot[9], result0[3], result1[3], result[4], orig[4];
    double *ph_p = php + 1;

    el_v[0] = el_comov[1] / el_comov[0];
    el_v[1] = el_comov[2] / el_comov[0];
    el_v[2] = el_comov[3] / el_comov[0];

    lorentz_boost(el_v, el_comov, elp, false);
    lorentz_boost(el_v, ph_comov, php, true);

    if (stokes) stokes_rotation(el_v, ph_comov + 1, php + 1, s);

    orig[0] = php[0]; orig[1] = php[1]; orig[2] = php[2]; orig[3] = php[3];

    double phi0 = atan2(php[2], php[1]);
    double s0m, c0m; // sin(-phi0), cos(-phi0)
    sincos(-phi0, &s0m, &c0m);
#pragma unroll
    for (int i = 0; i < 9; i++) rot[i] = 0;
    rot[8] = 1;
    rot[0] = c0m;
    rot[4] = c0m;
    rot[1] = -s0m;
    rot[3] = s0m;
    dgemv<3>(rot, ph_p, result0);
    php[1] = result0[0];
    php[2] = 0;
    php[3] = result0[2];

    double phi1 = atan2(result0[2], result0[0]);
    double s1m, c1m; // sin(-phi1), cos(-phi1)
    sincos(-phi1, &s1m, &c1m);
#pragma unroll
    for (int i = 0; i < 9; i++) rot[i] = 0;
    rot[4] = 1;
    rot[0] = c1m;
    rot[8] = c1m;
    rot[2] = -s1m;
    rot[6] = s1m;
    dgemv<3>(rot, ph_p, result1);
    php[1] = php[0];
    php[2] = result1[1];
    php[3] = 0;

    double theta = 0, phi = 0;
    int occurred = kn_scatter(stokes, theta, phi, php[0], s[1], s[2], rng);
    if (occurred == 1) {
        double sth, cth, sph, cph;
        sincos(theta, &sth, &cth);
        sincos(phi, &sph, &cph);
        result[0] = (php[0]) / (1 + (((php[0]) * (1 - cth)) / (M_EL * C_LIGHT)));
        result[1] = result[0] * cth;
        result[2] = result[0] * sth * sph;
        result[3] = result[0] * sth * cph;

        php[0] = result[0]; php[1] = result[1]; php[2] = result[2]; php[3] = result[3];
#pragma unroll
        for (int i = 0; i < 9; i++) rot[i] = 0;
        rot[4] = 1;
        rot[0] = c1m;
        rot[8] = c1m;
        rot[2] = s1m;
        rot[6] = -s1m;
        dgemv<3>(rot, ph_p, result1);
        php[1] = result1[0]; php[2] = result1[1]; php[3] = result1[2];
#pragma unroll
        for (int i = 0; i < 9; i++) rot[i] = 0;
        rot[8] = 1;
        rot[0] = c0m;
        rot[4] = c0m;
        rot[1] = s0m;
        rot[3] = -s0m;
        dgemv<3>(rot, ph_p, result0);

        if (stokes) {
            double xt[3], yt[3], xtn[3], ytn[3];
            find_xy(orig + 1, z_axis, xt, yt);
            find_xy(result0, orig + 1, xtn, ytn);
            double ph = find_phi(xt, yt, ytn);
            muller_rotation(ph, s);

            double th = acos((orig[1] * result0[0] + orig[2] * result0[1] + orig[3] * result0[2]) / (orig[0] * (php[0])));
            double ct, sn;
            sincos(th, &sn, &ct);
            double scatt[16];
#pragma unroll
            for (int i = 0; i < 16; i++) scatt[i] = 0;
            scatt[0] = 1.0 + ct * ct + ((1 - ct) * (orig[0] - result[0]) / (M_EL * C_LIGHT));
            scatt[1] = sn * sn;
            scatt[4] = sn * sn;
            scatt[5] = 1.0 + ct * ct;
            scatt[10] = 2.0 * ct;
            scatt[15] = 2.0 * ct + ((ct) * (1 - ct) * (orig[0] - result[0]) / (M_EL * C_LIGHT));
            double sr[4];
            dgemv<4>(scatt, s, sr);
            s[0] = sr[0] / sr[0];
            s[1] = sr[1] / sr[0];
            s[2] = sr[2] / sr[0];
            s[3] = sr[3] / sr[0];

            find_xy(result0, orig + 1, xt, yt);
            find_xy(result0, z_axis, xtn, ytn);
            ph = find_phi(xt, yt, ytn);
            muller_rotation(ph, s);
        }
        php[1] = result0[0]; php[2] = result0[1]; php[3] = result0[2];
        neg_el_v[0] = (-1 * el_v[0]);
        neg_el_v[1] = (-1 * el_v[1]);
        neg_el_v[2] = (-1 * el_v[2]);
        lorentz_boost(neg_el_v, php, ph_comov, true);
        if (stokes) stokes_rotation(neg_el_v, php + 1, ph_comov + 1, s);
    }
    return occurred;
}

// ----------------------------------------------------------------------------------------
// singleScatter cut into stages for the two-warp event (mcrat_b200.cu, scatter_candidate_2w):
// warp 0 runs the momentum chain below, warp 1 evaluates the Stokes-plane angles, which depend on
// momenta only.  Every stage is the corresponding statement block of single_scatter above,
// operation for operation, so both paths give bit-identical results.
// ----------------------------------------------------------------------------------------
struct ScatterRot {
    double s0m, c0m, s1m, c1m; // sin/cos of -phi0 and -phi1 (Src/mcrat_scattering.c:262-296)
};

// Src/mcrat_scattering.c:216-240: electron velocity, photon into the electron rest frame
__device__ inline void scatter_stage_boost(double *el_comov, double *ph_comov, double *el_v, double *php)
{
    el_v[0] = el_comov[1] / el_comov[0];
    el_v[1] = el_comov[2] / el_comov[0];
    el_v[2] = el_comov[3] / el_comov[0];
    lorentz_boost(el_v, ph_comov, php, true);
}

// Src/mcrat_scattering.c:262-296: rotate the photon onto +x; returns its energy (p0 is unchanged)
__device__ inline void scatter_stage_align(const double *php, ScatterRot &r)
{
    double rot[9], result0[3], tmp[3];
    double phi0 = atan2(php[2], php[1]);
    sincos(-phi0, &r.s0m, &r.c0m);
#pragma unroll
    for (int i = 0; i < 9; i++) rot[i] = 0;
    rot[8] = 1;
    rot[0] = r.c0m;
    rot[4] = r.c0m;
    rot[1] = -r.s0m;
    rot[3] = r.s0m;
    tmp[0] = php[1]; tmp[1] = php[2]; tmp[2] = php[3];
    dgemv<3>(rot, tmp, result0);
    double phi1 = atan2(result0[2], result0[0]);
    sincos(-phi1, &r.s1m, &r.c1m);
}

// Src/mcrat_scattering.c:318-386: scattered photon in the electron rest frame, original axes
// (sth, cth) = sincos(theta) as kn_accept_theta already evaluated it
__device__ inline void scatter_stage_out(double e_in, double sth, double cth, double phi, const ScatterRot &r, double *out)
{
    double rot[9], result[4], v[3], result1[3], result0[3];
    double sph, cph;
    sincos(phi, &sph, &cph);
    result[0] = (e_in) / (1 + (((e_in) * (1 - cth)) / (M_EL * C_LIGHT)));
    result[1] = result[0] * cth;
    result[2] = result[0] * sth * sph;
    result[3] = result[0] * sth * cph;
    v[0] = result[1]; v[1] = result[2]; v[2] = result[3];
#pragma unroll
    for (int i = 0; i < 9; i++) rot[i] = 0;
    rot[4] = 1;
    rot[0] = r.c1m;
    rot[8] = r.c1m;
    rot[2] = r.s1m;
    rot[6] = -r.s1m;
    dgemv<3>(rot, v, result1);
#pragma unroll
    for (int i = 0; i < 9; i++) rot[i] = 0;
    rot[8] = 1;
    rot[0] = r.c0m;
    rot[4] = r.c0m;
    rot[1] = r.s0m;
    rot[3] = -r.s0m;
    dgemv<3>(rot, result1, result0);
    out[0] = result[0]; out[1] = result0[0]; out[2] = result0[1]; out[3] = result0[2];
}

// Src/mcrat_scattering.c:408-416: scattering angle and the non-zero entries of the Fano matrix
__device__ inline void scatter_stage_fano(const double *orig, const double *out, double *f)
{
    double th = acos((orig[1] * out[1] + orig[2] * out[2] + orig[3] * out[3]) / (orig[0] * (out[0])));
    double ct, sn;
    sincos(th, &sn, &ct);
    f[0] = 1.0 + ct * ct + ((1 - ct) * (orig[0] - out[0]) / (M_EL * C_LIGHT));
    f[1] = sn * sn;
    f[2] = 1.0 + ct * ct;
    f[3] = 2.0 * ct;
    f[4] = 2.0 * ct + ((ct) * (1 - ct) * (orig[0] - out[0]) / (M_EL * C_LIGHT));
}

// Src/mcrat_scattering.c:418-433: s <- T s / (T s)_0 with the full 4x4 product of the reference
__device__ inline void fano_apply(const double *f, double *s)
{
    double scatt[16];
#pragma unroll
    for (int i = 0; i < 16; i++) scatt[i] = 0;
    scatt[0] = f[0];
    scatt[1] = f[1];
    scatt[4] = f[1];
    scatt[5] = f[2];
    scatt[10] = f[3];
    scatt[15] = f[4];
    double sr[4];
    dgemv<4>(scatt, s, sr);
    s[0] = sr[0] / sr[0];
    s[1] = sr[1] / sr[0];
    s[2] = sr[2] / sr[0];
    s[3] = sr[3] / sr[0];
}

// ----------------------------------------------------------------------------------------
// electron sampling (Src/electron.c:70-237)
// ----------------------------------------------------------------------------------------
// Src/electron.c:202-237 sampleThermalElectron
__device__ inline double sample_thermal_electron(double temp, EventRng &rng, double k2_known = 0)
{
    double gamma = 1;
    if (temp >= 1e7) {
        double factor = K_B * temp / (M_EL * C_LIGHT * C_LIGHT);
        double k2 = (k2_known > 0) ? k2_known : bessel_K2(1.0 / factor);
        double y = 1, f = 0, x = 0;
        while (((f != f) || (y > f)) && !rng.exhausted) {
            x = rng.uniform_pos() * (1 + 100 * factor);
            double bx = sqrt(1 - (1 / (x * x)));
            y = rng.uniform() / 2.0;
            f = x * x * (bx / k2) * exp(-1 * x / factor);
        }
        gamma = x;
    } else {
        double factor = sqrt(K_B * temp / M_EL);
        double g1 = ran_gaussian(rng, factor);
        double g2 = ran_gaussian(rng, factor);
        double g3 = ran_gaussian(rng, factor);
        double a = g1 / C_LIGHT, b = g2 / C_LIGHT, c = g3 / C_LIGHT;
        gamma = 1.0 / sqrt(1 - (a * a + b * b + c * c));
    }
    return gamma;
}

// Src/electron.c:126-175 rotateElectron; the two rotation angles depend on the photon only
struct ElRot {
    double stt, ctt, spm, cpm; // sin / cos of ph_theta and of -ph_phi
};

__device__ inline void electron_rot_angles(const double *ph_p, ElRot &r)
{
    double ph_phi = atan2(ph_p[2], ph_p[3]);
    double ph_theta = atan2(sqrt(ph_p[2] * ph_p[2] + ph_p[3] * ph_p[3]), ph_p[1]);
    sincos(ph_theta, &r.stt, &r.ctt);
    sincos(-ph_phi, &r.spm, &r.cpm);
}

__device__ inline void rotate_electron_sc(double *el_p, const ElRot &r)
{
    double rot[9], result[3];
    double *e = el_p + 1;
    const double stt = r.stt, ctt = r.ctt, spm = r.spm, cpm = r.cpm;
#pragma unroll
    for (int i = 0; i < 9; i++) rot[i] = 0;
    rot[4] = 1;
    rot[8] = ctt;
    rot[0] = ctt;
    rot[2] = -stt;
    rot[6] = stt;
    dgemv<3>(rot, e, result);
#pragma unroll
    for (int i = 0; i < 9; i++) rot[i] = 0;
    rot[0] = 1;
    rot[4] = cpm;
    rot[8] = cpm;
    rot[5] = -spm;
    rot[7] = spm;
    double out[3];
    dgemv<3>(rot, result, out);
    e[0] = out[0]; e[1] = out[1]; e[2] = out[2];
}

__device__ inline void rotate_electron(double *el_p, const double *ph_p)
{
    ElRot r;
    electron_rot_angles(ph_p, r);
    rotate_electron_sc(el_p, r);
}

// the three polar Box-Muller samples of the Maxwellian branch (Src/electron.c:231-233) evaluated by a
// whole warp: lane j works out candidate pair j of the prefetched Philox stream, the first three
// accepted candidates are the three gaussians the sequential loop would have produced (a candidate
// always costs exactly two draws: Philox doubles are never 0).  Returns the number of draws
// consumed, or 0 if the window held fewer than three accepted pairs (caller falls back).
__device__ inline int warp_gaussians3(const double *pre, int npre, uint64_t draw0, double sigma, double *g)
{
    const int lane = threadIdx.x & 31;
    const uint64_t k = draw0 + 2ull * (uint64_t)lane;
    const bool valid = (k + 1 < (uint64_t)npre);
    double x = 2, y = 2;
    if (valid) {
        x = -1 + 2 * pre[k];
        y = -1 + 2 * pre[k + 1];
    }
    const double r2 = x * x + y * y;
    const bool ok = valid && !(r2 > 1.0 || r2 == 0);
    const double val = sigma * y * sqrt(-2.0 * log(ok ? r2 : 0.5) / (ok ? r2 : 0.5));
    unsigned mask = __ballot_sync(0xffffffffu, ok);
    if (__popc(mask) < 3) return 0;
    int j = 0;
#pragma unroll
    for (int n = 0; n < 3; ++n) {
        j = __ffs(mask) - 1;
        mask &= mask - 1;
        g[n] = __shfl_sync(0xffffffffu, val, j);
    }
    return 2 * (j + 1);
}

// The Maxwell-Juttner branch (Src/electron.c:207-226) is a rejection loop over a flat envelope: 50 (theta ~ 1) to
// 330 (theta ~ 0.002) trials per electron, each two uniforms, a square root, a division and an exponential -- 40 000
// to 390 000 cycles on one lane.  Here the lanes of the warp evaluate 64 consecutive trials per round from the same Philox
// counters the sequential loop would reach (a trial always costs exactly two draws: Philox doubles are never 0), and
// the first accepted trial in stream order wins: same gamma, same number of draws consumed (the return value).
// 0 = no acceptance within `max_rounds` rounds; the caller then continues sequentially from where this stopped.
__device__ inline void mj_trial(uint32_t k0, uint32_t k1, uint64_t iter, uint64_t dd, double span, double factor, double k2,
                                double &x, bool &accepted)
{
    double a, b, u1, u2;
    philox_doubles((uint32_t)(dd >> 1), (uint32_t)iter, (uint32_t)(iter >> 32), 1u, k0, k1, a, b);
    if (dd & 1ull) {
        u1 = b;
        philox_doubles((uint32_t)((dd + 1) >> 1), (uint32_t)iter, (uint32_t)(iter >> 32), 1u, k0, k1, a, b);
        u2 = a;
    } else {
        u1 = a;
        u2 = b;
    }
    x = u1 * span;
    const double bx = sqrt(1 - (1 / (x * x)));
    const double y = u2 / 2.0;
    const double f = x * x * (bx / k2) * exp(-1 * x / factor);
    accepted = !((f != f) || (y > f));
}

// one round = 64 consecutive trials: lane j evaluates trials 64 r + j and 64 r + 32 + j (two independent chains)
__device__ inline uint64_t warp_mj_gamma(uint32_t k0, uint32_t k1, uint64_t iter, uint64_t draw0, double factor, double k2,
                                         int max_rounds, double &gamma)
{
    const int lane = threadIdx.x & 31;
    const double span = (1 + 100 * factor);
    for (int r = 0; r < max_rounds; ++r) {
        double xa, xb;
        bool acc_a, acc_b;
        mj_trial(k0, k1, iter, draw0 + 2ull * (uint64_t)(64 * r + lane), span, factor, k2, xa, acc_a);
        mj_trial(k0, k1, iter, draw0 + 2ull * (uint64_t)(64 * r + 32 + lane), span, factor, k2, xb, acc_b);
        const unsigned ma = __ballot_sync(0xffffffffu, acc_a);
        const unsigned mb = __ballot_sync(0xffffffffu, acc_b);
        if (ma) {
            const int j = __ffs(ma) - 1;
            gamma = __shfl_sync(0xffffffffu, xa, j);
            return 2ull * (uint64_t)(64 * r + j + 1);
        }
        if (mb) {
            const int j = __ffs(mb) - 1;
            gamma = __shfl_sync(0xffffffffu, xb, j);
            return 2ull * (uint64_t)(64 * r + 32 + j + 1);
        }
    }
    return 0;
}

// gamma of a Maxwellian electron from its three velocity components, Src/electron.c:231-236
__device__ inline double maxwellian_gamma(const double *g)
{
    double a = g[0] / C_LIGHT, b = g[1] / C_LIGHT, c = g[2] / C_LIGHT;
    return 1.0 / sqrt(1 - (a * a + b * b + c * c));
}

// singleThermalElectron after the Lorentz factor is known: direction draws, 4-momentum, rotation
__device__ inline void thermal_electron_from_gamma(double *el_p, double gamma, const ElRot &er, EventRng &rng)
{
    double beta = sqrt(1 - (1 / (gamma * gamma)));
    double phi = rng.uniform() * 2 * PI;
    double theta = acos((1 - sqrt(1 + beta * beta + 2 * beta - 4 * beta * rng.uniform())) / beta);
    double sth, cth, sph, cph;
    sincos(theta, &sth, &cth);
    sincos(phi, &sph, &cph);
    el_p[0] = gamma * (M_EL) * (C_LIGHT);
    el_p[1] = gamma * (M_EL) * (C_LIGHT)*beta * cth;
    el_p[2] = gamma * (M_EL) * (C_LIGHT)*beta * sth * sph;
    el_p[3] = gamma * (M_EL) * (C_LIGHT)*beta * sth * cph;
    rotate_electron_sc(el_p, er);
}

// Src/electron.c:70-94 singleThermalElectron (+ :177-200 sampleElectronTheta)
__device__ inline void single_thermal_electron(double *el_p, double temp, const double *ph_p, EventRng &rng)
{
    double gamma = sample_thermal_electron(temp, rng);
    double beta = sqrt(1 - (1 / (gamma * gamma)));
    double phi = rng.uniform() * 2 * PI;
    double theta = acos((1 - sqrt(1 + beta * beta + 2 * beta - 4 * beta * rng.uniform())) / beta);
    double sth, cth, sph, cph;
    sincos(theta, &sth, &cth);
    sincos(phi, &sph, &cph);
    el_p[0] = gamma * (M_EL) * (C_LIGHT);
    el_p[1] = gamma * (M_EL) * (C_LIGHT)*beta * cth;
    el_p[2] = gamma * (M_EL) * (C_LIGHT)*beta * sth * sph;
    el_p[3] = gamma * (M_EL) * (C_LIGHT)*beta * sth * cph;
    rotate_electron(el_p, ph_p);
}

// ----------------------------------------------------------------------------------------
// cyclo-synchrotron helpers (Src/mc_cyclosynch.c:30-92)
// ----------------------------------------------------------------------------------------
__device__ inline double calc_cyclotron_freq(double b) { return CHARGE_EL * b / (2 * PI * M_EL * C_LIGHT); }

__device__ inline double calc_b(int b_field_calc, double epsilon_b, double el_dens, double temp)
{
    if (b_field_calc == B_INTERNAL_E) return sqrt(epsilon_b * 8 * PI * 3 * el_dens * K_B * temp / 2);
    if (b_field_calc == B_TOTAL_E)
        return sqrt(8 * PI * epsilon_b * (el_dens * M_P * C_LIGHT * C_LIGHT + 4 * A_RAD * temp * temp * temp * temp / 3));
    return 0;
}

} // namespace mcrat
