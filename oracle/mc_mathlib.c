/*
 * mc_mathlib.c -- TEST INFRASTRUCTURE (oracle). See mc_mathlib.h.
 */
#include "mc_mathlib.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================== */
/* RANLXS (GSL rng/ranlxs.c, luxury levels 0/1/2 = 109/202/397)             */
/* ======================================================================== */
static const int lxs_next[12] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 0};
static const int lxs_snext[24] = {1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12,
                                  13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 0};
static const double lxs_sbase = 16777216.0;                 /* 2^24  */
static const double lxs_sone_bit = 1.0 / 16777216.0;        /* 2^-24 */
static const double lxs_one_bit = 1.0 / 281474976710656.0;  /* 2^-48 */
static const double lxs_shift = 268435456.0;                /* 2^28  */

#define LXS_STEP(x1, x2, i1, i2, i3) \
    x1 = xdbl[i1] - xdbl[i2];        \
    if (x2 < 0) {                    \
        x1 -= lxs_one_bit;           \
        x2 += 1;                     \
    }                                \
    xdbl[i3] = x2

static void lxs_increment_state(mc_ranlxs_state *state)
{
    int k, kmax, m;
    double x, y1, y2, y3;
    float *xflt = state->xflt;
    double *xdbl = state->xdbl;
    double *ydbl = state->ydbl;
    double carry = state->carry;
    unsigned int ir = state->ir;
    unsigned int jr = state->jr;

    for (k = 0; ir > 0; ++k) {
        y1 = xdbl[jr] - xdbl[ir];
        y2 = y1 - carry;
        if (y2 < 0) {
            carry = lxs_one_bit;
            y2 += 1;
        } else {
            carry = 0;
        }
        xdbl[ir] = y2;
        ir = lxs_next[ir];
        jr = lxs_next[jr];
    }

    kmax = (int)state->pr - 12;

    for (; k <= kmax; k += 12) {
        y1 = xdbl[7] - xdbl[0];
        y1 -= carry;

        LXS_STEP(y2, y1, 8, 1, 0);
        LXS_STEP(y3, y2, 9, 2, 1);
        LXS_STEP(y1, y3, 10, 3, 2);
        LXS_STEP(y2, y1, 11, 4, 3);
        LXS_STEP(y3, y2, 0, 5, 4);
        LXS_STEP(y1, y3, 1, 6, 5);
        LXS_STEP(y2, y1, 2, 7, 6);
        LXS_STEP(y3, y2, 3, 8, 7);
        LXS_STEP(y1, y3, 4, 9, 8);
        LXS_STEP(y2, y1, 5, 10, 9);
        LXS_STEP(y3, y2, 6, 11, 10);

        if (y3 < 0) {
            carry = lxs_one_bit;
            y3 += 1;
        } else {
            carry = 0;
        }
        xdbl[11] = y3;
    }

    kmax = (int)state->pr;

    for (; k < kmax; ++k) {
        y1 = xdbl[jr] - xdbl[ir];
        y2 = y1 - carry;
        if (y2 < 0) {
            carry = lxs_one_bit;
            y2 += 1;
        } else {
            carry = 0;
        }
        xdbl[ir] = y2;
        ydbl[ir] = y2 + lxs_shift;
        ir = lxs_next[ir];
        jr = lxs_next[jr];
    }

    ydbl[ir] = xdbl[ir] + lxs_shift;

    for (k = lxs_next[ir]; k > 0;) {
        ydbl[k] = xdbl[k] + lxs_shift;
        k = lxs_next[k];
    }

    for (k = 0, m = 0; k < 12; ++k) {
        x = xdbl[k];
        y2 = ydbl[k] - lxs_shift;
        if (y2 > x) y2 -= lxs_sone_bit;
        y1 = (x - y2) * lxs_sbase;

        xflt[m++] = (float)y1;
        xflt[m++] = (float)y2;
    }

    state->ir = ir;
    state->is = 2 * ir;
    state->is_old = 2 * ir;
    state->jr = jr;
    state->carry = carry;
}

double mc_ranlxs_get_double(mc_ranlxs_state *state)
{
    const unsigned int is = (unsigned int)lxs_snext[state->is];
    state->is = is;
    if (is == state->is_old) lxs_increment_state(state);
    return state->xflt[state->is];
}

unsigned long mc_ranlxs_get(mc_ranlxs_state *state)
{
    return (unsigned long)(mc_ranlxs_get_double(state) * 16777216.0);
}

void mc_ranlxs_set(mc_ranlxs_state *state, unsigned long s, unsigned int luxury)
{
    int ibit, jbit, i, k, m, xbit[31];
    double x, y;
    long int seed;

    if (s == 0) s = 1; /* GSL: default seed is 1 */
    seed = (long int)s;
    i = (int)(seed & 0x7FFFFFFFUL);

    for (k = 0; k < 31; ++k) {
        xbit[k] = i % 2;
        i /= 2;
    }

    ibit = 0;
    jbit = 18;

    for (k = 0; k < 12; ++k) {
        x = 0;
        for (m = 1; m <= 48; ++m) {
            y = (double)xbit[ibit];
            x += x + y;
            xbit[ibit] = (xbit[ibit] + xbit[jbit]) % 2;
            ibit = (ibit + 1) % 31;
            jbit = (jbit + 1) % 31;
        }
        state->xdbl[k] = lxs_one_bit * x;
    }

    state->carry = 0;
    state->ir = 0;
    state->jr = 7;
    state->is = 23;
    state->is_old = 0;
    state->pr = luxury;
}

/* ======================================================================== */
/* Philox4x32-10 (Salmon et al. 2011, Random123)                             */
/* ======================================================================== */
void mc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n1 = lo1;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        uint32_t n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void mc_philox_doubles(const uint32_t ctr[4], const uint32_t key[2], double out[2])
{
    uint32_t r[4];
    mc_philox4x32_10(ctr, key, r);
    uint64_t a = (((uint64_t)r[0] << 32) | r[1]) >> 12; /* 52 bits: a + 0.5 is exact */
    uint64_t b = (((uint64_t)r[2] << 32) | r[3]) >> 12;
    out[0] = ((double)a + 0.5) * (1.0 / 4503599627370496.0);
    out[1] = ((double)b + 0.5) * (1.0 / 4503599627370496.0);
}

/* ======================================================================== */
/* RNG handle                                                                */
/* ======================================================================== */
static inline double rng_note(mc_rng *r, double u)
{
    r->ndraws++;
    if (r->tee && r->tee_n < r->tee_cap) r->tee[r->tee_n] = u;
    if (r->tee) r->tee_n++;
    return u;
}

static double lxs_uniform(mc_rng *r) { return rng_note(r, mc_ranlxs_get_double(&r->lxs)); }
static double lxs_uniform_pos(mc_rng *r)
{
    double x;
    do {
        x = lxs_uniform(r);
    } while (x == 0);
    return x;
}
static unsigned long lxs_get(mc_rng *r) { return mc_ranlxs_get(&r->lxs); }
static void lxs_set(mc_rng *r, unsigned long s) { mc_ranlxs_set(&r->lxs, s, MC_RANLXS0_LUXURY); }

void mc_rng_init_ranlxs0(mc_rng *r, unsigned long seed)
{
    memset(r, 0, sizeof(*r));
    r->kind = MC_RNG_RANLXS0;
    r->uniform = lxs_uniform;
    r->uniform_pos = lxs_uniform_pos;
    r->get = lxs_get;
    r->set = lxs_set;
    mc_ranlxs_set(&r->lxs, seed, MC_RANLXS0_LUXURY);
}

static double replay_uniform(mc_rng *r)
{
    if (r->replay_pos >= r->replay_n) {
        /* running dry is a harness error; make it loud */
        abort();
    }
    return rng_note(r, r->replay[r->replay_pos++]);
}
static double replay_uniform_pos(mc_rng *r)
{
    double x;
    do {
        x = replay_uniform(r);
    } while (x == 0);
    return x;
}
static unsigned long replay_get(mc_rng *r) { return (unsigned long)(replay_uniform(r) * 16777216.0); }
static void replay_set(mc_rng *r, unsigned long s) { (void)r; (void)s; }

void mc_rng_init_replay(mc_rng *r, const double *buf, size_t n)
{
    memset(r, 0, sizeof(*r));
    r->kind = MC_RNG_REPLAY;
    r->uniform = replay_uniform;
    r->uniform_pos = replay_uniform_pos;
    r->get = replay_get;
    r->set = replay_set;
    r->replay = buf;
    r->replay_n = n;
}

static double philox_uniform(mc_rng *r)
{
    uint32_t ctr[4];
    double d[2];
    if (r->hint_stream == 0) {
        ctr[0] = r->hint_slot;
        ctr[1] = (uint32_t)r->hint_iter;
        ctr[2] = (uint32_t)(r->hint_iter >> 32);
        ctr[3] = 0u;
        mc_philox_doubles(ctr, r->key, d);
        return rng_note(r, d[0]);
    }
    if (r->hint_stream >= 3) {
        /* keyed side streams, counter (draw / 2, iter_lo, slot, stream + 8 * iter_hi):
         *   3  Monte Carlo integral of an out-of-table hot cross section (slot = photon slot, iter = loop iteration)
         *   4  Poisson count of a cell in photonEmitCyclosynch (slot = rank of the cell among the selected ones,
         *      iter = emission call << 32 | weight-search pass)
         *   5  direction / azimuth draws of an emitted photon (slot = index of the photon, iter = emission call << 32) */
        ctr[0] = (uint32_t)(r->hint_draw >> 1);
        ctr[1] = (uint32_t)r->hint_iter;
        ctr[2] = r->hint_slot;
        ctr[3] = r->hint_stream + ((uint32_t)(r->hint_iter >> 32) << 3);
        mc_philox_doubles(ctr, r->key, d);
        double v = d[r->hint_draw & 1u];
        r->hint_draw++;
        return rng_note(r, v);
    }
    ctr[0] = (uint32_t)(r->hint_draw >> 1);
    ctr[1] = (uint32_t)r->hint_iter;
    ctr[2] = (uint32_t)(r->hint_iter >> 32);
    ctr[3] = 1u;
    mc_philox_doubles(ctr, r->key, d);
    double u = d[r->hint_draw & 1u];
    r->hint_draw++;
    return rng_note(r, u);
}
static unsigned long philox_get(mc_rng *r) { return (unsigned long)(philox_uniform(r) * 16777216.0); }
static void philox_set(mc_rng *r, unsigned long s) { (void)r; (void)s; }

void mc_rng_init_philox(mc_rng *r, uint64_t seed, uint32_t shard)
{
    memset(r, 0, sizeof(*r));
    r->kind = MC_RNG_PHILOX;
    r->uniform = philox_uniform;
    r->uniform_pos = philox_uniform; /* already strictly inside (0,1) */
    r->get = philox_get;
    r->set = philox_set;
    r->key[0] = (uint32_t)seed ^ 0x4D435261u; /* 'MCRa' */
    r->key[1] = (uint32_t)(seed >> 32) ^ shard;
}

void mc_rng_set_tee(mc_rng *r, double *buf, size_t cap)
{
    r->tee = buf;
    r->tee_cap = cap;
    r->tee_n = 0;
}

void mc_rng_hint_mfp(mc_rng *r, uint64_t iter, uint32_t slot)
{
    r->hint_stream = 0;
    r->hint_iter = iter;
    r->hint_slot = slot;
}

void mc_rng_hint_keyed(mc_rng *r, uint32_t stream, uint32_t slot, uint64_t iter)
{
    r->hint_stream = stream;
    r->hint_slot = slot;
    r->hint_iter = iter;
    r->hint_draw = 0;
}

void mc_rng_hint_event(mc_rng *r, uint64_t event)
{
    r->hint_stream = 1;
    r->hint_iter = event;
    r->hint_draw = 0;
}

/* gsl_ran_gaussian: polar (Box-Muller, Knuth v2 3rd ed p122) */
double mc_ran_gaussian(mc_rng *r, double sigma)
{
    double x, y, r2;
    do {
        x = -1 + 2 * r->uniform_pos(r);
        y = -1 + 2 * r->uniform_pos(r);
        r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0);
    return sigma * y * sqrt(-2.0 * log(r2) / r2);
}

/* Poisson deviate.  GSL's version recurses through gamma and binomial
 * deviates for mu>10; this statistical restatement uses multiplication for
 * small mu and the PTRS transformed-rejection method (Hoermann 1993) above.
 * Parity against a real GSL build is therefore unpinned for Poisson draws
 * (they are used only by photon injection / cyclo-synchrotron emission). */
unsigned int mc_ran_poisson(mc_rng *r, double mu)
{
    if (!(mu > 0)) return 0;
    if (mu < 10.0) {
        double emu = exp(-mu), prod = 1.0;
        unsigned int k = 0;
        do {
            prod *= r->uniform(r);
            k++;
        } while (prod > emu);
        return k - 1;
    }
    {
        double smu = sqrt(mu);
        double b = 0.931 + 2.53 * smu;
        double a = -0.059 + 0.02483 * b;
        double inv_alpha = 1.1239 + 1.1328 / (b - 3.4);
        double v_r = 0.9277 - 3.6224 / (b - 2.0);
        for (;;) {
            double U = r->uniform(r) - 0.5;
            double V = r->uniform_pos(r);
            double us = 0.5 - fabs(U);
            double kf = floor((2.0 * a / us + b) * U + mu + 0.43);
            if (us >= 0.07 && V <= v_r) return (unsigned int)kf;
            if (kf < 0 || (us < 0.013 && V > us)) continue;
            if (log(V) + log(inv_alpha) - log(a / (us * us) + b) <=
                -mu + kf * log(mu) - lgamma(kf + 1.0))
                return (unsigned int)kf;
        }
    }
}

/* ======================================================================== */
/* Modified Bessel function K_n(x), integer n >= 0                            */
/* ======================================================================== */
static void bessel_K0K1(double x, double *k0, double *k1)
{
    const double EULER = 0.57721566490153286060651209008240243;
    if (x <= 2.0) {
        /* ascending series (Abramowitz & Stegun 9.6.11, 9.6.13) */
        double q = 0.25 * x * x;
        double lh = log(0.5 * x);
        /* I0 = sum q^k/(k!)^2 ;  K0 = -(lh+g) I0 + sum q^k/(k!)^2 H_k */
        double term = 1.0, i0 = 1.0, s0 = 0.0, h = 0.0;
        /* I1 = (x/2) sum q^k/(k!(k+1)!) ; K1 series with psi(k+1)+psi(k+2) */
        double term1 = 1.0, i1 = 1.0;
        double psi_sum = (-EULER) + (1.0 - EULER); /* psi(1)+psi(2) */
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
        *k0 = -(lh + EULER) * i0 + s0;
        *k1 = 1.0 / x + lh * (0.5 * x) * i1 - 0.25 * x * s1;
    } else {
        /* Steed's algorithm for Temme's continued fraction CF2 (nu = 0) */
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
        double rk0 = sqrt(M_PI / (2.0 * x)) * exp(-x) / s;
        *k0 = rk0;
        *k1 = rk0 * (x + 0.5 - h) / x;
    }
}

double mc_bessel_Kn(int n, double x)
{
    double k0, k1;
    if (n < 0) n = -n;
    if (!(x > 0.0)) return NAN;
    bessel_K0K1(x, &k0, &k1);
    if (n == 0) return k0;
    if (n == 1) return k1;
    double km = k0, k = k1;
    for (int j = 1; j < n; ++j) {
        double kp = km + (2.0 * (double)j / x) * k;
        km = k;
        k = kp;
    }
    return k;
}

/* ======================================================================== */
/* tiny BLAS (gsl cblas reference kernels)                                   */
/* ======================================================================== */
double mc_dnrm2(int n, const double *X)
{
    double scale = 0.0, ssq = 1.0;
    if (n <= 0) return 0;
    if (n == 1) return fabs(X[0]);
    for (int i = 0; i < n; i++) {
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

double mc_ddot(int n, const double *x, const double *y)
{
    double r = 0.0;
    for (int i = 0; i < n; i++) r += x[i] * y[i];
    return r;
}

void mc_dgemv(int n, const double *A, const double *x, double *y)
{
    /* beta == 0: y := 0;  then y[i] += alpha * sum_j x[j]*A[i][j] */
    for (int i = 0; i < n; i++) y[i] = 0.0;
    for (int i = 0; i < n; i++) {
        double temp = 0.0;
        for (int j = 0; j < n; j++) temp += x[j] * A[n * i + j];
        y[i] += 1.0 * temp;
    }
}

/* ======================================================================== */
/* bilinear interpolation (gsl interp2d.c + bilinear.c)                      */
/* ======================================================================== */
size_t mc_interp_bsearch(const double *xa, double x, size_t lo, size_t hi)
{
    size_t ilo = lo, ihi = hi;
    while (ihi > ilo + 1) {
        size_t i = (ihi + ilo) / 2;
        if (xa[i] > x)
            ihi = i;
        else
            ilo = i;
    }
    return ilo;
}

int mc_bilinear_eval(const double *xa, const double *ya, const double *za, size_t nx,
                     size_t ny, double x, double y, double *z)
{
    if (x < xa[0] || x > xa[nx - 1]) return 1; /* GSL_EDOM */
    if (y < ya[0] || y > ya[ny - 1]) return 1;
    size_t xi = mc_interp_bsearch(xa, x, 0, nx - 1);
    size_t yi = mc_interp_bsearch(ya, y, 0, ny - 1);
    double xmin = xa[xi], xmax = xa[xi + 1];
    double ymin = ya[yi], ymax = ya[yi + 1];
    double zminmin = za[yi * nx + xi];
    double zminmax = za[(yi + 1) * nx + xi];
    double zmaxmin = za[yi * nx + xi + 1];
    double zmaxmax = za[(yi + 1) * nx + xi + 1];
    double dx = xmax - xmin, dy = ymax - ymin;
    double t = (x - xmin) / dx;
    double u = (y - ymin) / dy;
    *z = (1. - t) * (1. - u) * zminmin + t * (1. - u) * zmaxmin + (1. - t) * u * zminmax +
         t * u * zmaxmax;
    return 0;
}

/* ======================================================================== */
/* gsl_monte_plain_integrate                                                 */
/* ======================================================================== */
void mc_monte_plain(mc_monte_fn f, void *params, const double *xl, const double *xu,
                    size_t dim, size_t calls, mc_rng *r, double *result, double *abserr)
{
    double vol = 1, m = 0, q = 0;
    double x[8];
    for (size_t i = 0; i < dim; i++) vol *= xu[i] - xl[i];
    for (size_t n = 0; n < calls; n++) {
        for (size_t i = 0; i < dim; i++) x[i] = xl[i] + r->uniform_pos(r) * (xu[i] - xl[i]);
        {
            double fval = f(x, dim, params);
            double d = fval - m;
            m += d / (n + 1.0);
            q += d * d * (n / (n + 1.0));
        }
    }
    *result = vol * m;
    if (calls < 2)
        *abserr = INFINITY;
    else
        *abserr = vol * sqrt(q / (calls * (calls - 1.0)));
}

/* ======================================================================== */
/* adaptive Gauss-Kronrod (7,15) quadrature with global bisection.           */
/* Stands in for gsl_integration_qags (no epsilon extrapolation); the only    */
/* call site integrates a smooth black-body tail at epsrel = 1e-2             */
/* (mc_cyclosynch.c:1285), far looser than either rule's accuracy.            */
/* ======================================================================== */
static void gk15(mc_quad_fn f, void *p, double a, double b, double *res, double *err)
{
    static const double xgk[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                                  0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                                  0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                                  0.207784955007898467600689403773245, 0.000000000000000000000000000000000};
    static const double wgk[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                                  0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                                  0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                                  0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
    static const double wg[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                                 0.381830050505118944950369775488975, 0.417959183673469387755102040816327};
    double c = 0.5 * (a + b), h = 0.5 * (b - a);
    double fc = f(c, p);
    double rk = fc * wgk[7], rg = fc * wg[3];
    for (int j = 0; j < 7; ++j) {
        double dx = h * xgk[j];
        double f1 = f(c - dx, p), f2 = f(c + dx, p);
        rk += wgk[j] * (f1 + f2);
        if (j & 1) rg += wg[j / 2] * (f1 + f2);
    }
    *res = rk * h;
    *err = fabs((rk - rg) * h);
}

int mc_integrate_adaptive(mc_quad_fn f, void *params, double a, double b, double epsabs,
                          double epsrel, size_t limit, double *result, double *abserr)
{
    typedef struct { double a, b, r, e; } seg;
    seg *s = (seg *)malloc(sizeof(seg) * (limit ? limit : 1));
    size_t n = 1;
    int status = 0;
    s[0].a = a; s[0].b = b;
    gk15(f, params, a, b, &s[0].r, &s[0].e);
    for (;;) {
        double tot = 0, err = 0;
        size_t worst = 0;
        for (size_t i = 0; i < n; ++i) {
            tot += s[i].r;
            err += s[i].e;
            if (s[i].e > s[worst].e) worst = i;
        }
        double tol = fmax(epsabs, epsrel * fabs(tot));
        if (err <= tol || n >= limit) {
            *result = tot;
            *abserr = err;
            status = (err <= tol) ? 0 : 1;
            break;
        }
        {
            seg w = s[worst];
            double mid = 0.5 * (w.a + w.b);
            s[worst].a = w.a; s[worst].b = mid;
            gk15(f, params, w.a, mid, &s[worst].r, &s[worst].e);
            s[n].a = mid; s[n].b = w.b;
            gk15(f, params, mid, w.b, &s[n].r, &s[n].e);
            n++;
        }
    }
    free(s);
    return status;
}

/* ---- gsl_histogram: uniform ranges and find ------------------------------------------------ */
void mc_hist_uniform_ranges(double *range, size_t n, double xmin, double xmax)
{
    size_t i;
    for (i = 0; i <= n; i++) {
        double f1 = ((double)(n - i) / (double)n);
        double f2 = ((double)i / (double)n);
        range[i] = f1 * xmin + f2 * xmax;
    }
}

int mc_hist_find(size_t n, const double *range, double x, size_t *i)
{
    size_t i_linear, lower, upper, mid;
    if (x < range[0]) return 1;
    if (x >= range[n]) return 1;
    {
        double u = (x - range[0]) / (range[n] - range[0]);
        i_linear = (size_t)(u * n);
    }
    if (i_linear < n && x >= range[i_linear] && x < range[i_linear + 1]) {
        *i = i_linear;
        return 0;
    }
    upper = n;
    lower = 0;
    while (upper - lower > 1) {
        mid = (upper + lower) / 2;
        if (x >= range[mid])
            lower = mid;
        else
            upper = mid;
    }
    *i = lower;
    return 0;
}
