/*
 * mc_mathlib.h -- TEST INFRASTRUCTURE (oracle). Not part of the product path.
 *
 * Restatement of the GSL numerics the MCRaT hot path depends on.  GSL is an
 * un-vendored, un-pinned dependency of the reference (Makefile:7 `-lgsl
 * -lgslcblas`; a site path `gsl-2.4` appears in Makefile_supercomputer:8) and is
 * absent from this image, so its published algorithms are restated here:
 *
 *   - RANLXS luxury level 0 (`gsl_rng_ranlxs0`, M. Luescher's ranlxs v2.1 as
 *     shipped in GSL rng/ranlxs.c); pinned by GSL's own rng/test.c known-answer
 *     values (10000th output for seed 1), see tests/test_mathlib.py.
 *   - `gsl_ran_gaussian` (polar Box-Muller), `gsl_ran_poisson` (statistical
 *     restatement only -- parity vs real GSL unpinned for Poisson).
 *   - `gsl_sf_bessel_Kn` (series for x<=2, Temme/Steed continued fraction
 *     above; checked against scipy.special.kn).
 *   - level-1/2 CBLAS as used on 3- and 4-vectors (dnrm2 scaled form, ddot,
 *     dgemv row-major NoTrans).
 *   - `gsl_interp2d_bilinear` evaluation with GSL_EDOM outside the grid.
 *   - `gsl_monte_plain_integrate`.
 *   - `gsl_integration_qags` replacement (adaptive Gauss-Kronrod 21).
 *
 * Both oracle/mcrat_oracle.c (the restatement of the reference) and
 * oracle/gsl_shim (the link-time stand-in used to compile the reference's own
 * sources into oracle/_ref) call these, so the two agree bit-for-bit wherever
 * they perform the same operations.
 */
#ifndef MC_MATHLIB_H
#define MC_MATHLIB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- RANLXS ------------------------------------------------------------ */
typedef struct {
    double xdbl[12], ydbl[12];
    double carry;
    float xflt[24];
    unsigned int ir, jr, is, is_old, pr;
} mc_ranlxs_state;

void mc_ranlxs_set(mc_ranlxs_state *st, unsigned long seed, unsigned int luxury);
double mc_ranlxs_get_double(mc_ranlxs_state *st); /* k/2^24 in [0,1) */
unsigned long mc_ranlxs_get(mc_ranlxs_state *st); /* k in [0,2^24) */

#define MC_RANLXS0_LUXURY 109u
#define MC_RANLXS1_LUXURY 202u
#define MC_RANLXS2_LUXURY 397u

/* ---- Philox4x32-10 (counter-based; same function as the device RNG) ----- */
void mc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* two doubles in (0,1) from one Philox block: ((52-bit int) + 0.5) * 2^-52, exact, strictly inside (0,1) */
void mc_philox_doubles(const uint32_t ctr[4], const uint32_t key[2], double out[2]);

/* ---- generic RNG handle used by the oracle ------------------------------ */
typedef struct mc_rng mc_rng;
struct mc_rng {
    double (*uniform)(mc_rng *);     /* [0,1)  (gsl_rng_uniform)     */
    double (*uniform_pos)(mc_rng *); /* (0,1)  (gsl_rng_uniform_pos) */
    unsigned long (*get)(mc_rng *);  /* gsl_rng_get                  */
    void (*set)(mc_rng *, unsigned long);
    int kind;                        /* MC_RNG_* */
    /* sequential generator */
    mc_ranlxs_state lxs;
    /* replay buffer */
    const double *replay;
    size_t replay_n, replay_pos;
    /* optional tee of every uniform handed out */
    double *tee;
    size_t tee_cap, tee_n;
    /* keyed Philox streams: hints set by the oracle before it draws */
    uint32_t key[2];
    uint64_t hint_iter;   /* while-loop iteration / event number */
    uint32_t hint_slot;   /* photon slot (stream 0 only)          */
    uint32_t hint_stream; /* 0 = free-path draw, 1 = event draws  */
    uint64_t hint_draw;   /* running draw index inside an event   */
    unsigned long long ndraws;
};
enum { MC_RNG_RANLXS0 = 0, MC_RNG_REPLAY = 1, MC_RNG_PHILOX = 2 };

void mc_rng_init_ranlxs0(mc_rng *r, unsigned long seed);
void mc_rng_init_replay(mc_rng *r, const double *buf, size_t n);
void mc_rng_init_philox(mc_rng *r, uint64_t seed, uint32_t shard);
void mc_rng_set_tee(mc_rng *r, double *buf, size_t cap);
/* stream selection for keyed generators (no-ops for sequential ones) */
void mc_rng_hint_mfp(mc_rng *r, uint64_t iter, uint32_t slot);
void mc_rng_hint_keyed(mc_rng *r, uint32_t stream, uint32_t slot, uint64_t iter);
void mc_rng_hint_event(mc_rng *r, uint64_t event);

double mc_ran_gaussian(mc_rng *r, double sigma);
unsigned int mc_ran_poisson(mc_rng *r, double mu);

/* ---- special functions --------------------------------------------------- */
double mc_bessel_Kn(int n, double x);

/* ---- tiny BLAS ------------------------------------------------------------ */
double mc_dnrm2(int n, const double *x);
double mc_ddot(int n, const double *x, const double *y);
/* y = A x, A row-major n x n (cblas_dgemv NoTrans, alpha=1, beta=0) */
void mc_dgemv(int n, const double *A, const double *x, double *y);

/* ---- bilinear interpolation ---------------------------------------------- */
/* za[j*nx + i]; returns 0 on success, 1 (GSL_EDOM) outside the grid */
int mc_bilinear_eval(const double *xa, const double *ya, const double *za,
                     size_t nx, size_t ny, double x, double y, double *z);
size_t mc_interp_bsearch(const double *xa, double x, size_t lo, size_t hi);

/* ---- plain Monte Carlo integration ---------------------------------------- */
typedef double (*mc_monte_fn)(double *x, size_t dim, void *params);
void mc_monte_plain(mc_monte_fn f, void *params, const double *xl, const double *xu,
                    size_t dim, size_t calls, mc_rng *r, double *result, double *abserr);

/* ---- adaptive quadrature ---------------------------------------------------- */
typedef double (*mc_quad_fn)(double x, void *params);
int mc_integrate_adaptive(mc_quad_fn f, void *params, double a, double b, double epsabs,
                          double epsrel, size_t limit, double *result, double *abserr);

#ifdef __cplusplus
}
#endif
/* gsl_histogram2d_set_ranges_uniform's edges (histogram/init.c make_uniform) and gsl_histogram find()
 * (histogram/find.c: linear guess, then bisection); find returns 0 and writes *i on success, 1 outside the range */
void mc_hist_uniform_ranges(double *range, size_t n, double xmin, double xmax);
int mc_hist_find(size_t n, const double *range, double x, size_t *i);

#endif
