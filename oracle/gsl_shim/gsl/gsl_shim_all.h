/*
 * gsl_shim_all.h -- TEST INFRASTRUCTURE (oracle).
 *
 * Minimal stand-in for the subset of the GNU Scientific Library API that the
 * MCRaT hot-path translation units reference (SURVEY.md section 8c lists the
 * symbols).  GSL itself is absent from this image; this header lets the
 * reference's own, unmodified sources compile into oracle/_ref.  Every numeric
 * routine forwards to oracle/mc_mathlib.c, which restates the published GSL
 * algorithms (RANLXS, polar Box-Muller, reference CBLAS kernels, bilinear
 * interp2d, plain Monte Carlo).  Containers follow GSL's documented layout
 * semantics (row-major matrix with tda, strided vector views).
 */
#ifndef GSL_SHIM_ALL_H
#define GSL_SHIM_ALL_H

#include <float.h>  /* real gsl_machine.h pulls in <limits.h> and <float.h> */
#include <limits.h>
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mc_mathlib.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- errno ---------------------------------------------------------------- */
enum { GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_EDOM = 1, GSL_EINVAL = 4 };
typedef void gsl_error_handler_t(const char *reason, const char *file, int line, int gsl_errno);
gsl_error_handler_t *gsl_set_error_handler_off(void);
gsl_error_handler_t *gsl_set_error_handler(gsl_error_handler_t *h);
const char *gsl_strerror(const int gsl_errno);

/* ---- math ------------------------------------------------------------------ */
#ifndef M_PI
#define M_PI 3.14159265358979323846264338328
#endif
typedef struct {
    double (*function)(double x, void *params);
    void *params;
} gsl_function;

/* ---- vector / matrix ------------------------------------------------------ */
typedef struct {
    size_t size;
    size_t stride;
    double *data;
    void *block;
    int owner;
} gsl_vector;
typedef struct {
    gsl_vector vector;
} gsl_vector_view;
typedef struct {
    size_t size1, size2, tda;
    double *data;
    void *block;
    int owner;
} gsl_matrix;

gsl_vector *gsl_vector_alloc(size_t n);
gsl_vector *gsl_vector_calloc(size_t n);
void gsl_vector_free(gsl_vector *v);
gsl_vector_view gsl_vector_view_array(double *base, size_t n);
static inline double gsl_vector_get(const gsl_vector *v, size_t i) { return v->data[i * v->stride]; }
static inline void gsl_vector_set(gsl_vector *v, size_t i, double x) { v->data[i * v->stride] = x; }
static inline double *gsl_vector_ptr(gsl_vector *v, size_t i) { return v->data + i * v->stride; }
int gsl_vector_add(gsl_vector *a, const gsl_vector *b);
int gsl_vector_sub(gsl_vector *a, const gsl_vector *b);
int gsl_vector_fprintf(FILE *stream, const gsl_vector *v, const char *format);

gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2);
gsl_matrix *gsl_matrix_calloc(size_t n1, size_t n2);
void gsl_matrix_free(gsl_matrix *m);
static inline double gsl_matrix_get(const gsl_matrix *m, size_t i, size_t j) { return m->data[i * m->tda + j]; }
static inline void gsl_matrix_set(gsl_matrix *m, size_t i, size_t j, double x) { m->data[i * m->tda + j] = x; }
void gsl_matrix_set_all(gsl_matrix *m, double x);
int gsl_matrix_scale(gsl_matrix *m, double x);

/* ---- blas ------------------------------------------------------------------ */
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE_t;
double gsl_blas_dnrm2(const gsl_vector *x);
int gsl_blas_ddot(const gsl_vector *x, const gsl_vector *y, double *result);
int gsl_blas_dgemv(CBLAS_TRANSPOSE_t TransA, double alpha, const gsl_matrix *A, const gsl_vector *x,
                   double beta, gsl_vector *y);

/* ---- rng / randist ---------------------------------------------------------- */
typedef struct {
    const char *name;
    unsigned int luxury;
} gsl_rng_type;
typedef struct {
    const gsl_rng_type *type;
    mc_rng impl;
} gsl_rng;
extern const gsl_rng_type *gsl_rng_ranlxs0;
extern const gsl_rng_type *gsl_rng_default;
const gsl_rng_type *gsl_rng_env_setup(void);
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_free(gsl_rng *r);
void gsl_rng_set(gsl_rng *r, unsigned long seed);
unsigned long gsl_rng_get(gsl_rng *r);
double gsl_rng_uniform(gsl_rng *r);
double gsl_rng_uniform_pos(gsl_rng *r);
double gsl_ran_gaussian(gsl_rng *r, double sigma);
unsigned int gsl_ran_poisson(gsl_rng *r, double mu);
/* harness hooks (not GSL): replace the stream behind a gsl_rng */
void gsl_shim_rng_use_replay(gsl_rng *r, const double *buf, size_t n);
void gsl_shim_rng_set_tee(gsl_rng *r, double *buf, size_t cap);
size_t gsl_shim_rng_tee_count(const gsl_rng *r);
unsigned long long gsl_shim_rng_draws(const gsl_rng *r);

/* ---- special functions ------------------------------------------------------ */
double gsl_sf_bessel_Kn(const int n, const double x);

/* ---- interp2d / spline2d ---------------------------------------------------- */
typedef struct {
    const char *name;
} gsl_interp2d_type;
extern const gsl_interp2d_type *gsl_interp2d_bilinear;
typedef struct {
    size_t cache;
} gsl_interp_accel;
typedef struct {
    size_t nx, ny;
    double *xarr, *yarr, *zarr;
} gsl_spline2d;
gsl_interp_accel *gsl_interp_accel_alloc(void);
void gsl_interp_accel_free(gsl_interp_accel *a);
gsl_spline2d *gsl_spline2d_alloc(const gsl_interp2d_type *T, size_t xsize, size_t ysize);
int gsl_spline2d_init(gsl_spline2d *s, const double xa[], const double ya[], const double za[], size_t xsize,
                      size_t ysize);
void gsl_spline2d_free(gsl_spline2d *s);
int gsl_spline2d_eval_e(const gsl_spline2d *s, const double x, const double y, gsl_interp_accel *xa,
                        gsl_interp_accel *ya, double *z);

/* ---- monte ------------------------------------------------------------------- */
typedef struct {
    double (*f)(double *x_array, size_t dim, void *params);
    size_t dim;
    void *params;
} gsl_monte_function;
typedef struct {
    size_t dim;
    double *x;
} gsl_monte_plain_state;
gsl_monte_plain_state *gsl_monte_plain_alloc(size_t dim);
void gsl_monte_plain_free(gsl_monte_plain_state *s);
int gsl_monte_plain_integrate(const gsl_monte_function *f, const double xl[], const double xu[], const size_t dim,
                              const size_t calls, gsl_rng *r, gsl_monte_plain_state *state, double *result,
                              double *abserr);

/* ---- integration -------------------------------------------------------------- */
typedef struct {
    size_t limit;
} gsl_integration_workspace;
gsl_integration_workspace *gsl_integration_workspace_alloc(const size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace *w);
int gsl_integration_qags(const gsl_function *f, double a, double b, double epsabs, double epsrel, size_t limit,
                         gsl_integration_workspace *workspace, double *result, double *abserr);

/* ---- histogram2d ---------------------------------------------------------------- */
typedef struct {
    size_t nx, ny;
    double *xrange;
    double *yrange;
    double *bin;
} gsl_histogram2d;
gsl_histogram2d *gsl_histogram2d_alloc(const size_t nx, const size_t ny);
void gsl_histogram2d_free(gsl_histogram2d *h);
int gsl_histogram2d_set_ranges_uniform(gsl_histogram2d *h, double xmin, double xmax, double ymin, double ymax);
int gsl_histogram2d_increment(gsl_histogram2d *h, double x, double y);
int gsl_histogram2d_find(const gsl_histogram2d *h, const double x, const double y, size_t *i, size_t *j);
double gsl_histogram2d_get(const gsl_histogram2d *h, const size_t i, const size_t j);
int gsl_histogram2d_fprintf(FILE *stream, const gsl_histogram2d *h, const char *range_format,
                            const char *bin_format);

#ifdef __cplusplus
}
#endif
#endif
