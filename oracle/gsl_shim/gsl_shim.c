/*
 * gsl_shim.c -- TEST INFRASTRUCTURE (oracle).
 * Implementation of the GSL API subset declared in gsl/gsl_shim_all.h on top of
 * oracle/mc_mathlib.c.  Linked only into oracle/_ref (the reference's own
 * sources compiled for validation / CPU-baseline timing).
 */
#include "gsl/gsl_shim_all.h"

/* ---- errno ---------------------------------------------------------------- */
static gsl_error_handler_t *g_handler = NULL;
gsl_error_handler_t *gsl_set_error_handler_off(void)
{
    gsl_error_handler_t *old = g_handler;
    g_handler = NULL;
    return old;
}
gsl_error_handler_t *gsl_set_error_handler(gsl_error_handler_t *h)
{
    gsl_error_handler_t *old = g_handler;
    g_handler = h;
    return old;
}
const char *gsl_strerror(const int e)
{
    switch (e) {
    case GSL_SUCCESS: return "success";
    case GSL_FAILURE: return "failure";
    case GSL_EDOM: return "input domain error";
    case GSL_EINVAL: return "invalid argument supplied by user";
    default: return "unknown error code";
    }
}

/* ---- vector / matrix -------------------------------------------------------- */
static gsl_vector *vec_new(size_t n, int zero)
{
    gsl_vector *v = (gsl_vector *)malloc(sizeof(*v));
    v->size = n;
    v->stride = 1;
    v->data = (double *)(zero ? calloc(n ? n : 1, sizeof(double)) : malloc((n ? n : 1) * sizeof(double)));
    v->block = v->data;
    v->owner = 1;
    return v;
}
gsl_vector *gsl_vector_alloc(size_t n) { return vec_new(n, 0); }
gsl_vector *gsl_vector_calloc(size_t n) { return vec_new(n, 1); }
void gsl_vector_free(gsl_vector *v)
{
    if (!v) return;
    if (v->owner) free(v->data);
    free(v);
}
gsl_vector_view gsl_vector_view_array(double *base, size_t n)
{
    gsl_vector_view view;
    view.vector.size = n;
    view.vector.stride = 1;
    view.vector.data = base;
    view.vector.block = NULL;
    view.vector.owner = 0;
    return view;
}
int gsl_vector_add(gsl_vector *a, const gsl_vector *b)
{
    for (size_t i = 0; i < a->size; i++) a->data[i * a->stride] += b->data[i * b->stride];
    return GSL_SUCCESS;
}
int gsl_vector_sub(gsl_vector *a, const gsl_vector *b)
{
    for (size_t i = 0; i < a->size; i++) a->data[i * a->stride] -= b->data[i * b->stride];
    return GSL_SUCCESS;
}
int gsl_vector_fprintf(FILE *stream, const gsl_vector *v, const char *format)
{
    for (size_t i = 0; i < v->size; i++) {
        fprintf(stream, format, v->data[i * v->stride]);
        fputc('\n', stream);
    }
    return GSL_SUCCESS;
}

static gsl_matrix *mat_new(size_t n1, size_t n2, int zero)
{
    gsl_matrix *m = (gsl_matrix *)malloc(sizeof(*m));
    size_t n = n1 * n2;
    m->size1 = n1;
    m->size2 = n2;
    m->tda = n2;
    m->data = (double *)(zero ? calloc(n ? n : 1, sizeof(double)) : malloc((n ? n : 1) * sizeof(double)));
    m->block = m->data;
    m->owner = 1;
    return m;
}
gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2) { return mat_new(n1, n2, 0); }
gsl_matrix *gsl_matrix_calloc(size_t n1, size_t n2) { return mat_new(n1, n2, 1); }
void gsl_matrix_free(gsl_matrix *m)
{
    if (!m) return;
    if (m->owner) free(m->data);
    free(m);
}
void gsl_matrix_set_all(gsl_matrix *m, double x)
{
    for (size_t i = 0; i < m->size1; i++)
        for (size_t j = 0; j < m->size2; j++) m->data[i * m->tda + j] = x;
}
int gsl_matrix_scale(gsl_matrix *m, double x)
{
    for (size_t i = 0; i < m->size1; i++)
        for (size_t j = 0; j < m->size2; j++) m->data[i * m->tda + j] *= x;
    return GSL_SUCCESS;
}

/* ---- blas ------------------------------------------------------------------ */
double gsl_blas_dnrm2(const gsl_vector *x)
{
    double tmp[8];
    if (x->stride == 1) return mc_dnrm2((int)x->size, x->data);
    for (size_t i = 0; i < x->size && i < 8; i++) tmp[i] = x->data[i * x->stride];
    return mc_dnrm2((int)x->size, tmp);
}
int gsl_blas_ddot(const gsl_vector *x, const gsl_vector *y, double *result)
{
    double r = 0.0;
    for (size_t i = 0; i < x->size; i++) r += x->data[i * x->stride] * y->data[i * y->stride];
    *result = r;
    return GSL_SUCCESS;
}
int gsl_blas_dgemv(CBLAS_TRANSPOSE_t TransA, double alpha, const gsl_matrix *A, const gsl_vector *x, double beta,
                   gsl_vector *y)
{
    /* reference cblas_dgemv, row-major, NoTrans (the only form the hot path uses) */
    const size_t M = A->size1, N = A->size2;
    if (TransA != CblasNoTrans) abort();
    if (beta == 0.0) {
        for (size_t i = 0; i < M; i++) y->data[i * y->stride] = 0.0;
    } else if (beta != 1.0) {
        for (size_t i = 0; i < M; i++) y->data[i * y->stride] *= beta;
    }
    if (alpha == 0.0) return GSL_SUCCESS;
    for (size_t i = 0; i < M; i++) {
        double temp = 0.0;
        for (size_t j = 0; j < N; j++) temp += x->data[j * x->stride] * A->data[A->tda * i + j];
        y->data[i * y->stride] += alpha * temp;
    }
    return GSL_SUCCESS;
}

/* ---- rng ---------------------------------------------------------------------- */
static const gsl_rng_type ranlxs0_type = {"ranlxs0", MC_RANLXS0_LUXURY};
const gsl_rng_type *gsl_rng_ranlxs0 = &ranlxs0_type;
const gsl_rng_type *gsl_rng_default = &ranlxs0_type;
const gsl_rng_type *gsl_rng_env_setup(void) { return gsl_rng_default; }
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T)
{
    gsl_rng *r = (gsl_rng *)malloc(sizeof(*r));
    r->type = T;
    mc_rng_init_ranlxs0(&r->impl, 0); /* gsl_rng_default_seed == 0 -> 1 */
    return r;
}
void gsl_rng_free(gsl_rng *r) { free(r); }
void gsl_rng_set(gsl_rng *r, unsigned long seed) { r->impl.set(&r->impl, seed); }
unsigned long gsl_rng_get(gsl_rng *r) { return r->impl.get(&r->impl); }
double gsl_rng_uniform(gsl_rng *r) { return r->impl.uniform(&r->impl); }
double gsl_rng_uniform_pos(gsl_rng *r) { return r->impl.uniform_pos(&r->impl); }
double gsl_ran_gaussian(gsl_rng *r, double sigma) { return mc_ran_gaussian(&r->impl, sigma); }
unsigned int gsl_ran_poisson(gsl_rng *r, double mu) { return mc_ran_poisson(&r->impl, mu); }
void gsl_shim_rng_use_replay(gsl_rng *r, const double *buf, size_t n) { mc_rng_init_replay(&r->impl, buf, n); }
void gsl_shim_rng_set_tee(gsl_rng *r, double *buf, size_t cap) { mc_rng_set_tee(&r->impl, buf, cap); }
size_t gsl_shim_rng_tee_count(const gsl_rng *r) { return r->impl.tee_n; }
unsigned long long gsl_shim_rng_draws(const gsl_rng *r) { return r->impl.ndraws; }

/* ---- special functions --------------------------------------------------------- */
double gsl_sf_bessel_Kn(const int n, const double x) { return mc_bessel_Kn(n, x); }

/* ---- interp2d -------------------------------------------------------------------- */
static const gsl_interp2d_type bilinear_type = {"bilinear"};
const gsl_interp2d_type *gsl_interp2d_bilinear = &bilinear_type;
gsl_interp_accel *gsl_interp_accel_alloc(void) { return (gsl_interp_accel *)calloc(1, sizeof(gsl_interp_accel)); }
void gsl_interp_accel_free(gsl_interp_accel *a) { free(a); }
gsl_spline2d *gsl_spline2d_alloc(const gsl_interp2d_type *T, size_t xsize, size_t ysize)
{
    gsl_spline2d *s = (gsl_spline2d *)malloc(sizeof(*s));
    (void)T;
    s->nx = xsize;
    s->ny = ysize;
    s->xarr = (double *)malloc(xsize * sizeof(double));
    s->yarr = (double *)malloc(ysize * sizeof(double));
    s->zarr = (double *)malloc(xsize * ysize * sizeof(double));
    return s;
}
int gsl_spline2d_init(gsl_spline2d *s, const double xa[], const double ya[], const double za[], size_t xsize,
                      size_t ysize)
{
    memcpy(s->xarr, xa, xsize * sizeof(double));
    memcpy(s->yarr, ya, ysize * sizeof(double));
    memcpy(s->zarr, za, xsize * ysize * sizeof(double));
    return GSL_SUCCESS;
}
void gsl_spline2d_free(gsl_spline2d *s)
{
    if (!s) return;
    free(s->xarr);
    free(s->yarr);
    free(s->zarr);
    free(s);
}
int gsl_spline2d_eval_e(const gsl_spline2d *s, const double x, const double y, gsl_interp_accel *xa,
                        gsl_interp_accel *ya, double *z)
{
    (void)xa;
    (void)ya;
    return mc_bilinear_eval(s->xarr, s->yarr, s->zarr, s->nx, s->ny, x, y, z) ? GSL_EDOM : GSL_SUCCESS;
}

/* ---- monte ------------------------------------------------------------------------ */
gsl_monte_plain_state *gsl_monte_plain_alloc(size_t dim)
{
    gsl_monte_plain_state *s = (gsl_monte_plain_state *)malloc(sizeof(*s));
    s->dim = dim;
    s->x = (double *)malloc(dim * sizeof(double));
    return s;
}
void gsl_monte_plain_free(gsl_monte_plain_state *s)
{
    if (!s) return;
    free(s->x);
    free(s);
}
int gsl_monte_plain_integrate(const gsl_monte_function *f, const double xl[], const double xu[], const size_t dim,
                              const size_t calls, gsl_rng *r, gsl_monte_plain_state *state, double *result,
                              double *abserr)
{
    (void)state;
    mc_monte_plain(f->f, f->params, xl, xu, dim, calls, &r->impl, result, abserr);
    return GSL_SUCCESS;
}

/* ---- integration ---------------------------------------------------------------------- */
gsl_integration_workspace *gsl_integration_workspace_alloc(const size_t n)
{
    gsl_integration_workspace *w = (gsl_integration_workspace *)malloc(sizeof(*w));
    w->limit = n;
    return w;
}
void gsl_integration_workspace_free(gsl_integration_workspace *w) { free(w); }
int gsl_integration_qags(const gsl_function *f, double a, double b, double epsabs, double epsrel, size_t limit,
                         gsl_integration_workspace *workspace, double *result, double *abserr)
{
    (void)workspace;
    return mc_integrate_adaptive(f->function, f->params, a, b, epsabs, epsrel, limit, result, abserr) ? GSL_FAILURE
                                                                                                      : GSL_SUCCESS;
}

/* ---- histogram2d -------------------------------------------------------------------------- */
gsl_histogram2d *gsl_histogram2d_alloc(const size_t nx, const size_t ny)
{
    gsl_histogram2d *h = (gsl_histogram2d *)malloc(sizeof(*h));
    h->nx = nx;
    h->ny = ny;
    h->xrange = (double *)calloc(nx + 1, sizeof(double));
    h->yrange = (double *)calloc(ny + 1, sizeof(double));
    h->bin = (double *)calloc(nx * ny ? nx * ny : 1, sizeof(double));
    return h;
}
void gsl_histogram2d_free(gsl_histogram2d *h)
{
    if (!h) return;
    free(h->xrange);
    free(h->yrange);
    free(h->bin);
    free(h);
}
static void make_uniform(double range[], size_t n, double xmin, double xmax)
{
    for (size_t i = 0; i <= n; i++) {
        double f1 = ((double)(n - i) / (double)n);
        double f2 = ((double)i / (double)n);
        range[i] = f1 * xmin + f2 * xmax;
    }
}
int gsl_histogram2d_set_ranges_uniform(gsl_histogram2d *h, double xmin, double xmax, double ymin, double ymax)
{
    make_uniform(h->xrange, h->nx, xmin, xmax);
    make_uniform(h->yrange, h->ny, ymin, ymax);
    for (size_t i = 0; i < h->nx * h->ny; i++) h->bin[i] = 0;
    return GSL_SUCCESS;
}
static int find1(size_t n, const double range[], double x, size_t *i)
{
    if (x < range[0] || x >= range[n]) return 1;
    /* linear-guess then bisection, as gsl histogram/find.c */
    {
        double u = (x - range[0]) / (range[n] - range[0]);
        size_t g = (size_t)(u * n);
        if (g < n && x >= range[g] && x < range[g + 1]) {
            *i = g;
            return 0;
        }
    }
    {
        size_t lower = 0, upper = n;
        while (upper - lower > 1) {
            size_t mid = (upper + lower) / 2;
            if (x >= range[mid])
                lower = mid;
            else
                upper = mid;
        }
        *i = lower;
    }
    return 0;
}
int gsl_histogram2d_find(const gsl_histogram2d *h, const double x, const double y, size_t *i, size_t *j)
{
    if (find1(h->nx, h->xrange, x, i)) return GSL_EDOM;
    if (find1(h->ny, h->yrange, y, j)) return GSL_EDOM;
    return GSL_SUCCESS;
}
int gsl_histogram2d_increment(gsl_histogram2d *h, double x, double y)
{
    size_t i = 0, j = 0;
    if (gsl_histogram2d_find(h, x, y, &i, &j)) return GSL_EDOM;
    h->bin[i * h->ny + j] += 1.0;
    return GSL_SUCCESS;
}
double gsl_histogram2d_get(const gsl_histogram2d *h, const size_t i, const size_t j) { return h->bin[i * h->ny + j]; }
int gsl_histogram2d_fprintf(FILE *stream, const gsl_histogram2d *h, const char *range_format,
                            const char *bin_format)
{
    for (size_t i = 0; i < h->nx; i++)
        for (size_t j = 0; j < h->ny; j++) {
            fprintf(stream, range_format, h->xrange[i]);
            fputc(' ', stream);
            fprintf(stream, range_format, h->xrange[i + 1]);
            fputc(' ', stream);
            fprintf(stream, range_format, h->yrange[j]);
            fputc(' ', stream);
            fprintf(stream, range_format, h->yrange[j + 1]);
            fputc(' ', stream);
            fprintf(stream, bin_format, h->bin[i * h->ny + j]);
            fputc('\n', stream);
        }
    return GSL_SUCCESS;
}
