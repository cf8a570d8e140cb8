/*
 * ref_harness.c -- TEST INFRASTRUCTURE (oracle).
 *
 * Flat C entry points (ctypes-friendly) around the reference's own hot-path
 * functions.  This file is compiled TOGETHER WITH THE REFERENCE'S UNMODIFIED
 * SOURCES (where they lie under /root/reference/Src, via oracle/build_ref.py)
 * into oracle/_ref/libmcrat_ref_<cfg>.so, one library per compile-time
 * configuration (mcrat_input.h is #define-only).  It contains no reference
 * code: it only builds `struct hydro_dataframe` / `struct photonList` values
 * from flat arrays, calls the reference functions, and restates the driver's
 * scatter-frame while-loop (Src/mcrat.c:754-851) so that whole frames can be
 * replayed.
 */
#include "mcrat.h"

#define REF_API __attribute__((visibility("default")))

/* ---- configuration report ---------------------------------------------------- */
REF_API int ref_sizeof_photon(void) { return (int)sizeof(struct photon); }

REF_API void ref_get_config(int *out)
{
    out[0] = DIMENSIONS;
    out[1] = GEOMETRY;
    out[2] = STOKES_SWITCH;
    out[3] = TAU_CALCULATION;
    out[4] = CYCLOSYNCHROTRON_SWITCH;
    out[5] = COMV_SWITCH;
#if CYCLOSYNCHROTRON_SWITCH == ON
    out[6] = B_FIELD_CALC;
#else
    out[6] = -1;
#endif
    out[7] = SIM_SWITCH;
}

REF_API void ref_get_constants(double *out)
{
    out[0] = C_LIGHT; out[1] = A_RAD; out[2] = PL_CONST; out[3] = K_B; out[4] = M_P;
    out[5] = THOM_X_SECT; out[6] = M_EL; out[7] = FINE_STRUCT; out[8] = CHARGE_EL; out[9] = R_EL;
}

/* ---- hydro frame ----------------------------------------------------------------- */
static double *dup_or_zero(const double *src, int n)
{
    double *d = malloc((n > 0 ? n : 1) * sizeof(double));
    if (src)
        memcpy(d, src, n * sizeof(double));
    else
        for (int i = 0; i < n; i++) d[i] = 0.0;
    return d;
}

/* arrays: 19 pointers in struct order (r0 r1 r2 r0_size r1_size r2_size r theta v0 v1 v2 dens
 * dens_lab pres temp gamma B0 B1 B2); NULL entries become zero arrays */
REF_API void *ref_hydro_new(int n, const double **arrays, const double *domains, double fps, int scatt_frame,
                            int inj_frame)
{
    struct hydro_dataframe *h = calloc(1, sizeof(*h));
    h->num_elements = n;
    h->r0 = dup_or_zero(arrays[0], n);
    h->r1 = dup_or_zero(arrays[1], n);
    h->r2 = dup_or_zero(arrays[2], n);
    h->r0_size = dup_or_zero(arrays[3], n);
    h->r1_size = dup_or_zero(arrays[4], n);
    h->r2_size = dup_or_zero(arrays[5], n);
    h->r = dup_or_zero(arrays[6], n);
    h->theta = dup_or_zero(arrays[7], n);
    h->v0 = dup_or_zero(arrays[8], n);
    h->v1 = dup_or_zero(arrays[9], n);
    h->v2 = dup_or_zero(arrays[10], n);
    h->dens = dup_or_zero(arrays[11], n);
    h->dens_lab = dup_or_zero(arrays[12], n);
    h->pres = dup_or_zero(arrays[13], n);
    h->temp = dup_or_zero(arrays[14], n);
    h->gamma = dup_or_zero(arrays[15], n);
    h->B0 = dup_or_zero(arrays[16], n);
    h->B1 = dup_or_zero(arrays[17], n);
    h->B2 = dup_or_zero(arrays[18], n);
    h->r0_domain[0] = domains[0]; h->r0_domain[1] = domains[1];
    h->r1_domain[0] = domains[2]; h->r1_domain[1] = domains[3];
    h->r2_domain[0] = domains[4]; h->r2_domain[1] = domains[5];
    h->fps = fps;
    h->scatt_frame_number = scatt_frame;
    h->inj_frame_number = inj_frame;
    h->last_frame = scatt_frame;
    h->increment_inj_frame = 1;
    h->increment_scatt_frame = 1;
    h->grid = NULL; /* Src/mcrat_io.c:1985 */
    return h;
}

REF_API void ref_hydro_free(void *vh)
{
    struct hydro_dataframe *h = vh;
    if (!h) return;
    free(h->r0); free(h->r1); free(h->r2); free(h->r0_size); free(h->r1_size); free(h->r2_size);
    free(h->r); free(h->theta); free(h->v0); free(h->v1); free(h->v2); free(h->dens); free(h->dens_lab);
    free(h->pres); free(h->temp); free(h->gamma); free(h->B0); free(h->B1); free(h->B2);
    free(h);
}

/* copy (possibly analytic-outflow-overwritten) cell fields back out, struct order */
REF_API void ref_hydro_get(void *vh, int field, double *out)
{
    struct hydro_dataframe *h = vh;
    double *f[19] = {h->r0, h->r1, h->r2, h->r0_size, h->r1_size, h->r2_size, h->r, h->theta, h->v0, h->v1,
                     h->v2, h->dens, h->dens_lab, h->pres, h->temp, h->gamma, h->B0, h->B1, h->B2};
    memcpy(out, f[field], h->num_elements * sizeof(double));
}

/* analytic outflows (Src/analytic_outflows.c); kind 1 cylindrical, 2 spherical, 3 structured */
REF_API void ref_hydro_analytic(void *vh, int kind, const char *logpath)
{
    struct hydro_dataframe *h = vh;
    FILE *f = fopen(logpath ? logpath : "/dev/null", "a");
    fillHydroCoordinateToSpherical(h);
    if (kind == 1) cylindricalPrep(h, f);
    if (kind == 2) sphericalPrep(h, f);
    if (kind == 3) structuredFireballPrep(h, f);
    fclose(f);
}

/* ---- photon list ------------------------------------------------------------------- */
REF_API void *ref_list_new(const void *photons, int n)
{
    struct photonList *l = malloc(sizeof(*l));
    initalizePhotonList(l);
    if (n > 0) {
        setPhotonList(l, (struct photon *)photons, n);
        /* setPhotonList (Src/photons.c:83-108) counts the null photons of the array but leaves num_photons at n; the
         * driver never hands it a list with nulls, the tests do: restore the invariant verifyPhotonNum checks */
        l->num_photons = n - l->num_null_photons;
    }
    return l;
}
REF_API void ref_list_free(void *vl)
{
    struct photonList *l = vl;
    if (!l) return;
    freePhotonList(l);
    free(l);
}
REF_API int ref_list_capacity(void *vl) { return ((struct photonList *)vl)->list_capacity; }
REF_API int ref_list_num_photons(void *vl) { return ((struct photonList *)vl)->num_photons; }
REF_API int ref_list_num_null(void *vl) { return ((struct photonList *)vl)->num_null_photons; }
REF_API void ref_list_get(void *vl, void *out)
{
    struct photonList *l = vl;
    memcpy(out, l->photons, (size_t)l->list_capacity * sizeof(struct photon));
}
REF_API void ref_list_get_sorted(void *vl, int *out)
{
    struct photonList *l = vl;
    memcpy(out, l->sorted_indexes, (size_t)l->list_capacity * sizeof(int));
}

/* ---- rng ------------------------------------------------------------------------------ */
REF_API void *ref_rng_new(unsigned long seed)
{
    gsl_rng *r;
    gsl_rng_env_setup();
    r = gsl_rng_alloc(gsl_rng_ranlxs0);
    if (seed) gsl_rng_set(r, seed);
    return r;
}
REF_API void ref_rng_free(void *r) { gsl_rng_free(r); }
REF_API void ref_rng_use_replay(void *r, const double *buf, size_t n) { gsl_shim_rng_use_replay(r, buf, n); }
REF_API void ref_rng_set_tee(void *r, double *buf, size_t cap) { gsl_shim_rng_set_tee(r, buf, cap); }
REF_API size_t ref_rng_tee_count(void *r) { return gsl_shim_rng_tee_count(r); }
REF_API unsigned long long ref_rng_draws(void *r) { return gsl_shim_rng_draws(r); }
REF_API double ref_rng_uniform(void *r) { return gsl_rng_uniform(r); }
REF_API unsigned long ref_rng_get(void *r) { return gsl_rng_get(r); }
REF_API void ref_rng_set(void *r, unsigned long s) { gsl_rng_set(r, s); }

/* ---- log file ---------------------------------------------------------------------------- */
static FILE *g_log = NULL;
REF_API void ref_set_log(const char *path)
{
    if (g_log) fclose(g_log);
    g_log = fopen(path ? path : "/dev/null", "w");
}
static FILE *ref_logf(void)
{
    if (!g_log) g_log = fopen("/dev/null", "w");
    return g_log;
}

/* ---- the boundary functions ------------------------------------------ */
REF_API int ref_findContainingHydroCell(void *list, void *hydro, int sw, void *rng)
{
    return findContainingHydroCell(list, hydro, sw, rng, ref_logf());
}
REF_API void ref_calcMeanFreePath(void *list, void *hydro, void *rng) { calcMeanFreePath(list, hydro, rng, ref_logf()); }
REF_API double ref_photonEvent(void *list, double dt_max, void *hydro, int *scattered_ph_index, int *frame_scatt_cnt,
                               int *frame_abs_cnt, void *rng)
{
    return photonEvent(list, dt_max, hydro, scattered_ph_index, frame_scatt_cnt, frame_abs_cnt, rng, ref_logf());
}
REF_API void ref_updatePhotonPosition(void *list, double t) { updatePhotonPosition(list, t, ref_logf()); }

/* statistics readers the driver calls between kernels */
REF_API void ref_phMinMax(void *list, double *out4)
{
    phMinMax(list, &out4[0], &out4[1], &out4[2], &out4[3], ref_logf());
}
REF_API void ref_phScattStats(void *list, int *max, int *min, double *avg, double *r_avg)
{
    phScattStats(list, max, min, avg, r_avg, ref_logf());
}
REF_API double ref_averagePhotonEnergy(void *list) { return averagePhotonEnergy(list); }

/* ---- unit-level entry points -------------------------------------------------------------------- */
REF_API void ref_mcratCoordinateToHydroCoordinate(double *out3, double x, double y, double z)
{
    mcratCoordinateToHydroCoordinate(out3, x, y, z);
}
REF_API void ref_hydroVectorToCartesian(double *out3, double v0, double v1, double v2, double x0, double x1, double x2)
{
    hydroVectorToCartesian(out3, v0, v1, v2, x0, x1, x2);
}
REF_API int ref_checkInBlock(double a, double b, double c, void *hydro, int idx) { return checkInBlock(a, b, c, hydro, idx); }
REF_API int ref_findContainingBlock(double a, double b, double c, void *hydro)
{
    return findContainingBlock(a, b, c, hydro, ref_logf());
}
REF_API double ref_hydroElementVolume(void *hydro, int idx) { return hydroElementVolume(hydro, idx); }
REF_API void ref_lorentzBoost(double *boost, double *p, double *result, char object)
{
    lorentzBoost(boost, p, result, object, ref_logf());
}
REF_API void ref_calculateOpticalDepth(void *photon, void *hydro, void *rng)
{
    calculateOpticalDepth(photon, hydro, rng, ref_logf());
}
REF_API double ref_kleinNishinaCrossSection(double x) { return kleinNishinaCrossSection(x); }
REF_API int ref_kleinNishinaScatter(double *theta, double *phi, double p0, double q, double u, void *rng)
{
    return kleinNishinaScatter(theta, phi, p0, q, u, rng, ref_logf());
}
REF_API int ref_singleScatter(double *el, double *ph, double *s, void *rng) { return singleScatter(el, ph, s, rng, ref_logf()); }
REF_API void ref_singleThermalElectron(double *el_p, double temp, double *ph_p, void *rng)
{
    singleThermalElectron(el_p, temp, ph_p, rng, ref_logf());
}
REF_API double ref_sampleThermalElectron(double temp, void *rng) { return sampleThermalElectron(temp, rng, ref_logf()); }
REF_API void ref_stokesRotation(double *v, double *v_ph, double *v_ph_boosted, double *s)
{
    stokesRotation(v, v_ph, v_ph_boosted, s, ref_logf());
}
REF_API void ref_mullerMatrixRotation(double theta, double *s) { mullerMatrixRotation(theta, s, ref_logf()); }
REF_API double ref_singleMaxwellJuttner(double gamma, double theta) { return singleMaxwellJuttner(gamma, theta); }
REF_API double ref_boostedCrossSection(double e, double mu, double gamma) { return boostedCrossSection(e, mu, gamma); }
REF_API double ref_calculateTotalThermalCrossSection(double e, double theta, void *rng)
{
    return calculateTotalThermalCrossSection(e, theta, rng, ref_logf());
}

/* ---- hot cross-section table (TAU_CALCULATION == TABLE) ------------------------------------------ */
extern double thermal_table[N_PH_E + 1][N_T + 1];
static int g_table_ready = 0;
REF_API void ref_table_dims(int *out) { out[0] = N_PH_E + 1; out[1] = N_T + 1; }
REF_API void ref_set_thermal_table(const double *tab)
{
    if (g_table_ready) cleanupInterpolationData();
    memcpy(thermal_table, tab, sizeof(double) * (N_PH_E + 1) * (N_T + 1));
    initalizeHotCrossSectionInterp();
    g_table_ready = 1;
}
REF_API double ref_interpolateThermalHotCrossSection(double log_e, double log_theta, void *rng)
{
    return interpolateThermalHotCrossSection(log_e, log_theta, rng, ref_logf());
}
REF_API double ref_getThermalCrossSection(double comv_e, double temp, void *rng)
{
    return getThermalCrossSection(comv_e, temp, rng, ref_logf());
}

/* ---- cyclo-synchrotron ------------------------------------------------------------------------------ */
REF_API double ref_calcCyclosynchRLimits(int fs, int fi, double fps, double r_inj, const char *which)
{
    return calcCyclosynchRLimits(fs, fi, fps, r_inj, (char *)which);
}
REF_API double ref_getMagneticFieldMagnitude(void *hydro, int idx)
{
#if CYCLOSYNCHROTRON_SWITCH == ON
    return getMagneticFieldMagnitude(hydro, idx);
#else
    (void)hydro; (void)idx;
    return 0;
#endif
}
REF_API int ref_photonEmitCyclosynch(void *list, double r_inj, double ph_weight, int max_photons, double theta_min,
                                     double theta_max, void *hydro, void *rng, int single, int scatt_idx)
{
#if CYCLOSYNCHROTRON_SWITCH == ON
    return photonEmitCyclosynch(list, r_inj, ph_weight, max_photons, theta_min, theta_max, hydro, rng, single,
                                scatt_idx, ref_logf());
#else
    (void)list; (void)r_inj; (void)ph_weight; (void)max_photons; (void)theta_min; (void)theta_max; (void)hydro;
    (void)rng; (void)single; (void)scatt_idx;
    return -1;
#endif
}
REF_API double ref_phAbsCyclosynch(void *list, int *num_abs, int *scatt_cs_num, void *hydro)
{
#if CYCLOSYNCHROTRON_SWITCH == ON
    return phAbsCyclosynch(list, num_abs, scatt_cs_num, hydro, ref_logf());
#else
    (void)list; (void)num_abs; (void)scatt_cs_num; (void)hydro;
    return -1;
#endif
}
REF_API int ref_rebinCyclosynchCompPhotons(void *list, int *num_emit, int *scatt_cs_num, int max_photons,
                                           double theta_min, double theta_max, void *rng)
{
#if CYCLOSYNCHROTRON_SWITCH == ON
    return rebinCyclosynchCompPhotons(list, num_emit, scatt_cs_num, max_photons, theta_min, theta_max, rng, ref_logf());
#else
    (void)list; (void)num_emit; (void)scatt_cs_num; (void)max_photons; (void)theta_min; (void)theta_max; (void)rng;
    return -1;
#endif
}

/* ---- photon injection (used to generate inputs; Src/mclib.c:9-300) ------------------------------------ */
REF_API int ref_photonInjection(void *list, double r_inj, double ph_weight, int min_photons, int max_photons,
                                char spect, double theta_min, double theta_max, void *hydro, void *rng)
{
    photonInjection(list, r_inj, ph_weight, min_photons, max_photons, spect, theta_min, theta_max, hydro, rng, ref_logf());
    return ((struct photonList *)list)->num_photons;
}

/* ---- one scatter frame: the driver's while-loop, Src/mcrat.c:754-851 ----------------------------------- */
struct ref_frame_stats {
    long long iterations;
    long long scatterings;     /* frame_scatt_cnt */
    long long relocations;     /* num_photons_find_new_element */
    long long photon_slots;    /* sum over iterations of list_capacity */
    double time_now;
    double last_time_step;
    int cs_emitted;
    int scatt_cyclosynch_num_ph;
};

/* cs_* arguments are used only when CYCLOSYNCHROTRON_SWITCH == ON.
 * max_iters < 0 means run until remaining_time reaches 0. */
REF_API void ref_run_frame(void *vlist, void *vhydro, void *vrng, double time_now, double remaining_time,
                           long long max_iters, int find_nearest_grid_switch, double cs_r_inj, double cs_ph_weight,
                           int cs_max_photons, double cs_theta_min, double cs_theta_max,
                           struct ref_frame_stats *st)
{
    struct photonList *photon_list = vlist;
    struct hydro_dataframe *hydrodata = vhydro;
    gsl_rng *rng = vrng;
    FILE *fPtr = ref_logf();
    int frame_scatt_cnt = 0, frame_abs_cnt = 0, ph_scatt_index = 0;
    int num_photons_find_new_element = 0;
    int num_cyclosynch_ph_emit = 0, scatt_cyclosynch_num_ph = st->scatt_cyclosynch_num_ph;
    double time_step = 0, n_comptonized = 0;
    long long iters = 0, slots = 0;
    struct photon *scattered_photon = NULL;
    (void)cs_r_inj; (void)cs_ph_weight; (void)cs_max_photons; (void)cs_theta_min; (void)cs_theta_max;
    (void)n_comptonized; (void)scattered_photon;

    while (remaining_time > 0 && (max_iters < 0 || iters < max_iters)) {
        num_photons_find_new_element +=
            findContainingHydroCell(photon_list, hydrodata, find_nearest_grid_switch, rng, fPtr);
        calcMeanFreePath(photon_list, hydrodata, rng, fPtr);
        find_nearest_grid_switch = 0;
        slots += photon_list->list_capacity;

        if (getPhoton(photon_list, photon_list->sorted_indexes[0])->time_to_scatter < remaining_time) {
            time_step = photonEvent(photon_list, remaining_time, hydrodata, &ph_scatt_index, &frame_scatt_cnt,
                                    &frame_abs_cnt, rng, fPtr);
            time_now += time_step;
            remaining_time -= time_step;
            scattered_photon = getPhoton(photon_list, ph_scatt_index);
#if CYCLOSYNCHROTRON_SWITCH == ON
            if (scattered_photon->type == CS_POOL_PHOTON) {
                n_comptonized += scattered_photon->weight;
                scattered_photon->type = COMPTONIZED_PHOTON;
                num_cyclosynch_ph_emit +=
                    photonEmitCyclosynch(photon_list, cs_r_inj, cs_ph_weight, cs_max_photons, cs_theta_min,
                                         cs_theta_max, hydrodata, rng, 1, ph_scatt_index, fPtr);
                scatt_cyclosynch_num_ph++;
            }
            if ((frame_scatt_cnt % 1000 == 0) && (frame_scatt_cnt != 0)) {
                if (scatt_cyclosynch_num_ph > cs_max_photons)
                    rebinCyclosynchCompPhotons(photon_list, &num_cyclosynch_ph_emit, &scatt_cyclosynch_num_ph,
                                               cs_max_photons, cs_theta_min, cs_theta_max, rng, fPtr);
            }
#endif
        } else {
            time_now += remaining_time;
            updatePhotonPosition(photon_list, remaining_time, fPtr);
            time_step = remaining_time;
            remaining_time = 0;
        }
        iters++;
    }
    st->iterations = iters;
    st->scatterings = frame_scatt_cnt;
    st->relocations = num_photons_find_new_element;
    st->photon_slots = slots;
    st->time_now = time_now;
    st->last_time_step = time_step;
    st->cs_emitted = num_cyclosynch_ph_emit;
    st->scatt_cyclosynch_num_ph = scatt_cyclosynch_num_ph;
}
