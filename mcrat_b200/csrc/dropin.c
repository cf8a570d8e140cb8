/*
 * dropin.c -- the reference's function surface on top of the C ABI (see
 * include/mcrat_b200_dropin.h for the contract and the mirror policy).
 * Error behaviour follows the reference: failures are written to the rank log (fPtr)
 * and the process exits (Src/photons.c:54, Src/hot_x_section.c:102: `exit`), since the
 * reference's signatures have no error channel.
 */
#include "mcrat_b200_dropin.h"

#include <stdlib.h>
#include <string.h>

static mcrat_b200_ctx *g_ctx = NULL;
static mcrat_b200_config g_cfg;
static int g_configured = 0;
static int g_host_dirty = 1; /* host list newer than the device mirror */
static int g_cs = 0;
static int g_full_sort = -1; /* -1: ask the environment (MCRAT_B200_FULL_SORT) on first use */

static mcrat_dropin_photonList *g_last_list = NULL; /* the host list of the last call: where a failing run leaves its photons */

static void die(FILE *fPtr, const char *where)
{
    const char *msg = mcrat_b200_last_error(g_ctx);
    char copy[512];
    snprintf(copy, sizeof(copy), "%s", msg ? msg : "?"); /* the flush below may overwrite the message */
    if (fPtr) {
        fprintf(fPtr, "mcrat_b200 (%s): %s\n", where, copy);
        fflush(fPtr);
    }
    fprintf(stderr, "mcrat_b200 (%s): %s\n", where, copy);
    /* the reference exits with its list as it stands (Src/photons.c:54); ours stands on the device: bring back what
     * can be brought back, so that an atexit handler / core dump of the host sees the state the failure left */
    if (g_ctx && g_last_list && g_last_list->photons && !g_host_dirty) {
        int n = mcrat_b200_list_capacity(g_ctx);
        if (n > g_last_list->list_capacity) n = g_last_list->list_capacity;
        if (n > 0 && mcrat_b200_get_photons(g_ctx, g_last_list->photons, n) == 0 && fPtr) {
            fprintf(fPtr, "mcrat_b200: %d photons flushed from the device to the host list before exiting\n", n);
            fflush(fPtr);
        }
    }
    exit(1);
}

/* findContainingBlock's log line for every photon without a containing cell, Src/geometry.c:373-388 */
static void log_not_found(FILE *fPtr)
{
    int slots[32], total = 0, n, k;
    double h[96];
    if (!g_ctx) return;
    n = mcrat_b200_get_not_found(g_ctx, 32, slots, h, &total);
    if (n <= 0 || !fPtr) return;
    for (k = 0; k < n; k++) {
        if (g_cfg.dimensions == MCRAT_THREE) fprintf(fPtr, "3D switch is: %d and SIM switch is: %d\n", g_cfg.dimensions, 0);
        if (g_cfg.dimensions == MCRAT_THREE)
            fprintf(fPtr, "MCRaT Couldn't find a block for the photon located at r0=%e r1=%e r2=%e in the hydro simulation coordinate system.\n",
                    h[3 * k], h[3 * k + 1], h[3 * k + 2]);
        else
            fprintf(fPtr, "MCRaT Couldn't find a block for the photon located at r0=%e r1=%e\n", h[3 * k], h[3 * k + 1]);
    }
    if (total > n) fprintf(fPtr, "mcrat_b200: ... and %d more photons without a containing block\n", total - n);
    fflush(fPtr);
}

int mcrat_b200_dropin_configure(const mcrat_b200_config *cfg)
{
    if (g_ctx) {
        mcrat_b200_destroy(g_ctx);
        g_ctx = NULL;
    }
    g_cfg = *cfg;
    g_cfg.abi_version = MCRAT_B200_ABI_VERSION;
    g_configured = 1;
    g_cs = cfg->cyclosynch_switch;
    g_host_dirty = 1;
    return mcrat_b200_create(&g_cfg, &g_ctx);
}

void mcrat_b200_dropin_shutdown(void)
{
    if (g_ctx) mcrat_b200_destroy(g_ctx);
    g_ctx = NULL;
    g_configured = 0;
}

mcrat_b200_ctx *mcrat_b200_dropin_context(void) { return g_ctx; }

void mcrat_b200_dropin_mark_host_dirty(void) { g_host_dirty = 1; }

void mcrat_b200_dropin_set_full_sort(int on) { g_full_sort = on ? 1 : 0; }

static void ensure_ctx(FILE *fPtr)
{
    if (g_ctx) return;
    if (!g_configured) {
        const char *env = getenv("MCRAT_B200_CONFIG");
        mcrat_b200_config c;
        int dev = 0;
        memset(&c, 0, sizeof(c));
        if (!env || sscanf(env, "%d,%d,%d,%d,%d,%d,%lf,%d", &c.dimensions, &c.geometry, &c.stokes_switch,
                           &c.tau_calculation, &c.cyclosynch_switch, &c.b_field_calc, &c.epsilon_b, &dev) < 7) {
            fprintf(stderr, "mcrat_b200: call mcrat_b200_dropin_configure() or set MCRAT_B200_CONFIG\n");
            exit(1);
        }
        c.device = dev;
        c.rng_mode = MCRAT_RNG_PHILOX;
        if (mcrat_b200_dropin_configure(&c) != 0) die(fPtr, "create");
    } else if (mcrat_b200_create(&g_cfg, &g_ctx) != 0) {
        die(fPtr, "create");
    }
}

static void upload_hydro(mcrat_dropin_hydro_dataframe *h, FILE *fPtr)
{
    const double *fields[19] = {h->r0, h->r1, h->r2, h->r0_size, h->r1_size, h->r2_size, h->r, h->theta, h->v0, h->v1,
                                h->v2, h->dens, h->dens_lab, h->pres, h->temp, h->gamma, h->B0, h->B1, h->B2};
    double dom[6] = {h->r0_domain[0], h->r0_domain[1], h->r1_domain[0], h->r1_domain[1], h->r2_domain[0], h->r2_domain[1]};
    if (mcrat_b200_set_hydro(g_ctx, h->num_elements, fields, dom, h->fps, h->scatt_frame_number, h->inj_frame_number) != 0)
        die(fPtr, "set_hydro");
}

static void upload_photons(mcrat_dropin_photonList *l, FILE *fPtr)
{
    if (mcrat_b200_set_photons(g_ctx, l->photons, l->list_capacity) != 0) die(fPtr, "set_photons");
    g_host_dirty = 0;
}

int mcrat_b200_dropin_download(mcrat_dropin_photonList *l)
{
    if (!g_ctx) return MCRAT_B200_ERR_STATE;
    int n = mcrat_b200_list_capacity(g_ctx);
    if (n > l->list_capacity) n = l->list_capacity;
    return mcrat_b200_get_photons(g_ctx, l->photons, n);
}

/* findContainingHydroCell, Src/mclib.c:436 */
int __wrap_findContainingHydroCell(mcrat_dropin_photonList *photon_list, mcrat_dropin_hydro_dataframe *hydro_data,
                                   int find_nearest_block_switch, void *rand, FILE *fPtr)
{
    int n = 0;
    (void)rand;
    ensure_ctx(fPtr);
    if (find_nearest_block_switch != 0) {
        upload_hydro(hydro_data, fPtr); /* a new hydro frame was just read, Src/mcrat.c:721, 756 */
        upload_photons(photon_list, fPtr);
    } else if (g_host_dirty || mcrat_b200_list_capacity(g_ctx) != photon_list->list_capacity) {
        upload_photons(photon_list, fPtr);
    }
    g_last_list = photon_list;
    if (mcrat_b200_find_containing_hydro_cell(g_ctx, find_nearest_block_switch, &n) != 0) die(fPtr, "findContainingHydroCell");
    log_not_found(fPtr);
    return n;
}

/* calcMeanFreePath, Src/mclib.c:617.  Only the head of the time order is materialised on the
 * host (sorted_indexes[0], its time_to_scatter): that is all the driver reads, Src/mcrat.c:777. */
void __wrap_calcMeanFreePath(mcrat_dropin_photonList *photon_list, mcrat_dropin_hydro_dataframe *hydro_data, void *rand,
                             FILE *fPtr)
{
    int head = 0;
    double t = 0;
    (void)hydro_data;
    (void)rand;
    ensure_ctx(fPtr);
    if (mcrat_b200_calc_mean_free_path(g_ctx, &head, &t) != 0) die(fPtr, "calcMeanFreePath");
    if (g_full_sort < 0) g_full_sort = getenv("MCRAT_B200_FULL_SORT") ? 1 : 0;
    if (g_full_sort && photon_list->list_capacity > 0) {
        /* a host that reads more of the order than its head: all of sorted_indexes, sorted on the device */
        if (mcrat_b200_get_sorted_indexes(g_ctx, photon_list->sorted_indexes, photon_list->list_capacity) != 0) die(fPtr, "sorted_indexes");
    }
    if (photon_list->list_capacity > 0 && head >= 0 && head < photon_list->list_capacity) {
        photon_list->sorted_indexes[0] = head;
        photon_list->photons[head].time_to_scatter = t;
    }
}

/* photonEvent, Src/mclib.c:1107 */
double __wrap_photonEvent(mcrat_dropin_photonList *photon_list, double dt_max, mcrat_dropin_hydro_dataframe *hydro_data,
                          int *scattered_ph_index, int *frame_scatt_cnt, int *frame_abs_cnt, void *rand, FILE *fPtr)
{
    double ts = 0;
    int idx = 0;
    (void)hydro_data;
    (void)rand;
    ensure_ctx(fPtr);
    g_last_list = photon_list;
    if (mcrat_b200_photon_event(g_ctx, dt_max, &ts, &idx, frame_scatt_cnt, frame_abs_cnt) != 0) die(fPtr, "photonEvent");
    *scattered_ph_index = idx;
    if (!(ts < dt_max)) {
        /* the frame is over (remaining_time reaches 0, Src/mcrat.c:784): the whole list goes back.  With
         * CYCLOSYNCHROTRON_SWITCH ON what the driver does next to the list (pool replacement :799, rebinning :825) has its
         * own wrapped entry points, so one record suffices there too */
        if (mcrat_b200_dropin_download(photon_list) != 0) die(fPtr, "download");
    } else if (idx >= 0 && idx < photon_list->list_capacity) {
        if (mcrat_b200_get_photon(g_ctx, idx, &photon_list->photons[idx]) != 0) die(fPtr, "get_photon");
    }
    return ts;
}

/* updatePhotonPosition, Src/mclib.c:1054: the driver's own call ends a frame (Src/mcrat.c:841) */
void __wrap_updatePhotonPosition(mcrat_dropin_photonList *photon_list, double t, FILE *fPtr)
{
    ensure_ctx(fPtr);
    if (g_host_dirty || mcrat_b200_list_capacity(g_ctx) != photon_list->list_capacity) upload_photons(photon_list, fPtr);
    if (mcrat_b200_update_photon_position(g_ctx, t) != 0) die(fPtr, "updatePhotonPosition");
    if (mcrat_b200_dropin_download(photon_list) != 0) die(fPtr, "download");
}

/* averagePhotonEnergy, Src/mclib.c:1358 (logged every 1000 scatterings, Src/mcrat.c:814) */
double __wrap_averagePhotonEnergy(mcrat_dropin_photonList *photon_list)
{
    double e = 0;
    ensure_ctx(NULL);
    if (g_host_dirty || mcrat_b200_list_capacity(g_ctx) != photon_list->list_capacity) upload_photons(photon_list, NULL);
    if (mcrat_b200_average_photon_energy(g_ctx, &e) != 0) die(NULL, "averagePhotonEnergy");
    return e;
}

/* phAbsCyclosynch, Src/mc_cyclosynch.c:1571: runs after the loop on the host's list */
double __wrap_phAbsCyclosynch(mcrat_dropin_photonList *photon_list, int *num_abs_ph, int *scatt_cyclosynch_num_ph,
                              mcrat_dropin_hydro_dataframe *hydro_data, FILE *fPtr)
{
    double w = 0;
    int i, nulls = 0;
    (void)hydro_data;
    ensure_ctx(fPtr);
    upload_photons(photon_list, fPtr);
    if (mcrat_b200_ph_abs_cyclosynch(g_ctx, num_abs_ph, scatt_cyclosynch_num_ph, &w) != 0) die(fPtr, "phAbsCyclosynch");
    if (mcrat_b200_dropin_download(photon_list) != 0) die(fPtr, "download");
    /* setNullPhoton bookkeeping, Src/photons.c:247-249 */
    for (i = 0; i < photon_list->list_capacity; i++)
        if (photon_list->photons[i].type == 'N') nulls++;
    photon_list->num_null_photons = nulls;
    photon_list->num_photons = photon_list->list_capacity - nulls;
    return w;
}

/* ---- the rest of the surface the driver calls inside the hydro-frame loop --------------------------------- */
static void sync_list(mcrat_dropin_photonList *photon_list, FILE *fPtr)
{
    ensure_ctx(fPtr);
    if (g_host_dirty || mcrat_b200_list_capacity(g_ctx) != photon_list->list_capacity) upload_photons(photon_list, fPtr);
}

/* phMinMax, Src/mclib.c:1465 (Src/mcrat.c:704: before the hydro frame is read) */
void __wrap_phMinMax(mcrat_dropin_photonList *photon_list, double *min, double *max, double *min_theta, double *max_theta,
                     FILE *fPtr)
{
    sync_list(photon_list, fPtr);
    if (mcrat_b200_ph_min_max(g_ctx, min, max, min_theta, max_theta) != 0) die(fPtr, "phMinMax");
}

/* phScattStats, Src/mclib.c:1385 (Src/mcrat.c:745, 881) */
void __wrap_phScattStats(mcrat_dropin_photonList *photon_list, int *max, int *min, double *avg, double *r_avg, FILE *fPtr)
{
    sync_list(photon_list, fPtr);
    if (mcrat_b200_ph_scatt_stats(g_ctx, max, min, avg, r_avg) != 0) die(fPtr, "phScattStats");
}

/* calcCyclosynchRLimits, Src/mc_cyclosynch.c:225-242 */
double __wrap_calcCyclosynchRLimits(int frame_scatt, int frame_inj, double fps, double r_inj, char *min_or_max)
{
    return mcrat_b200_calc_cyclosynch_r_limits(frame_scatt, frame_inj, fps, r_inj, min_or_max);
}

/* grow the host list to the device's capacity the way reallocatePhotonListMemory does (Src/photons.c:57-80):
 * new slots are null photons */
static void grow_host_list(mcrat_dropin_photonList *l, int new_cap, FILE *fPtr)
{
    int i;
    mcrat_photon *p;
    int *s;
    if (new_cap <= l->list_capacity) return;
    p = (mcrat_photon *)realloc(l->photons, (size_t)new_cap * sizeof(mcrat_photon));
    s = (int *)realloc(l->sorted_indexes, (size_t)new_cap * sizeof(int));
    if (!p || !s) {
        if (fPtr) fprintf(fPtr, "mcrat_b200: could not grow the photon list to %d slots\n", new_cap);
        exit(1); /* Src/photons.c:66 */
    }
    for (i = l->list_capacity; i < new_cap; i++) {
        memset(&p[i], 0, sizeof(mcrat_photon));
        p[i].type = 'N';
        p[i].nearest_block_index = -1;
        s[i] = i;
    }
    l->num_null_photons += new_cap - l->list_capacity;
    l->photons = p;
    l->sorted_indexes = s;
    l->list_capacity = new_cap;
}

static void recount(mcrat_dropin_photonList *l)
{
    int i, nulls = 0;
    for (i = 0; i < l->list_capacity; i++)
        if (l->photons[i].type == 'N') nulls++;
    l->num_null_photons = nulls;
    l->num_photons = l->list_capacity - nulls;
}

/* photonEmitCyclosynch, Src/mc_cyclosynch.c:1176-1569: all-cells mode (K6) and single mode on the device */
int __wrap_photonEmitCyclosynch(mcrat_dropin_photonList *photon_list, double r_inj, double ph_weight, int maximum_photons,
                                double theta_min, double theta_max, mcrat_dropin_hydro_dataframe *hydro_data, void *rand,
                                int inject_single_switch, int scatt_ph_index, FILE *fPtr)
{
    int n = 0, ncells = 0, cap;
    double w = 0;
    (void)rand;
    ensure_ctx(fPtr);
    if (inject_single_switch != 0) {
        /* Src/mcrat.c:792-799: the driver has just set the scattered photon's type to COMPTONIZED on its own copy; the
         * device does the same to its copy, emits the replacement and re-positions the scattered photon */
        int slot = -1;
        if (mcrat_b200_photon_emit_cyclosynch_single(g_ctx, scatt_ph_index, &slot) != 0) die(fPtr, "photonEmitCyclosynch");
        cap = mcrat_b200_list_capacity(g_ctx);
        grow_host_list(photon_list, cap, fPtr);
        if (mcrat_b200_get_photon(g_ctx, scatt_ph_index, &photon_list->photons[scatt_ph_index]) != 0) die(fPtr, "get_photon");
        if (mcrat_b200_get_photon(g_ctx, slot, &photon_list->photons[slot]) != 0) die(fPtr, "get_photon");
        photon_list->num_photons += 1; /* incrementPhotonNum, Src/photons.c:253-262 */
        photon_list->num_null_photons -= 1;
        return 1;
    }
    upload_hydro(hydro_data, fPtr); /* the frame getHydroData has just read, Src/mcrat.c:721, 747 */
    upload_photons(photon_list, fPtr);
    if (mcrat_b200_photon_emit_cyclosynch(g_ctx, r_inj, ph_weight, maximum_photons, theta_min, theta_max, &n, &w, &ncells) != 0)
        die(fPtr, "photonEmitCyclosynch");
    if (fPtr) {
        fprintf(fPtr, "MCRaT has chosen %d hydro elements that it will emit cyclosynchrotron photons into.\n", ncells);
        if (ncells != 0)
            fprintf(fPtr, "Emitting %d cyclosynchrotron photon(s) with weight %e\n", n, w);
        else
            fprintf(fPtr, "Emitting 0 cyclosynchrotron photons\n");
        fflush(fPtr);
    }
    cap = mcrat_b200_list_capacity(g_ctx);
    grow_host_list(photon_list, cap, fPtr);
    if (mcrat_b200_dropin_download(photon_list) != 0) die(fPtr, "download");
    recount(photon_list);
    return n;
}

/* rebinCyclosynchCompPhotons, Src/mc_cyclosynch.c:600-710 */
int __wrap_rebinCyclosynchCompPhotons(mcrat_dropin_photonList *photon_list, int *num_cyclosynch_ph_emit,
                                      int *scatt_cyclosynch_num_ph, int max_photons, double thread_theta_min,
                                      double thread_theta_max, void *rand, FILE *fPtr)
{
    int nnull = 0;
    (void)thread_theta_min;
    (void)thread_theta_max;
    (void)rand;
    sync_list(photon_list, fPtr);
    if (mcrat_b200_rebin_cyclosynch_comp_photons(g_ctx, max_photons, num_cyclosynch_ph_emit, scatt_cyclosynch_num_ph, &nnull) != 0)
        die(fPtr, "rebinCyclosynchCompPhotons");
    if (mcrat_b200_dropin_download(photon_list) != 0) die(fPtr, "download");
    recount(photon_list);
    return nnull;
}

/* initalizeHotCrossSection / cleanupInterpolationData, Src/hot_x_section.c:17-80, 507-542 */
static char g_table_path[1024] = "";

void mcrat_b200_dropin_set_table_path(const char *path)
{
    snprintf(g_table_path, sizeof(g_table_path), "%s", path ? path : "");
}

static const char *table_path(void)
{
    const char *env = getenv("MCRAT_B200_HOT_X_SECTION_FILE");
    if (g_table_path[0]) return g_table_path;
    if (env && env[0]) return env;
    return "thermal_hot_x_section.dat"; /* HOT_THERMAL_X_SECTION_FILE */
}

#define DROPIN_N_PH_E 220 /* Src/hot_x_section.h:2-10 */
#define DROPIN_N_T 80

static int is_dash_line(const char *line) /* Src/hot_x_section.c:307-322 */
{
    int found = 0;
    for (; *line; ++line) {
        if (*line == ' ' || *line == '\t' || *line == '\r' || *line == '\n') continue;
        if (*line != '-') return 0;
        found = 1;
    }
    return found;
}

void __wrap_initalizeHotCrossSection(int rank, void *rand, FILE *fPtr)
{
    static double table[(DROPIN_N_PH_E + 1) * (DROPIN_N_T + 1)];
    const char *path = table_path();
    FILE *fp;
    char line[1024];
    int i, j;
    (void)rand;
    ensure_ctx(fPtr);
    fp = fopen(path, "r");
    if (fp) { /* readHotCrossSection, Src/hot_x_section.c:208-258: skip to the dashed line, then `i j logx logtheta value` rows */
        int rows = 0;
        double a, b, v;
        if (fPtr) fprintf(fPtr, "Reading thermal hot cross section data from %s...\n", path);
        while (fgets(line, sizeof(line), fp) && !is_dash_line(line)) {}
        while (fgets(line, sizeof(line), fp)) {
            if (sscanf(line, "%d\t%d\t%lf\t%lf\t%lf", &i, &j, &a, &b, &v) != 5) continue;
            if (i < 0 || i > DROPIN_N_PH_E || j < 0 || j > DROPIN_N_T) {
                if (fPtr) fprintf(fPtr, "The bounds of the thermal input file exceed what MCRaT has been compiled with.\n");
                exit(0);
            }
            table[i * (DROPIN_N_T + 1) + j] = v;
            rows++;
        }
        fclose(fp);
        if (rows != (DROPIN_N_PH_E + 1) * (DROPIN_N_T + 1)) {
            if (fPtr) fprintf(fPtr, "mcrat_b200: %s holds %d rows, expected %d\n", path, rows, (DROPIN_N_PH_E + 1) * (DROPIN_N_T + 1));
            exit(0);
        }
        if (mcrat_b200_set_thermal_table(g_ctx, table) != 0) die(fPtr, "set_thermal_table");
        return;
    }
    /* createHotCrossSection, Src/hot_x_section.c:82-206: 500 000 samples per point, on the device; every rank builds the same
     * table from the same keyed streams, rank 0 writes the file */
    {
        float ms = 0;
        if (fPtr) fprintf(fPtr, "Creating the thermal hot cross section table on the device...\n");
        if (mcrat_b200_build_thermal_table(g_ctx, 500000, 20130313ull, table, &ms) != 0) die(fPtr, "build_thermal_table");
        if (fPtr) fprintf(fPtr, "done in %.1f ms.\n", ms);
    }
    if (rank == 0) {
        const double dt = (4.0 - (-4.0)) / DROPIN_N_T, dph_e = (6.0 - (-12.0)) / DROPIN_N_PH_E;
        fp = fopen(path, "w");
        if (!fp) {
            if (fPtr) fprintf(fPtr, "couldn't write to file %s\n", path);
            exit(0);
        }
        fprintf(fp, "The comoving photon energy and the temperatures are normalized by the electron rest mass\n");
        fprintf(fp, "The calculated hot cross sections are normalized by the thompson cross section.\n");
        fprintf(fp, "Photon index\tTheta Index\tlog10(Comoving Photon Energy)\tlog10(Theta)\tlog10(Hot Cross Section)\n");
        fprintf(fp, "------------------------------------------------\n");
        for (i = 0; i <= DROPIN_N_PH_E; i++)
            for (j = 0; j <= DROPIN_N_T; j++)
                fprintf(fp, "%d\t%d\t%g\t%g\t%15.10g\n", i, j, -12.0 + i * dph_e, -4.0 + j * dt, table[i * (DROPIN_N_T + 1) + j]);
        fclose(fp);
        if (fPtr) fprintf(fPtr, "done writing to file.\n\n");
    }
}

/* the interpolation state lives in the device context; nothing to free before mcrat_b200_dropin_shutdown() */
void __wrap_cleanupInterpolationData(void) {}
