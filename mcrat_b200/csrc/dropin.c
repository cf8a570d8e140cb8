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

static void die(FILE *fPtr, const char *where)
{
    const char *msg = mcrat_b200_last_error(g_ctx);
    if (fPtr) {
        fprintf(fPtr, "mcrat_b200 (%s): %s\n", where, msg ? msg : "?");
        fflush(fPtr);
    }
    fprintf(stderr, "mcrat_b200 (%s): %s\n", where, msg ? msg : "?");
    exit(1);
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
    if (mcrat_b200_find_containing_hydro_cell(g_ctx, find_nearest_block_switch, &n) != 0) die(fPtr, "findContainingHydroCell");
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
    if (mcrat_b200_photon_event(g_ctx, dt_max, &ts, &idx, frame_scatt_cnt, frame_abs_cnt) != 0) die(fPtr, "photonEvent");
    *scattered_ph_index = idx;
    if (g_cs || !(ts < dt_max)) {
        /* host code may touch the whole list next (cyclo-synchrotron emission / rebinning,
         * Src/mcrat.c:799, 825) or the frame is over (remaining_time reaches 0, :784) */
        if (mcrat_b200_dropin_download(photon_list) != 0) die(fPtr, "download");
        if (g_cs) g_host_dirty = 1;
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
