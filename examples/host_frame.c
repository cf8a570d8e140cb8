/* host_frame.c -- a C host driving the B200 hot path through the C ABI only (no Python, no torch).
 *
 * It is the body of MCRaT's hydro-frame loop (Src/mcrat.c:664-918) with the device path in place of
 * the while-loop at :761-851, reading the same inputs a MCRaT run has at that point:
 *
 *   host_frame <dir> [frames] [max_iters]
 *     <dir>/mcrat_input.h   compile-time switches of the reference, read at run time
 *     <dir>/mc.par          readMcPar's file
 *     <dir>/hydro.bin       int32 n, then 19 arrays of n doubles in struct hydro_dataframe order
 *     <dir>/photons.bin     int32 n, then n records of struct photon (176 bytes)
 *   writes <dir>/mc_proc_0.h5 (one group per frame) and <dir>/mcdata_<frame>.h5, prints one line per frame.
 *
 * Build: gcc -O2 -std=gnu11 examples/host_frame.c -Iinclude -Lmcrat_b200/csrc -lmcrat_b200 -lmcrat_b200_io \
 *            -Wl,-rpath,$PWD/mcrat_b200/csrc -o host_frame
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mcrat_b200.h"
#include "mcrat_b200_io.h"

static void *read_all(const char *path, size_t *bytes)
{
    FILE *f = fopen(path, "rb");
    if (!f) {
        fprintf(stderr, "cannot open %s\n", path);
        exit(2);
    }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    void *p = malloc((size_t)n);
    if (fread(p, 1, (size_t)n, f) != (size_t)n) {
        fprintf(stderr, "short read on %s\n", path);
        exit(2);
    }
    fclose(f);
    *bytes = (size_t)n;
    return p;
}

int main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s <dir> [frames] [max_iters]\n", argv[0]);
        return 2;
    }
    const char *dir = argv[1];
    const int frames = argc > 2 ? atoi(argv[2]) : 2;
    const long long max_iters = argc > 3 ? atoll(argv[3]) : 200;
    char path[1024];

    mcrat_b200_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    mcrat_b200_io_switches sw;
    mcrat_b200_mc_par par;
    snprintf(path, sizeof(path), "%s/mcrat_input.h", dir);
    if (mcrat_b200_config_from_input_header(path, &cfg, &sw) != MCRAT_IO_OK) {
        fprintf(stderr, "%s\n", mcrat_b200_io_last_error());
        return 1;
    }
    snprintf(path, sizeof(path), "%s/%s", dir, sw.mcpar);
    if (mcrat_b200_read_mc_par(path, &par) != MCRAT_IO_OK) {
        fprintf(stderr, "%s\n", mcrat_b200_io_last_error());
        return 1;
    }
    cfg.device = 0;
    cfg.shard = 0; /* MPI rank of the reference */
    cfg.seed = 20261018;
    cfg.rng_mode = MCRAT_RNG_PHILOX;

    size_t hb = 0, pb = 0;
    snprintf(path, sizeof(path), "%s/hydro.bin", dir);
    unsigned char *hraw = read_all(path, &hb);
    snprintf(path, sizeof(path), "%s/photons.bin", dir);
    unsigned char *praw = read_all(path, &pb);
    int ncell = 0, nph = 0;
    memcpy(&ncell, hraw, 4);
    memcpy(&nph, praw, 4);
    if (hb != 8 + (size_t)19 * ncell * 8 || pb != 8 + (size_t)nph * sizeof(mcrat_photon)) {
        fprintf(stderr, "hydro.bin / photons.bin have unexpected sizes\n");
        return 1;
    }
    const double *fields[19];
    for (int k = 0; k < 19; ++k) fields[k] = (const double *)(hraw + 8) + (size_t)k * ncell;
    mcrat_photon *photons = (mcrat_photon *)(praw + 8);
    const double dom[6] = {par.r0_domain[0], par.r0_domain[1], par.r1_domain[0], par.r1_domain[1], par.r2_domain[0], par.r2_domain[1]};

    mcrat_b200_ctx *g = NULL;
    if (mcrat_b200_create(&cfg, &g) != MCRAT_B200_OK) {
        fprintf(stderr, "mcrat_b200_create: %s\n", mcrat_b200_last_error(NULL));
        return 1;
    }
#define CK(call)                                                                  \
    do {                                                                          \
        int rc__ = (call);                                                        \
        if (rc__ != MCRAT_B200_OK) {                                              \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc__, mcrat_b200_last_error(g)); \
            return 1;                                                             \
        }                                                                         \
    } while (0)

    const int frame0 = par.frm0[0];
    double time_now = (double)frame0 / par.fps;
    CK(mcrat_b200_set_photons(g, photons, nph));
    for (int scatt_frame = frame0; scatt_frame < frame0 + frames; ++scatt_frame) {
        /* getHydroData(&hydrodata, ...) would run here (Src/mcrat.c:721); the same frame is re-used */
        CK(mcrat_b200_set_hydro(g, ncell, fields, dom, par.fps, scatt_frame, frame0));
        double remaining_time = ((double)(scatt_frame + 1) / par.fps) - time_now; /* Src/mcrat.c:754 */
        mcrat_b200_frame_stats st;
        CK(mcrat_b200_run_frame(g, time_now, remaining_time, max_iters, 1, &st));
        time_now = st.time_now;
        int max_s = 0, min_s = 0;
        double avg_s = 0, avg_r = 0, avg_e = 0;
        CK(mcrat_b200_ph_scatt_stats(g, &max_s, &min_s, &avg_s, &avg_r)); /* phScattStats, Src/mcrat.c:881 */
        CK(mcrat_b200_average_photon_energy(g, &avg_e));
        CK(mcrat_b200_get_photons(g, photons, nph));                      /* for printPhotons / saveCheckpoint, :898-907 */
        if (mcrat_b200_print_photons(dir, 0, scatt_frame, photons, nph, &sw) != MCRAT_IO_OK ||
            mcrat_b200_merge_frame(dir, scatt_frame, (const int[]){0}, 1, &sw) != MCRAT_IO_OK) {
            fprintf(stderr, "%s\n", mcrat_b200_io_last_error());
            return 1;
        }
        printf("frame %d iterations %lld scatterings %lld relocations %lld time_now %.17g avg_scatt %.17g avg_r %.17g avg_e %.17g\n",
               scatt_frame, st.iterations, st.scatterings, st.relocations, st.time_now, avg_s, avg_r, avg_e);
    }
    mcrat_b200_destroy(g);
    free(hraw);
    free(praw);
    return 0;
}
