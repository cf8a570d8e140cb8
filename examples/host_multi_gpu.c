/* host_multi_gpu.c -- the job of host_frame.c split over several GPUs of one node, in C, through the C ABI only.
 *
 * One host thread per GPU stands for one MPI rank of the reference (Src/mcrat.c:93-95, 139-164): every rank owns a contiguous
 * slice of the photon list, a full replica of the hydro frame and its own Philox shard key.  Nothing is exchanged inside the
 * frame loop.  Around it the ranks use the library's communicator (mcrat_b200_comm_*, NCCL) the way the reference uses MPI:
 *   - the hot cross-section table is built by all GPUs together and assembled by an all-gather (TAU_CALCULATION TABLE only;
 *     the reference builds it on rank 0 and MPI_Bcasts it, Src/hot_x_section.c:717),
 *   - after every hydro frame the frame counters are reduced over the ranks (rank 0 prints the job's totals),
 *   - at the end the photons of all ranks are gathered on rank 0 in rank order (what Src/merge.c:840-876 does with
 *     MPI_Allgatherv) and written as one mc_proc / mcdata pair.
 * In an MPI host the unique id goes through MPI_Bcast instead of the shared buffer used here.
 *
 *   host_multi_gpu <dir> <gpus> [frames] [max_iters]        inputs as for host_frame (mcrat_input.h, mc.par, hydro.bin, photons.bin)
 *
 * Build: gcc -O2 -std=gnu11 -pthread examples/host_multi_gpu.c -Iinclude -Lmcrat_b200/csrc -lmcrat_b200 -lmcrat_b200_io \
 *            -Wl,-rpath,$PWD/mcrat_b200/csrc -o host_multi_gpu
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mcrat_b200.h"
#include "mcrat_b200_io.h"

typedef struct {
    int rank, nranks, frames;
    long long max_iters;
    const char *dir;
    const mcrat_b200_config *cfg;
    const mcrat_b200_io_switches *sw;
    const mcrat_b200_mc_par *par;
    int ncells;
    const double *const *fields;
    const double *domains;
    const mcrat_photon *photons; /* the whole list; this rank takes [first, first + count) */
    int first, count;
    unsigned char *id;           /* MCRAT_B200_COMM_ID_BYTES, written by rank 0 */
    pthread_barrier_t *barrier;
    int rc;
} rank_args;

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

#define CHECK(call)                                                                                  \
    do {                                                                                             \
        if ((call) != 0) {                                                                           \
            fprintf(stderr, "rank %d: %s: %s\n", a->rank, #call, mcrat_b200_last_error(ctx));        \
            exit(1); /* the other ranks would wait for this one in the next collective: MPI_Abort */ \
        }                                                                                            \
    } while (0)

static void *rank_main(void *arg)
{
    rank_args *a = (rank_args *)arg;
    mcrat_b200_ctx *ctx = NULL;
    mcrat_b200_comm *comm = NULL;
    mcrat_b200_config cfg = *a->cfg;
    cfg.device = a->rank;            /* one GPU per rank */
    cfg.shard = (uint32_t)a->rank;   /* mixed into the Philox key: ranks draw independent streams */
    cfg.seed = 20261018;
    cfg.rng_mode = MCRAT_RNG_PHILOX;
    if (mcrat_b200_create(&cfg, &ctx) != 0) {
        fprintf(stderr, "rank %d: create: %s\n", a->rank, mcrat_b200_last_error(NULL));
        a->rc = 1;
        /* the other ranks would wait for this one in the communicator: a real host aborts the job here (MPI_Abort) */
        exit(1);
    }
    if (a->rank == 0 && mcrat_b200_comm_unique_id(a->id, MCRAT_B200_COMM_ID_BYTES) != 0) {
        fprintf(stderr, "rank 0: comm_unique_id: %s\n", mcrat_b200_last_error(NULL));
        exit(1);
    }
    pthread_barrier_wait(a->barrier); /* MPI_Bcast(id, ...) in an MPI host */
    CHECK(mcrat_b200_comm_create(ctx, a->nranks, a->rank, a->id, MCRAT_B200_COMM_ID_BYTES, &comm));
    if (cfg.tau_calculation == MCRAT_TABLE) CHECK(mcrat_b200_comm_build_thermal_table(comm, 500000, 1, NULL, NULL));

    CHECK(mcrat_b200_set_photons(ctx, a->photons + a->first, a->count));
    const int frame0 = a->par->frm0[0];
    double time_now = (double)frame0 / a->par->fps;
    for (int k = 0; k < a->frames; ++k) {
        const int frame = frame0 + k;
        /* getHydroData(&hydrodata, ...) would run here (Src/mcrat.c:721); the same frame is re-used */
        CHECK(mcrat_b200_set_hydro(ctx, a->ncells, a->fields, a->domains, a->par->fps, frame, frame0));
        const double remaining_time = ((double)(frame + 1) / a->par->fps) - time_now; /* Src/mcrat.c:754 */
        mcrat_b200_frame_stats st, job;
        CHECK(mcrat_b200_run_frame(ctx, time_now, remaining_time, a->max_iters, 1, &st));
        CHECK(mcrat_b200_comm_reduce_frame_stats(comm, &st, &job)); /* collective: every rank, every frame */
        if (a->rank == 0)
            printf("frame %d: %lld scatterings, %lld re-locations, %lld photon-iterations over %d GPUs (slowest rank: %lld iterations)\n",
                   frame, job.scatterings, job.relocations, job.photon_slots, a->nranks, job.iterations);
        time_now = st.time_now;
    }

    /* the merged output: all photons on rank 0, in rank order */
    long long *counts = (long long *)calloc((size_t)a->nranks, sizeof(long long)), total = 0;
    CHECK(mcrat_b200_comm_photon_counts(comm, counts, NULL, NULL)); /* collective: list lengths of all ranks */
    for (int r = 0; r < a->nranks; ++r) total += counts[r];
    mcrat_photon *all = a->rank == 0 ? (mcrat_photon *)malloc((size_t)(total ? total : 1) * sizeof(mcrat_photon)) : NULL;
    CHECK(mcrat_b200_comm_gather_photons(comm, 0, all, a->rank == 0 ? total : 0, counts, &total));
    if (a->rank == 0) {
        const int last = a->par->frm0[0] + a->frames - 1, ranks[1] = {0};
        if (mcrat_b200_print_photons(a->dir, 0, last, all, (int)total, a->sw) != MCRAT_IO_OK ||
            mcrat_b200_merge_frame(a->dir, last, ranks, 1, a->sw) != MCRAT_IO_OK) {
            fprintf(stderr, "output: %s\n", mcrat_b200_io_last_error());
            a->rc = 1;
        } else {
            printf("wrote %lld photons of %d ranks to %s/mcdata_%d.h5\n", total, a->nranks, a->dir, last);
        }
    }
    free(all);
    free(counts);
    mcrat_b200_comm_destroy(comm);
    mcrat_b200_destroy(ctx);
    return NULL;
}

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s <dir> <gpus> [frames] [max_iters]\n", argv[0]);
        return 2;
    }
    const char *dir = argv[1];
    const int gpus = atoi(argv[2]);
    const int frames = argc > 3 ? atoi(argv[3]) : 2;
    const long long max_iters = argc > 4 ? atoll(argv[4]) : 200;
    if (gpus < 1 || gpus > 64 || gpus > mcrat_b200_device_count()) {
        fprintf(stderr, "%d GPUs asked for, %d CUDA devices present (there is no CPU fallback)\n", gpus, mcrat_b200_device_count());
        return 2;
    }
    char path[1024];
    mcrat_b200_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    mcrat_b200_io_switches sw;
    mcrat_b200_mc_par par;
    snprintf(path, sizeof(path), "%s/mcrat_input.h", dir);
    if (mcrat_b200_config_from_input_header(path, &cfg, &sw) != MCRAT_IO_OK) {
        fprintf(stderr, "%s\n", mcrat_b200_io_last_error());
        return 2;
    }
    snprintf(path, sizeof(path), "%s/%s", dir, sw.mcpar);
    if (mcrat_b200_read_mc_par(path, &par) != MCRAT_IO_OK) {
        fprintf(stderr, "%s\n", mcrat_b200_io_last_error());
        return 2;
    }
    size_t hbytes = 0, pbytes = 0;
    snprintf(path, sizeof(path), "%s/hydro.bin", dir);
    char *hb = (char *)read_all(path, &hbytes);
    snprintf(path, sizeof(path), "%s/photons.bin", dir);
    char *pb = (char *)read_all(path, &pbytes);
    int ncells = 0, nph = 0;
    memcpy(&ncells, hb, 4);
    memcpy(&nph, pb, 4);
    if (hbytes != 8 + (size_t)19 * (size_t)ncells * 8 || pbytes != 8 + (size_t)nph * sizeof(mcrat_photon)) {
        fprintf(stderr, "hydro.bin / photons.bin have unexpected sizes\n");
        return 2;
    }
    const double *cols = (const double *)(hb + 8);
    const double *fields[19];
    for (int f = 0; f < 19; ++f) fields[f] = cols + (size_t)f * (size_t)ncells;
    const double domains[6] = {par.r0_domain[0], par.r0_domain[1], par.r1_domain[0], par.r1_domain[1], par.r2_domain[0], par.r2_domain[1]};
    const mcrat_photon *photons = (const mcrat_photon *)(pb + 8);

    unsigned char id[MCRAT_B200_COMM_ID_BYTES];
    pthread_barrier_t barrier;
    pthread_barrier_init(&barrier, NULL, (unsigned)gpus);
    pthread_t *th = (pthread_t *)calloc((size_t)gpus, sizeof(pthread_t));
    rank_args *ra = (rank_args *)calloc((size_t)gpus, sizeof(rank_args));
    const int base = nph / gpus, extra = nph % gpus; /* contiguous, near-equal slices: rank r gets the r-th */
    int rc = 0;
    for (int r = 0; r < gpus; ++r) {
        rank_args *a = &ra[r];
        a->rank = r; a->nranks = gpus; a->frames = frames; a->max_iters = max_iters; a->dir = dir;
        a->cfg = &cfg; a->sw = &sw; a->par = &par; a->ncells = ncells; a->fields = fields; a->domains = domains;
        a->photons = photons;
        a->first = r * base + (r < extra ? r : extra);
        a->count = base + (r < extra ? 1 : 0);
        a->id = id; a->barrier = &barrier;
        pthread_create(&th[r], NULL, rank_main, a);
    }
    for (int r = 0; r < gpus; ++r) {
        pthread_join(th[r], NULL);
        rc |= ra[r].rc;
    }
    free(th);
    free(ra);
    free(hb);
    free(pb);
    return rc;
}
