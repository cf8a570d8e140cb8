/*
 * mcrat_b200.h -- C ABI of the B200-native MCRaT photon-propagation / scattering hot path.
 *
 * One shared library (libmcrat_b200.so: hand-written sm_100a CUDA, FP64) that replaces
 * the body of MCRaT's scatter-frame while-loop, Src/mcrat.c:761-851.  Plain pointers and
 * sizes only; the host stays C.  Every entry point names the reference interface it
 * replaces (file:line under the reference tree).  Photon records cross the boundary in
 * the reference's own `struct photon` layout (Src/mcrat.h:142-171, 176 bytes) and cell
 * data as the `struct hydro_dataframe` arrays (Src/mcrat.h:194-244); on the device both
 * live as SoA columns.
 *
 * There is no CPU fallback: every call returns MCRAT_B200_ERR_CUDA if the device path
 * cannot run.  All functions return 0 on success, a negative MCRAT_B200_ERR_* otherwise;
 * mcrat_b200_last_error() gives the message.
 *
 * Thread-safety: one context per host thread / MPI rank / GPU (the reference runs one
 * rank per shard, Src/mcrat.c:139-164).  A context owns one CUDA stream.
 */
#ifndef MCRAT_B200_H
#define MCRAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCRAT_B200_ABI_VERSION 2

/* reference codes, Src/mcrat.h:36-65 */
enum { MCRAT_CARTESIAN = 0, MCRAT_SPHERICAL = 1, MCRAT_CYLINDRICAL = 2, MCRAT_POLAR = 3 };
enum { MCRAT_TWO = 0, MCRAT_TWO_POINT_FIVE = 1, MCRAT_THREE = 2 };
enum { MCRAT_INTERNAL_E = 0, MCRAT_TOTAL_E = 1, MCRAT_SIMULATION = 2 };
enum { MCRAT_DIRECT = 1, MCRAT_TABLE = 2 };

enum { MCRAT_RNG_PHILOX = 0, MCRAT_RNG_REPLAY = 1 };

enum {
    MCRAT_B200_OK = 0,
    MCRAT_B200_ERR_CUDA = -1,     /* CUDA runtime / no device */
    MCRAT_B200_ERR_ARG = -2,      /* bad argument */
    MCRAT_B200_ERR_STATE = -3,    /* call out of order (e.g. no hydro frame loaded) */
    MCRAT_B200_ERR_REPLAY = -4,   /* replay uniform buffer exhausted */
    MCRAT_B200_ERR_TABLE = -5     /* hot cross-section lookup outside the table */
};

/* `struct photon`, Src/mcrat.h:142-171 with NONTHERMAL_E_DIST == OFF.  Same member order
 * and types, hence the same 176-byte layout the reference dumps into checkpoints
 * (Src/mcrat_io.c:902). */
typedef struct mcrat_photon {
    char type;
    double p0, p1, p2, p3;
    double comv_p0, comv_p1, comv_p2, comv_p3;
    double r0, r1, r2;
    double s0, s1, s2, s3;
    double num_scatt;
    int recalc_properties;
    double weight;
    int nearest_block_index;
    double time_to_scatter;
    double total_optical_depth;
} mcrat_photon;

/* The reference's compile-time switches (Src/mcrat_input.h, Src/mcrat.h:262-427) as an
 * init-time struct, plus device selection. */
typedef struct mcrat_b200_config {
    int abi_version;       /* MCRAT_B200_ABI_VERSION */
    int dimensions;        /* DIMENSIONS */
    int geometry;          /* GEOMETRY */
    int stokes_switch;     /* STOKES_SWITCH */
    int tau_calculation;   /* TAU_CALCULATION */
    int cyclosynch_switch; /* CYCLOSYNCHROTRON_SWITCH */
    int b_field_calc;      /* B_FIELD_CALC */
    double epsilon_b;      /* EPSILON_B */
    int device;            /* CUDA device ordinal */
    int rng_mode;          /* MCRAT_RNG_PHILOX (production) or MCRAT_RNG_REPLAY (parity harness) */
    uint64_t seed;         /* Philox key, low/high words */
    uint32_t shard;        /* shard (rank) id mixed into the Philox key */
    int profile;           /* 1: time each kernel class with CUDA events (see kernel_times) */
    void *stream;          /* cudaStream_t to run on, or NULL to let the library create one */
    int scan_index;        /* 0: the rescan of a new hydro frame (find_nearest_grid_switch = 1) and the streamed
                            *    loop re-locate with the full photon x cell scan of the reference (K1 / K1b);
                            * 1: the same first-match search through a bounding-box index over the
                            *    cells in array order (identical results, far fewer tests).
                            * The persistent loop always re-locates the few photons that change cell in a
                            * steady-state iteration through the index (one warp per photon). */
} mcrat_b200_config;

typedef struct mcrat_b200_ctx mcrat_b200_ctx;

/* what one call of the device-resident frame loop did */
typedef struct mcrat_b200_frame_stats {
    long long iterations;   /* while-loop iterations executed, Src/mcrat.c:761 (max over sub-shards) */
    long long scatterings;  /* frame_scatt_cnt, Src/mclib.c:1318 */
    long long relocations;  /* num_photons_find_new_element, Src/mcrat.c:768 */
    long long photon_slots; /* sum over iterations of list_capacity (photon-iterations) */
    long long cell_evals;   /* photon-cell containment tests executed by the scan kernels */
    long long box_evals;    /* photon-box tests executed by the bounding-box index (scan_index = 1) */
    double time_now;        /* clock of sub-shard 0 */
    double last_time_step;
    int last_scattered_index;
    int not_found;          /* photons for which no containing cell exists (Src/mclib.c:583) */
    int cs_host_pending;    /* loop paused for the host: 1 = a pool photon scattered and the list has no null slot
                             * left (host grows the list and emits, Src/mcrat.c:792-807); 2 = rebin due (:820-830) */
    int error;              /* 0 or MCRAT_B200_ERR_* raised on the device */
    int cs_emitted;         /* pool photons replaced on the device (photonEmitCyclosynch, single mode) */
    int scatt_cyclosynch_num_ph;      /* running scatt_cyclosynch_num_ph, Src/mcrat.c:803 */
    double cs_comptonized_weight;     /* n_comptonized added in this call, Src/mcrat.c:794 */
    long long ref_equiv_evals; /* checkInBlock calls the reference's early-exit loop (Src/geometry.c:356-370) would have made
                                * for the photons re-located by full rescans in this call: first-hit index + 1 each, or
                                * num_elements for a photon without a containing cell.  cell_evals counts what the device
                                * executed (every photon x every cell for K1). */
} mcrat_b200_frame_stats;

typedef struct mcrat_b200_kernel_times {
    double scan_ms;   long long scan_launches;   /* K1 photon x cell containment scan */
    double pass_ms;   long long pass_launches;   /* K4+K2 fused push / re-check / free-path / block arg-min */
    double event_ms;  long long event_launches;  /* K3 fused scatter */
    double other_ms;  long long other_launches;
} mcrat_b200_kernel_times;

/* ---- life cycle ---------------------------------------------------------------------- */
int mcrat_b200_abi_version(void);
int mcrat_b200_device_count(void);
int mcrat_b200_create(const mcrat_b200_config *cfg, mcrat_b200_ctx **out);
void mcrat_b200_destroy(mcrat_b200_ctx *ctx);
const char *mcrat_b200_last_error(const mcrat_b200_ctx *ctx); /* ctx may be NULL: last create() error */
int mcrat_b200_synchronize(mcrat_b200_ctx *ctx);

/* ---- inputs --------------------------------------------------------------------------- */
/* Upload one hydro frame (replaces what getHydroData leaves in `struct hydro_dataframe`,
 * Src/mcrat.c:721).  `fields` holds 19 host pointers in struct order: r0 r1 r2 r0_size
 * r1_size r2_size r theta v0 v1 v2 dens dens_lab pres temp gamma B0 B1 B2 (unused ones
 * may be NULL); `domains` = {r0_domain[2], r1_domain[2], r2_domain[2]}. */
int mcrat_b200_set_hydro(mcrat_b200_ctx *ctx, int num_elements, const double *const *fields, const double *domains,
                         double fps, int scatt_frame_number, int inj_frame_number);
/* thermal_table[N_PH_E+1][N_T+1] of Src/hot_x_section.c:15 (221 x 81, log10 sigma/sigma_T) */
int mcrat_b200_set_thermal_table(mcrat_b200_ctx *ctx, const double *table);
/* createHotCrossSection (Src/hot_x_section.c:82-206) on the device: every table point is the
 * reference's plain Monte Carlo integral with `calls` samples (the reference uses 500000,
 * :348) from a Philox stream keyed by the point; the table is installed in the context and, if
 * table_out != NULL, copied out in the reference's [N_PH_E+1][N_T+1] order (write it with the
 * layout of Src/hot_x_section.c:116-131 to interoperate with thermal_hot_x_section.dat). */
int mcrat_b200_build_thermal_table(mcrat_b200_ctx *ctx, long long calls, uint64_t seed, double *table_out,
                                   float *elapsed_ms);
/* Upload / download the photon list (`struct photonList`.photons, Src/mcrat.h:173-180). */
int mcrat_b200_set_photons(mcrat_b200_ctx *ctx, const mcrat_photon *photons, int list_capacity);
int mcrat_b200_get_photons(mcrat_b200_ctx *ctx, mcrat_photon *photons, int list_capacity);
int mcrat_b200_get_photon(mcrat_b200_ctx *ctx, int index, mcrat_photon *out);
int mcrat_b200_list_capacity(const mcrat_b200_ctx *ctx);
/* Sub-shards per context.  The reference decomposes a run into independent ranks, each with its
 * own photon list, clock and time-ordered scatter sequence (Src/mcrat.c:139-164, 457-479; no
 * exchange inside the frame loop).  A context can hold `num_shards` such ranks at once: the next
 * mcrat_b200_set_photons() splits the list into that many contiguous, equal slot ranges, and the
 * frame loop then advances every sub-shard concurrently (one scattering event per sub-shard and
 * iteration).  Sub-shard s behaves exactly like a stand-alone context created with
 * shard = cfg.shard + s whose list is that slot range (same Philox streams, same results).
 * Default 1.  The step-by-step surface and the replay harness need num_shards == 1. */
int mcrat_b200_set_num_shards(mcrat_b200_ctx *ctx, int num_shards);
int mcrat_b200_num_shards(const mcrat_b200_ctx *ctx);
/* parity harness: the uniform stream the reference's gsl_rng would hand out */
int mcrat_b200_set_replay_uniforms(mcrat_b200_ctx *ctx, const double *u, size_t n);
long long mcrat_b200_replay_consumed(mcrat_b200_ctx *ctx);

/* ---- the reference's function surface ----------------------------- */
/* findContainingHydroCell, Src/mclib.h:8, Src/mclib.c:436 */
int mcrat_b200_find_containing_hydro_cell(mcrat_b200_ctx *ctx, int find_nearest_block_switch,
                                          int *num_photons_find_new_element);
/* calcMeanFreePath, Src/mclib.h:10, Src/mclib.c:617.  Returns the head of the time-ordered
 * list: sorted_indexes[0] and its time_to_scatter (what Src/mcrat.c:777 reads). */
int mcrat_b200_calc_mean_free_path(mcrat_b200_ctx *ctx, int *first_index, double *first_time_to_scatter);
/* The rest of photonList.sorted_indexes, for hosts that read more than its head: the slot indices 0 .. n-1 ordered by the
 * time_to_scatter the last calcMeanFreePath left (Src/mclib.c:717-729), sorted on the device.  Equal times -- the 1e12 / c of
 * every photon outside the domain -- come in ascending slot order (the reference's qsort_r leaves their order to the C
 * library).  Call after mcrat_b200_calc_mean_free_path; single shard. */
int mcrat_b200_get_sorted_indexes(mcrat_b200_ctx *ctx, int *sorted_indexes, int n);
/* photonEvent, Src/mclib.h:23, Src/mclib.c:1107 */
int mcrat_b200_photon_event(mcrat_b200_ctx *ctx, double dt_max, double *time_step, int *scattered_ph_index,
                            int *frame_scatt_cnt, int *frame_abs_cnt);
/* updatePhotonPosition, Src/mclib.h:21, Src/mclib.c:1054 */
int mcrat_b200_update_photon_position(mcrat_b200_ctx *ctx, double t);
/* phAbsCyclosynch, Src/mc_cyclosynch.h:92, Src/mc_cyclosynch.c:1571 */
int mcrat_b200_ph_abs_cyclosynch(mcrat_b200_ctx *ctx, int *num_abs_ph, int *scatt_cyclosynch_num_ph,
                                 double *absorbed_weight);
/* calcCyclosynchRLimits, Src/mc_cyclosynch.h:84, Src/mc_cyclosynch.c:225; min_or_max is "min" or "max" */
double mcrat_b200_calc_cyclosynch_r_limits(int frame_scatt, int frame_inj, double fps, double r_inj,
                                           const char *min_or_max);
/* rebin threshold (mc.par max_photons) and the running scatt_cyclosynch_num_ph of the driver
 * (Src/mcrat.c:803, 820): with CYCLOSYNCHROTRON_SWITCH ON the frame loop replaces scattered pool
 * photons on the device (photonEmitCyclosynch in single mode, Src/mc_cyclosynch.c:1465-1555) and
 * pauses for the host only to grow the list or to rebin */
int mcrat_b200_set_cs_limits(mcrat_b200_ctx *ctx, int max_photons, int scatt_cyclosynch_num_ph);
/* rebinCyclosynchCompPhotons, Src/mc_cyclosynch.h:86, Src/mc_cyclosynch.c:600-710, on the device: every 'k' / 'c' photon
 * is replaced by the weighted-mean photon of its (log E, theta[, phi]) bin, placed into the null slots of the list in slot
 * order (addToPhotonList, Src/photons.c:167-205).  The counters are the ones the reference updates (:680-684); the return
 * value is MCRAT_B200_OK or an error (the reference returns the number of empty bins, given here through
 * num_null_rebin_ph).  If the list has fewer null slots than bins the call fails with MCRAT_B200_ERR_STATE and the host
 * must grow the list.  Single shard only (CYCLOSYNCHROTRON_SWITCH ON). */
int mcrat_b200_rebin_cyclosynch_comp_photons(mcrat_b200_ctx *ctx, int max_photons, int *num_cyclosynch_ph_emit,
                                             int *scatt_cyclosynch_num_ph, int *num_null_rebin_ph);
/* photonEmitCyclosynch with inject_single_switch == 0, Src/mc_cyclosynch.h:90, Src/mc_cyclosynch.c:1176-1464, on the device
 * (K6): cells of the emission shell of the hydro frame last uploaded (its fps, scatt_frame_number and inj_frame_number
 * give rmin / rmax, :1206-1207), photon-weight search with Poisson counts per cell, the new 'p' photons placed into the
 * null slots of the list in slot order (addToPhotonList).  A list with too few null slots grows on the device
 * (mcrat_b200_list_capacity changes; download accordingly).  Outputs: the return value of the reference's function
 * (photons emitted), the weight the search settled on, the number of cells in the shell.  Single shard.
 * (inject_single_switch == 1, the replacement of one scattered pool photon, happens inside the frame loop.) */
int mcrat_b200_photon_emit_cyclosynch(mcrat_b200_ctx *ctx, double r_inj, double ph_weight, int maximum_photons, double theta_min,
                                      double theta_max, int *num_emitted, double *ph_weight_adjusted, int *num_cells_selected);
/* photonEmitCyclosynch with inject_single_switch == 1 for hosts that drive the loop call by call (Src/mcrat.c:792-803, right
 * after the photonEvent that scattered pool photon `scatt_ph_index`): the photon becomes a comptonised one ('k'), a fresh
 * pool photon goes into the first null slot (*new_photon_index; the list grows when it has none) and the scattered photon
 * is re-positioned inside its cell.  Draws continue the event's stream.  mcrat_b200_run_frame does all this by itself. */
int mcrat_b200_photon_emit_cyclosynch_single(mcrat_b200_ctx *ctx, int scatt_ph_index, int *new_photon_index);
/* CYCLOSYNCHROTRON_REBIN_E_PERC, _REBIN_ANG (degrees), _REBIN_ANG_PHI (degrees); defaults 0.1, 0.5, 10 (Src/mcrat.h:308-322) */
int mcrat_b200_set_cs_rebin_params(mcrat_b200_ctx *ctx, double rebin_e_perc, double rebin_ang_deg, double rebin_ang_phi_deg);
/* phMinMax / phScattStats / averagePhotonEnergy, Src/mclib.c:1465 / 1385 / 1358 */
int mcrat_b200_ph_min_max(mcrat_b200_ctx *ctx, double *min_r, double *max_r, double *min_theta, double *max_theta);
int mcrat_b200_ph_scatt_stats(mcrat_b200_ctx *ctx, int *max_scatt, int *min_scatt, double *avg_scatt, double *avg_r);
int mcrat_b200_average_photon_energy(mcrat_b200_ctx *ctx, double *avg_energy);

/* ---- device-resident frame loop ------------------------------------------------------------- */
/* The whole `while (remaining_time > 0)` loop of Src/mcrat.c:761-851 on the device.
 * Stops when remaining_time reaches 0, after max_iters iterations (max_iters < 0: no cap),
 * or when the host must act (stats->cs_host_pending, stats->error). */
int mcrat_b200_run_frame(mcrat_b200_ctx *ctx, double time_now, double remaining_time, long long max_iters,
                         int find_nearest_grid_switch, mcrat_b200_frame_stats *stats);

/* The reference re-checks every photon's cached cell in every iteration (findContainingHydroCell, Src/mclib.c:469-597).
 * The pass kernel skips that re-check for a photon while the total length of the pushes since its last re-check is
 * provably smaller than its distance to the boundary of its cell and of the domain (a per-shard path counter against
 * a per-photon threshold); a skipped re-check would have succeeded and has no side effect, so results are identical.
 * mode 0: never skip; 1 (default): skip; 2: decide as in 1 but re-check anyway and fail the frame with
 * MCRAT_B200_ERR_STATE if a skipped photon had left its cell (test mode). */
int mcrat_b200_set_recheck_skip(mcrat_b200_ctx *ctx, int mode);
/* How the loop is driven.
 * STREAMED: stream-ordered kernel launches per iteration, enqueued in growing batches.  With two or more sub-shards
 *   an iteration is two launches per half of the sub-shards (pass; re-locate + event), the halves on two streams half a
 *   period apart, so that one half's scattering events run beside the other half's pass and the HBM pipe never waits
 *   for a scattering.  New hydro frames (everything re-locates) and optically thin flows (many photons change cell per
 *   iteration) go through the grid-wide re-location kernels: pass, K1 / K1b / K1c, finish, event.
 * PERSISTENT: one cooperative launch per frame; every sub-shard is iterated by resident blocks that hand over through
 *   a generation word in global memory (no launch boundary inside the loop).
 * AUTO (default): PERSISTENT_STREAM once the photon columns outgrow L2 (>= 1.2 x 10^6 photons) and there are 16 or more
 *   sub-shards of at most 200 000 photons; otherwise PERSISTENT up to 2^21 photons and STREAMED above.
 * In all of them sub-shards advance independently, like MPI ranks, and the photons are bit-identical; the replay
 * harness always runs STREAMED_GLOBAL. */
#define MCRAT_B200_LOOP_AUTO 0
#define MCRAT_B200_LOOP_STREAMED 1
#define MCRAT_B200_LOOP_PERSISTENT 2
#define MCRAT_B200_LOOP_STREAMED_GLOBAL 3 /* streamed, every iteration through the grid-wide re-location kernels (four launches;
                                          * what STREAMED falls back to for new hydro frames and optically thin flows) */
#define MCRAT_B200_LOOP_PERSISTENT_STREAM 4 /* lists larger than L2: resident event blocks (each serving a few sub-shards
                                           * in turn) beside resident pass blocks that pull (iteration, sub-shard, slice)
                                           * items from a counter -- the PERSISTENT protocol without tying pass blocks to a
                                           * shard, so the photon columns stream through the SMs without a launch boundary
                                           * while the scatterings run beside them.  AUTO: see above. */
int mcrat_b200_set_loop_mode(mcrat_b200_ctx *ctx, int mode);

/* per-sub-shard view of the counters (cumulative since the shard layout was set) and its slot range */
int mcrat_b200_get_shard_stats(mcrat_b200_ctx *ctx, int shard, mcrat_b200_frame_stats *stats, int *first_slot,
                               int *num_slots);

/* findContainingBlock logs every photon it finds no cell for with its hydro coordinates (Src/geometry.c:373-388).  The
 * device keeps the first 32 of them since the last call: slot index and (r0, r1, r2) in the hydro coordinate system.
 * Returns the number of entries copied (<= max_entries), *n_total = photons not found since the last call; the log is
 * cleared.  frame_stats.not_found counts them per frame. */
int mcrat_b200_get_not_found(mcrat_b200_ctx *ctx, int max_entries, int *slots, double *hydro_coords, int *n_total);

/* ---- measurement ------------------------------------------------------------------------------ */
int mcrat_b200_get_kernel_times(mcrat_b200_ctx *ctx, mcrat_b200_kernel_times *out, int reset);
/* switch per-kernel-class timing (mcrat_b200_config.profile) on or off for the following calls; while it is on every
 * launch is bracketed by CUDA events and the loop runs streamed */
int mcrat_b200_set_profile(mcrat_b200_ctx *ctx, int on);
/* number of kernels launched through this context so far */
long long mcrat_b200_launch_count(const mcrat_b200_ctx *ctx);
/* full photon x cell rescan only (the K1 kernel on the current list), for roofline timing */
int mcrat_b200_rescan_all(mcrat_b200_ctx *ctx, long long *cell_evals, float *elapsed_ms);
/* sustained FP64-pipe instruction issue rate of this GPU (independent DFMA chains on all SMs),
 * G thread-instructions/s; x2 = DFMA GFLOP/s */
int mcrat_b200_measure_fp64_peak(mcrat_b200_ctx *ctx, double *ginstr_per_s);
/* self-test: the free path divides by C_LIGHT (Src/mclib.c:684) with an FMA-corrected multiplication by the
 * reciprocal; this compares it bit for bit with the general division on 4 x n doubles and returns the number of
 * differing results (0 expected) */
int mcrat_b200_selftest_div_by_c(mcrat_b200_ctx *ctx, long long n, unsigned seed, long long *mismatches);
/* streaming copy bandwidth of this GPU, GB/s (read+write bytes) */
int mcrat_b200_measure_hbm_peak(mcrat_b200_ctx *ctx, double *gb_per_s);

/* ---- multi-GPU: the reference's MPI exchanges either side of the frame loop, over NCCL ------------------ */
/* One process per GPU, one communicator per context.  The frame loop needs no exchange (a GPU's sub-shards are MPI
 * ranks of the reference, Src/mcrat.c:139-164, 457-479); these calls cover what the reference does exchange.  NCCL is
 * bound at run time: without it every comm call fails with MCRAT_B200_ERR_STATE, the rest of the library is unaffected.
 * All calls are collective over the communicator and run on the context's stream. */
typedef struct mcrat_b200_comm mcrat_b200_comm;
#define MCRAT_B200_COMM_ID_BYTES 128
/* ncclGetUniqueId on one rank; the host distributes the bytes (MPI_Bcast in mcrat.c, a store in the test harness) */
int mcrat_b200_comm_unique_id(unsigned char *id, size_t len);
int mcrat_b200_comm_nccl_version(void); /* 0: NCCL not available */
int mcrat_b200_comm_create(mcrat_b200_ctx *ctx, int nranks, int rank, const unsigned char *id, size_t len,
                           mcrat_b200_comm **out);
void mcrat_b200_comm_destroy(mcrat_b200_comm *comm);
int mcrat_b200_comm_rank(const mcrat_b200_comm *comm);
int mcrat_b200_comm_size(const mcrat_b200_comm *comm);
long long mcrat_b200_comm_collectives(const mcrat_b200_comm *comm); /* NCCL operations issued so far */
/* broadcastInterpolationData, Src/hot_x_section.c:709-826 (MPI_Bcast :717): the table installed in `root`'s context
 * replaces the one of every other rank */
int mcrat_b200_comm_bcast_thermal_table(mcrat_b200_comm *comm, int root);
/* createHotCrossSection (Src/hot_x_section.c:82-206) by all GPUs at once: rank r integrates 1/nranks of the table's
 * points (K7), one all-gather assembles and installs the table on every rank.  Same result as
 * mcrat_b200_build_thermal_table with the same (calls, seed), whatever the number of ranks. */
int mcrat_b200_comm_build_thermal_table(mcrat_b200_comm *comm, long long calls, uint64_t seed, double *table_out,
                                        float *elapsed_ms);
/* per-frame counters over all ranks: sums of scatterings, relocations, photon_slots, cell_evals, box_evals,
 * ref_equiv_evals, not_found, cs_emitted, scatt_cyclosynch_num_ph, cs_comptonized_weight; maxima of iterations,
 * time_now, last_time_step, cs_host_pending; error = the most severe (most negative) error of any rank */
int mcrat_b200_comm_reduce_frame_stats(mcrat_b200_comm *comm, const mcrat_b200_frame_stats *mine,
                                       mcrat_b200_frame_stats *total);
/* per-rank photon counts (load balance; what Src/merge.c:784-790 gathers before the merge): list capacity, photons
 * with weight != 0 (the ones printPhotons writes, Src/mcrat_io.c:150-160), null slots; arrays of comm_size entries,
 * any of them may be NULL */
int mcrat_b200_comm_photon_counts(mcrat_b200_comm *comm, long long *list_capacity, long long *output_photons,
                                  long long *null_slots);
/* the photon lists of all ranks concatenated in rank order -- what Src/merge.c:840-876 assembles column by column with
 * MPI_Allgatherv -- packed on the device and moved GPU to GPU; `root` receives them in `photons` (host memory,
 * `capacity` records; root = -1: every rank does), ready for mcrat_b200_print_photons.  counts[r] = records of rank r
 * (may be NULL), *total their sum.  If a receiver's capacity is too small no rank transfers anything, *total says how
 * much is needed and the call returns MCRAT_B200_ERR_ARG everywhere. */
int mcrat_b200_comm_gather_photons(mcrat_b200_comm *comm, int root, mcrat_photon *photons, long long capacity,
                                   long long *counts, long long *total);

#ifdef __cplusplus
}
#endif
#endif
