/* mcrat_b200_io.h -- host-side configuration and output surface of the hot path (plain C, no CUDA).
 *
 * the callers and data formats either side of the device path.
 *   - mc.par                       readMcPar,      Src/mcrat_io.c:1136-1237 (file: sample_mc.par:1-25)
 *   - mcrat_input.h                the compile-time switches of the reference (Src/mcrat_input.h,
 *                                  defaults Src/mcrat.h:262-427) -> run-time mcrat_b200_config
 *   - mc_proc_<rank>.h5            printPhotons,   Src/mcrat_io.c:113-530: one group per hydro frame,
 *                                  datasets P0..P3 [COMV_P0..3] R0..R2 [S0..S3] [PT] NS PW
 *   - mcdata_<frame>.h5            dirFileMerge,   Src/mcrat_io.c:1239-1770: the same names at the file
 *                                  root, ranks concatenated in id order
 *
 * HDF5 is written and read by a self-contained implementation of the subset of the HDF5 file
 * format the reference's output uses (superblock version 0, version-1 object headers, symbol-table
 * groups: B-tree v1 + local heap + SNOD; IEEE F64LE and STD_I8LE datatypes; contiguous layout on
 * write, contiguous and un-filtered chunked layout on read).  No libhdf5 is needed, and files
 * written here open with libhdf5 / h5py / ProcessMCRaT.  Differences from the reference's files are
 * confined to storage layout: datasets are contiguous (the reference's per-rank files use chunked,
 * extendible datasets), and appending to a frame group rewrites the file.
 */
#ifndef MCRAT_B200_IO_H
#define MCRAT_B200_IO_H

#include <stddef.h>
#include "mcrat_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MCRAT_IO_OK 0
#define MCRAT_IO_ERR_OPEN (-11)   /* cannot open / create the file */
#define MCRAT_IO_ERR_FORMAT (-12) /* not in the supported HDF5 subset, or malformed text input */
#define MCRAT_IO_ERR_ARG (-13)
#define MCRAT_IO_ERR_NOTFOUND (-14)
#define MCRAT_IO_ERR_NOMEM (-15)

#define MCRAT_IO_MAX_ANGLE_BINS 64

/* ---- mc.par (readMcPar, Src/mcrat_io.c:1136-1237) ---------------------------------------- */
typedef struct mcrat_b200_mc_par {
    double fps;                 /* hydro_data->fps */
    int last_frame;             /* hydro_data->last_frame */
    double r0_domain[2], r1_domain[2], r2_domain[2];
    double theta_jmin, theta_j; /* degrees, as the reference keeps them */
    int n_theta_j;              /* number of angle bins */
    int frm0[MCRAT_IO_MAX_ANGLE_BINS];       /* first injection frame per bin */
    int frm2[MCRAT_IO_MAX_ANGLE_BINS];       /* frm0 + number of injection frames (Src/mcrat_io.c:1201) */
    double inj_radius[MCRAT_IO_MAX_ANGLE_BINS];
    char spect;                 /* 'w' wien, 'b' blackbody */
    int min_photons, max_photons;
    char restart;               /* 'i' initialise, 'c' continue */
} mcrat_b200_mc_par;

int mcrat_b200_read_mc_par(const char *path, mcrat_b200_mc_par *out);

/* ---- mcrat_input.h -> mcrat_b200_config ----------------------------------------------------- */
/* output / bookkeeping switches that do not change device arithmetic */
typedef struct mcrat_b200_io_switches {
    int comv_switch;     /* COMV_SWITCH  (default OFF, Src/mcrat.h) */
    int save_type;       /* SAVE_TYPE    (default OFF; forced ON with CYCLOSYNCHROTRON_SWITCH) */
    int stokes_switch;   /* copy of cfg->stokes_switch */
    int sim_switch;      /* FLASH 0, PLUTO_CHOMBO 1, PLUTO 2 (Src/mcrat.h:17-20) */
    int simulation_type; /* SCIENCE 0, CYLINDRICAL_OUTFLOW 1, SPHERICAL_OUTFLOW 2, STRUCTURED_SPHERICAL_OUTFLOW 3 */
    char mc_path[256], filepath[256], fileroot[256], mcpar[64];
} mcrat_b200_io_switches;

/* Parses the `#define NAME VALUE` lines of a mcrat_input.h and fills the run-time configuration with
 * the reference's codes (Src/mcrat.h:17-65) and defaults (Src/mcrat.h:262-427).  Fields of `cfg`
 * that are not compile-time switches of the reference (device, seed, rng_mode, ...) are left as
 * they are; cfg->abi_version is set.  Returns MCRAT_IO_ERR_FORMAT for combinations the reference
 * rejects with #error. */
int mcrat_b200_config_from_input_header(const char *path, mcrat_b200_config *cfg, mcrat_b200_io_switches *sw);

/* ---- photon output ---------------------------------------------------------------------------- */
/* printPhotons (Src/mcrat_io.c:113): appends the photons with weight != 0 to group "<frame>" of
 * <dir>/mc_proc_<angle_rank>.h5 (created if missing; the group is created or extended). */
int mcrat_b200_print_photons(const char *dir, int angle_rank, int frame, const mcrat_photon *photons, int list_capacity,
                             const mcrat_b200_io_switches *sw);

/* dirFileMerge for one frame (Src/mcrat_io.c:1239): concatenates group "<frame>" of
 * <dir>/mc_proc_<id>.h5, id = ranks[0..nranks), into <dir>/mcdata_<frame>.h5 (datasets at the root).
 * Ranks whose file has no such group are skipped, as in the reference. */
int mcrat_b200_merge_frame(const char *dir, int frame, const int *ranks, int nranks, const mcrat_b200_io_switches *sw);

/* ---- generic access to the HDF5 subset (tests, CONTINUE, interoperability) ------------------- */
/* Number of elements of dataset `name` ("P0", or "200/P0" inside a group); < 0 on error. */
long long mcrat_b200_h5_dataset_length(const char *path, const char *name);
/* Reads a 1-D dataset: F64 into `out_f64` (n doubles) or I8 into `out_i8` (n chars); the other may be NULL. */
int mcrat_b200_h5_read_dataset(const char *path, const char *name, double *out_f64, signed char *out_i8, size_t n);
/* Lists the links of a group ("" or "/" = root) as a '\n'-separated string; returns the count or < 0. */
int mcrat_b200_h5_list(const char *path, const char *group, char *buf, size_t buflen);

const char *mcrat_b200_io_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
