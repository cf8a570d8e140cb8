/*
 * mcrat_b200_dropin.h -- the reference's own function surface, backed by libmcrat_b200.so.
 *
 * libmcrat_b200_dropin.so defines the functions MCRaT's frame loop calls
 * (Src/mcrat.c:761-851) with exactly the reference's names, argument lists and return
 * values (Src/mclib.h:8-23, Src/mc_cyclosynch.h:92), each exported as `__wrap_<name>` so
 * that an unmodified MCRaT build picks them up with GNU ld's symbol wrapping:
 *
 *     -Wl,--wrap=findContainingHydroCell -Wl,--wrap=calcMeanFreePath
 *     -Wl,--wrap=photonEvent -Wl,--wrap=updatePhotonPosition -Wl,--wrap=averagePhotonEnergy
 *     -Wl,--wrap=phMinMax -Wl,--wrap=phScattStats -Wl,--wrap=phAbsCyclosynch -Wl,--wrap=photonEmitCyclosynch
 *     -Wl,--wrap=rebinCyclosynchCompPhotons -Wl,--wrap=calcCyclosynchRLimits
 *     -Wl,--wrap=initalizeHotCrossSection -Wl,--wrap=cleanupInterpolationData
 *     -lmcrat_b200_dropin -lmcrat_b200
 *
 * (see INTEGRATION.md).  The structs below mirror the member sequence of the reference's
 * `struct photon` / `struct photonList` / `struct hydro_dataframe` (Src/mcrat.h:142-244,
 * NONTHERMAL_E_DIST == OFF) so that pointers to the reference's objects can be passed
 * straight through; `gsl_rng` and `struct SpatialGrid` stay opaque.
 *
 * Mirror policy (who owns the photons when):
 *   findContainingHydroCell(switch=1)  uploads the hydro frame and the photon list
 *                                      (the driver sets switch=1 right after getHydroData,
 *                                      Src/mcrat.c:721,756); with switch=0 works on the device;
 *   calcMeanFreePath                   device; writes back sorted_indexes[0] and that photon's
 *                                      time_to_scatter (all Src/mcrat.c:777 reads);
 *   photonEvent                        device; writes back the scattered photon's record
 *                                      (Src/mcrat.c:787-795, 813); when the frame's time is used
 *                                      up, or with CYCLOSYNCHROTRON_SWITCH ON (host code mutates
 *                                      the list, Src/mcrat.c:799, 825), the whole list;
 *   updatePhotonPosition               device push, then the whole list is downloaded (the driver
 *                                      calls it only as the last step of a frame, Src/mcrat.c:841).
 */
#ifndef MCRAT_B200_DROPIN_H
#define MCRAT_B200_DROPIN_H

#include <stdio.h>

#include "mcrat_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef MCRAT_B200_DROPIN_NO_STRUCTS
/* Src/mcrat.h:173-180 */
struct mcrat_dropin_photonList {
    mcrat_photon *photons;
    int *sorted_indexes;
    int num_photons;
    int num_null_photons;
    int list_capacity;
};

/* Src/mcrat.h:194-244 */
struct mcrat_dropin_hydro_dataframe {
    int num_elements;
    double *r0, *r1, *r2;
    double *r0_size, *r1_size, *r2_size;
    double *r, *theta;
    double *v0, *v1, *v2;
    double *dens, *dens_lab, *pres, *temp, *gamma;
    double *B0, *B1, *B2;
    double r0_domain[2], r1_domain[2], r2_domain[2];
    double fps;
    int scatt_frame_number, inj_frame_number, last_frame, increment_inj_frame, increment_scatt_frame;
    void *grid; /* struct SpatialGrid *, always NULL in the reference (Src/mcrat_io.c:1985) */
};
#endif

typedef struct mcrat_dropin_photonList mcrat_dropin_photonList;
typedef struct mcrat_dropin_hydro_dataframe mcrat_dropin_hydro_dataframe;

/* One-time configuration (the reference's compile-time switches).  If it is never called,
 * the first wrapped call reads the environment variable
 *   MCRAT_B200_CONFIG="dimensions,geometry,stokes,tau_calculation,cyclosynch,b_field_calc,epsilon_b[,device]"
 * (integer codes of Src/mcrat.h:36-65) and aborts with a message if it is absent. */
int mcrat_b200_dropin_configure(const mcrat_b200_config *cfg);
void mcrat_b200_dropin_shutdown(void);
mcrat_b200_ctx *mcrat_b200_dropin_context(void);
/* explicit mirror control for hosts that mutate the list between wrapped calls */
int mcrat_b200_dropin_download(mcrat_dropin_photonList *photon_list);
void mcrat_b200_dropin_mark_host_dirty(void);
/* __wrap_calcMeanFreePath materialises only sorted_indexes[0] (all Src/mcrat.c reads, :777).  A host that walks the order
 * further switches the full sort on (or sets MCRAT_B200_FULL_SORT=1): the whole array is then filled from a stable sort of
 * the time column on the device, equal times in slot order. */
void mcrat_b200_dropin_set_full_sort(int on);

/* the wrapped surface; `rand` (gsl_rng *) is accepted and ignored: the device draws from
 * counter-based Philox streams keyed per photon and iteration */
int __wrap_findContainingHydroCell(mcrat_dropin_photonList *photon_list, mcrat_dropin_hydro_dataframe *hydro_data,
                                   int find_nearest_block_switch, void *rand, FILE *fPtr);
void __wrap_calcMeanFreePath(mcrat_dropin_photonList *photon_list, mcrat_dropin_hydro_dataframe *hydro_data, void *rand,
                             FILE *fPtr);
double __wrap_photonEvent(mcrat_dropin_photonList *photon_list, double dt_max, mcrat_dropin_hydro_dataframe *hydro_data,
                          int *scattered_ph_index, int *frame_scatt_cnt, int *frame_abs_cnt, void *rand, FILE *fPtr);
void __wrap_updatePhotonPosition(mcrat_dropin_photonList *photon_list, double t, FILE *fPtr);
double __wrap_averagePhotonEnergy(mcrat_dropin_photonList *photon_list);
double __wrap_phAbsCyclosynch(mcrat_dropin_photonList *photon_list, int *num_abs_ph, int *scatt_cyclosynch_num_ph,
                              mcrat_dropin_hydro_dataframe *hydro_data, FILE *fPtr);
/* phMinMax / phScattStats, Src/mclib.h:27-29 (Src/mcrat.c:704, 745, 881): device reductions over the mirrored list */
void __wrap_phMinMax(mcrat_dropin_photonList *photon_list, double *min, double *max, double *min_theta, double *max_theta,
                     FILE *fPtr);
void __wrap_phScattStats(mcrat_dropin_photonList *photon_list, int *max, int *min, double *avg, double *r_avg, FILE *fPtr);
/* calcCyclosynchRLimits, Src/mc_cyclosynch.h:84 (Src/mcrat.c:711, 714) */
double __wrap_calcCyclosynchRLimits(int frame_scatt, int frame_inj, double fps, double r_inj, char *min_or_max);
/* rebinCyclosynchCompPhotons, Src/mc_cyclosynch.h:86 (Src/mcrat.c:825, 865): on the device; when the list has fewer null
 * slots than rebinned photons the host list is grown the way addToPhotonList does (Src/photons.c:117-129: realloc to
 * capacity + missing slots, new slots nulled) and the call is repeated.  Returns the number of empty bins. */
int __wrap_rebinCyclosynchCompPhotons(mcrat_dropin_photonList *photon_list, int *num_cyclosynch_ph_emit,
                                      int *scatt_cyclosynch_num_ph, int max_photons, double thread_theta_min,
                                      double thread_theta_max, void *rand, FILE *fPtr);
/* photonEmitCyclosynch, Src/mc_cyclosynch.h:90 (Src/mcrat.c:747 all cells, :799 single): on the device
 * (mcrat_b200_photon_emit_cyclosynch); the host list is grown first when it cannot take the new photons */
int __wrap_photonEmitCyclosynch(mcrat_dropin_photonList *photon_list, double r_inj, double ph_weight, int maximum_photons,
                                double theta_min, double theta_max, mcrat_dropin_hydro_dataframe *hydro_data, void *rand,
                                int inject_single_switch, int scatt_ph_index, FILE *fPtr);
/* initalizeHotCrossSection / cleanupInterpolationData, Src/hot_x_section.h:31, 59 (Src/mcrat.c:585, 931).  The table file
 * (reference layout, Src/hot_x_section.c:116-131) is read if it exists; otherwise the table is built on the device (K7,
 * 500 000 samples per point like the reference) and rank 0 writes the file.  Path: mcrat_b200_dropin_set_table_path(), else
 * $MCRAT_B200_HOT_X_SECTION_FILE, else "thermal_hot_x_section.dat" in the working directory (the reference composes it from
 * its compile-time FILEPATH / MC_PATH). */
void __wrap_initalizeHotCrossSection(int rank, void *rand, FILE *fPtr);
void __wrap_cleanupInterpolationData(void);
void mcrat_b200_dropin_set_table_path(const char *path);

#ifdef __cplusplus
}
#endif
#endif
