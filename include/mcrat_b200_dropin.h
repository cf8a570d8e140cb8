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

#ifdef __cplusplus
}
#endif
#endif
