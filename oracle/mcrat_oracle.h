/*
 * mcrat_oracle.h -- TEST INFRASTRUCTURE (oracle).  Never linked into the product.
 *
 * CPU restatement, in plain C with run-time configuration, of MCRaT's
 * photon-propagation / scattering hot path.  Each
 * function cites the reference file:line it follows.  It exists to check the
 * CUDA path: tests/ compare the GPU results with this code on the same inputs
 * and the same uniform stream, and tests/test_oracle_vs_ref.py pins this code
 * against the reference's own sources compiled into oracle/_ref.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * call into this library.
 */
#ifndef MCRAT_ORACLE_H
#define MCRAT_ORACLE_H

#include "mc_mathlib.h"
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference codes, Src/mcrat.h:36-65 */
enum { MC_CARTESIAN = 0, MC_SPHERICAL = 1, MC_CYLINDRICAL = 2, MC_POLAR = 3 };
enum { MC_TWO = 0, MC_TWO_POINT_FIVE = 1, MC_THREE = 2 };
enum { MC_INTERNAL_E = 0, MC_TOTAL_E = 1, MC_SIMULATION = 2 };
enum { MC_DIRECT = 1, MC_TABLE = 2 };
/* photon types, Src/mcrat.h:52-57 */
#define MC_INJECTED_PHOTON 'i'
#define MC_COMPTONIZED_PHOTON 'k'
#define MC_CS_POOL_PHOTON 'p'
#define MC_UNABSORBED_CS_PHOTON 'c'
#define MC_REBINNED_PHOTON 'r'
#define MC_NULL_PHOTON 'N'

/* hot cross-section table extents, Src/hot_x_section.h:2-10 */
#define MC_LOG_PH_E_MIN (-12.0)
#define MC_LOG_PH_E_MAX 6.0
#define MC_N_PH_E 220
#define MC_LOG_T_MIN (-4.0)
#define MC_LOG_T_MAX 4.0
#define MC_N_T 80

/* same member order and types as `struct photon`, Src/mcrat.h:142-171
 * (NONTHERMAL_E_DIST == OFF): 176 bytes, an on-disk format of the reference */
typedef struct mc_photon {
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
} mc_photon;

/* Src/mcrat.h:173-180 */
typedef struct mc_photon_list {
    mc_photon *photons;
    int *sorted_indexes;
    int num_photons;
    int num_null_photons;
    int list_capacity;
} mc_photon_list;

/* Src/mcrat.h:194-244 (fields the hot path reads) */
typedef struct mc_hydro {
    int num_elements;
    double *r0, *r1, *r2, *r0_size, *r1_size, *r2_size, *r, *theta;
    double *v0, *v1, *v2, *dens, *dens_lab, *pres, *temp, *gamma, *B0, *B1, *B2;
    double r0_domain[2], r1_domain[2], r2_domain[2];
    double fps;
    int scatt_frame_number, inj_frame_number;
} mc_hydro;

/* compile-time switches of the reference as run-time fields */
typedef struct mc_config {
    int dimensions;      /* MC_TWO / MC_TWO_POINT_FIVE / MC_THREE */
    int geometry;        /* MC_CARTESIAN ... */
    int stokes_switch;   /* STOKES_SWITCH */
    int tau_calculation; /* MC_DIRECT / MC_TABLE */
    int cyclosynch_switch;
    int b_field_calc;
    double epsilon_b;
    double cs_rebin_e_perc; /* CYCLOSYNCHROTRON_REBIN_E_PERC, default 0.1 */
} mc_config;

typedef struct mc_oracle {
    mc_config cfg;
    /* thermal hot cross-section table + bilinear grids (hot_x_section.c:461-500) */
    int table_ready;
    double xa[MC_N_PH_E + 1];
    double ya[MC_N_T + 1];
    double za[(MC_N_PH_E + 1) * (MC_N_T + 1)]; /* za[j*(N_PH_E+1)+i] = table[i][j] */
    /* while-loop iteration counter: stream key for keyed generators */
    unsigned long long iter;
    /* photonEmitCyclosynch (all cells) calls so far: part of the key of the keyed emission streams; and the weight
     * the last call settled on (ph_weight_adjusted) */
    unsigned int emit_epoch;
    double last_emit_weight;
    FILE *log;
    long long checkinblock_evals; /* instrumentation: containment tests executed */
} mc_oracle;

extern const double MC_C_LIGHT, MC_A_RAD, MC_PL_CONST, MC_K_B, MC_M_P, MC_THOM_X_SECT, MC_M_EL, MC_FINE_STRUCT,
    MC_CHARGE_EL, MC_R_EL;

mc_oracle *mc_oracle_new(const mc_config *cfg);
void mc_oracle_free(mc_oracle *o);
void mc_oracle_set_log(mc_oracle *o, const char *path);
void mc_oracle_set_iter(mc_oracle *o, unsigned long long iter);
unsigned long long mc_oracle_get_iter(const mc_oracle *o);
/* table[i*(N_T+1)+j], i over photon energy, j over temperature */
void mc_oracle_set_thermal_table(mc_oracle *o, const double *table);
int mc_sizeof_photon(void);

/* geometry.c */
void mc_coord_to_hydro(const mc_oracle *o, double *out3, double x, double y, double z);
void mc_hydro_coord_to_mcrat(const mc_oracle *o, double *out3, double r0, double r1, double r2);
void mc_hydro_coord_to_spherical(const mc_oracle *o, double *r, double *theta, double r0, double r1, double r2);
void mc_hydro_vector_to_cartesian(const mc_oracle *o, double *out3, double v0, double v1, double v2, double x0,
                                  double x1, double x2);
double mc_hydro_element_volume(const mc_oracle *o, const mc_hydro *h, int index);
int mc_check_in_block(const mc_oracle *o, double r0, double r1, double r2, const mc_hydro *h, int idx);
int mc_find_containing_block(mc_oracle *o, double r0, double r1, double r2, const mc_hydro *h);

/* mclib.c */
void mc_lorentz_boost(const double *boost, const double *p_ph, double *result, char object);
void mc_zero_norm(double *p_ph);
int mc_find_containing_hydro_cell(mc_oracle *o, mc_photon_list *l, const mc_hydro *h, int find_nearest_block_switch,
                                  mc_rng *rng);
void mc_calc_mean_free_path(mc_oracle *o, mc_photon_list *l, const mc_hydro *h, mc_rng *rng);
void mc_update_photon_position(mc_photon_list *l, double t);
double mc_photon_event(mc_oracle *o, mc_photon_list *l, double dt_max, const mc_hydro *h, int *scattered_ph_index,
                       int *frame_scatt_cnt, int *frame_abs_cnt, mc_rng *rng);
double mc_average_photon_energy(const mc_oracle *o, const mc_photon_list *l);
void mc_ph_scatt_stats(const mc_oracle *o, const mc_photon_list *l, int *max, int *min, double *avg, double *r_avg);
void mc_ph_min_max(const mc_photon_list *l, double *min, double *max, double *min_theta, double *max_theta);

/* optical_depth.c / hot_x_section.c */
void mc_calculate_optical_depth(mc_oracle *o, mc_photon *ph, const mc_hydro *h, mc_rng *rng);
double mc_thermal_cross_section(mc_oracle *o, double photon_comv_e, double fluid_temp, mc_rng *rng);
double mc_interpolate_thermal_hot_cross_section(mc_oracle *o, double log_e, double log_theta, mc_rng *rng);
double mc_total_thermal_cross_section(double ph_comv, double theta, mc_rng *rng);
double mc_single_maxwell_juttner(double gamma, double theta);
double mc_boosted_cross_section(double norm_ph_comv, double mu, double gamma);
double mc_calc_dimless_theta(double temp);

/* mcrat_scattering.c */
void mc_muller_matrix_rotation(double theta, double *s);
void mc_find_xy(const double *v_ph, const double *vector, double *x, double *y);
double mc_find_phi(const double *x_old, const double *y_old, const double *x_new, const double *y_new);
void mc_stokes_rotation(const double *v, const double *v_ph, const double *v_ph_boosted, double *s);
int mc_single_scatter(const mc_oracle *o, double *el_comov, double *ph_comov, double *s, mc_rng *rng);
int mc_klein_nishina_scatter(const mc_oracle *o, double *theta, double *phi, double p0, double q, double u,
                             mc_rng *rng);
double mc_klein_nishina_cross_section(double energy_ratio);

/* electron.c */
void mc_single_thermal_electron(double *el_p, double temp, const double *ph_p, mc_rng *rng);
void mc_rotate_electron(double *el_p, const double *ph_p);
double mc_sample_electron_theta(double beta, mc_rng *rng);
double mc_sample_thermal_electron(double temp, mc_rng *rng);

/* photons.c */
void mc_list_init(mc_photon_list *l);
void mc_list_free(mc_photon_list *l);
void mc_list_set(mc_photon_list *l, const mc_photon *arr, int n);
void mc_list_add(mc_photon_list *l, const mc_photon *ph, size_t n);
void mc_list_set_null(mc_photon_list *l, int index);

/* mc_cyclosynch.c */
double mc_calc_cyclotron_freq(double b);
double mc_calc_b(const mc_oracle *o, double el_dens, double temp);
double mc_magnetic_field_magnitude(const mc_oracle *o, const mc_hydro *h, int idx);
double mc_cyclosynch_r_limits(int frame_scatt, int frame_inj, double fps, double r_inj, const char *min_or_max);
int mc_photon_emit_cyclosynch(mc_oracle *o, mc_photon_list *l, double r_inj, double ph_weight, int maximum_photons,
                              double theta_min, double theta_max, const mc_hydro *h, mc_rng *rng,
                              int inject_single_switch, int scatt_ph_index);
/* rebinCyclosynchCompPhotons, Src/mc_cyclosynch.c:600-710 (+ static helpers :244-598); returns the number of empty
 * bins or -1 */
int mc_rebin_cyclosynch_comp_photons(mc_oracle *o, mc_photon_list *l, int *num_cyclosynch_ph_emit,
                                     int *scatt_cyclosynch_num_ph, int max_photons);
double mc_ph_abs_cyclosynch(mc_oracle *o, mc_photon_list *l, int *num_abs_ph, int *scatt_cyclosynch_num_ph,
                            const mc_hydro *h);

/* one scatter frame: Src/mcrat.c:754-851 */
typedef struct mc_frame_stats {
    long long iterations;
    long long scatterings;
    long long relocations;
    long long photon_slots;
    double time_now;
    double last_time_step;
    int cs_emitted;
    int scatt_cyclosynch_num_ph;
} mc_frame_stats;

void mc_run_frame(mc_oracle *o, mc_photon_list *l, const mc_hydro *h, mc_rng *rng, double time_now,
                  double remaining_time, long long max_iters, int find_nearest_grid_switch, double cs_r_inj,
                  double cs_ph_weight, int cs_max_photons, double cs_theta_min, double cs_theta_max,
                  mc_frame_stats *st);

#ifdef __cplusplus
}
#endif
#endif
