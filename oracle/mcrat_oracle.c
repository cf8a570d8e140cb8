/*
 * mcrat_oracle.c -- TEST INFRASTRUCTURE (oracle).  See mcrat_oracle.h.
 *
 * Arithmetic is written operation-for-operation in the order the reference
 * evaluates it (C left-to-right association, no FMA contraction: build with
 * -ffp-contract=off) so that, linked against the same libm, it agrees with
 * oracle/_ref bit-for-bit wherever evaluation order is defined by the C
 * standard.
 */
#define _GNU_SOURCE
#include "mcrat_oracle.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdbool.h>
#include <stdlib.h>
#include <string.h>

/* Src/mclib.c:4-5 -- digits copied exactly: they are physical constants of the model */
const double MC_A_RAD = 7.56e-15, MC_C_LIGHT = 2.99792458e10, MC_PL_CONST = 6.6260755e-27,
             MC_FINE_STRUCT = 7.29735308e-3, MC_CHARGE_EL = 4.8032068e-10;
const double MC_K_B = 1.380658e-16, MC_M_P = 1.6726231e-24, MC_THOM_X_SECT = 6.65246e-25, MC_M_EL = 9.1093879e-28,
             MC_R_EL = 2.817941499892705e-13;

#define C_LIGHT MC_C_LIGHT
#define A_RAD MC_A_RAD
#define PL_CONST MC_PL_CONST
#define K_B MC_K_B
#define M_P MC_M_P
#define THOM_X_SECT MC_THOM_X_SECT
#define M_EL MC_M_EL
#define CHARGE_EL MC_CHARGE_EL

int mc_sizeof_photon(void) { return (int)sizeof(mc_photon); }

mc_oracle *mc_oracle_new(const mc_config *cfg)
{
    mc_oracle *o = (mc_oracle *)calloc(1, sizeof(*o));
    o->cfg = *cfg;
    if (o->cfg.cs_rebin_e_perc == 0) o->cfg.cs_rebin_e_perc = 0.1; /* Src/mcrat.h:310-312 */
    o->log = NULL;
    return o;
}

void mc_oracle_free(mc_oracle *o)
{
    if (!o) return;
    if (o->log) fclose(o->log);
    free(o);
}

/* position of the keyed (Philox) streams: the while-loop iteration number the next iteration draws with */
void mc_oracle_set_iter(mc_oracle *o, unsigned long long iter) { o->iter = iter; }
unsigned long long mc_oracle_get_iter(const mc_oracle *o) { return o->iter; }

void mc_oracle_set_log(mc_oracle *o, const char *path)
{
    if (o->log) fclose(o->log);
    o->log = path ? fopen(path, "w") : NULL;
}

/* Src/hot_x_section.c:461-500 (grids and the [j*(N_PH_E+1)+i] transposition) */
void mc_oracle_set_thermal_table(mc_oracle *o, const double *table)
{
    int i, j;
    double dt = (MC_LOG_T_MAX - MC_LOG_T_MIN) / MC_N_T, dph_e = (MC_LOG_PH_E_MAX - MC_LOG_PH_E_MIN) / MC_N_PH_E;
    for (i = 0; i <= MC_N_PH_E; i++) o->xa[i] = MC_LOG_PH_E_MIN + i * dph_e;
    for (i = 0; i <= MC_N_T; i++) o->ya[i] = MC_LOG_T_MIN + i * dt;
    for (i = 0; i <= MC_N_PH_E; i++)
        for (j = 0; j <= MC_N_T; j++) o->za[j * (MC_N_PH_E + 1) + i] = table[i * (MC_N_T + 1) + j];
    o->table_ready = 1;
}

/* ============================================================================ */
/* geometry.c                                                                     */
/* ============================================================================ */

/* Src/geometry.c:15-64 mcratCoordinateToHydroCoordinate */
void mc_coord_to_hydro(const mc_oracle *o, double *out, double mcrat_r0, double mcrat_r1, double mcrat_r2)
{
    double r0 = -1, r1 = -1, r2 = -1;
    const int g = o->cfg.geometry;
    if (o->cfg.dimensions == MC_TWO || o->cfg.dimensions == MC_TWO_POINT_FIVE) {
        if (g == MC_CARTESIAN || g == MC_CYLINDRICAL) {
            r0 = sqrt(mcrat_r0 * mcrat_r0 + mcrat_r1 * mcrat_r1);
            r1 = mcrat_r2;
        }
        if (g == MC_SPHERICAL) {
            r0 = sqrt(mcrat_r0 * mcrat_r0 + mcrat_r1 * mcrat_r1 + mcrat_r2 * mcrat_r2);
            r1 = acos(mcrat_r2 / r0);
        }
    } else {
        if (g == MC_CARTESIAN) {
            r0 = mcrat_r0;
            r1 = mcrat_r1;
            r2 = mcrat_r2;
        }
        if (g == MC_SPHERICAL) {
            r0 = sqrt(mcrat_r0 * mcrat_r0 + mcrat_r1 * mcrat_r1 + mcrat_r2 * mcrat_r2);
            r1 = acos(mcrat_r2 / r0);
            r2 = fmod(atan2(mcrat_r1, mcrat_r0) * 180.0 / M_PI + 360.0, 360.0) * M_PI / 180;
        }
        if (g == MC_POLAR) {
            r0 = sqrt(mcrat_r0 * mcrat_r0 + mcrat_r1 * mcrat_r1);
            r1 = fmod(atan2(mcrat_r1, mcrat_r0) * 180.0 / M_PI + 360.0, 360.0) * M_PI / 180;
            r2 = mcrat_r2;
        }
    }
    out[0] = r0;
    out[1] = r1;
    out[2] = r2;
}

/* Src/geometry.c:66-106 hydroCoordinateToSpherical */
void mc_hydro_coord_to_spherical(const mc_oracle *o, double *r, double *theta, double r0, double r1, double r2)
{
    double sph_r = 0, sph_theta = 0;
    const int g = o->cfg.geometry;
    if (o->cfg.dimensions == MC_TWO || o->cfg.dimensions == MC_TWO_POINT_FIVE) {
        if (g == MC_CARTESIAN || g == MC_CYLINDRICAL) {
            sph_r = sqrt(r0 * r0 + r1 * r1);
            sph_theta = atan2(r0, r1);
        }
        if (g == MC_SPHERICAL) {
            sph_r = r0;
            sph_theta = r1;
        }
    } else {
        if (g == MC_CARTESIAN) {
            sph_r = sqrt(r0 * r0 + r1 * r1 + r2 * r2);
            sph_theta = acos(r2 / sph_r);
        }
        if (g == MC_SPHERICAL) {
            sph_r = r0;
            sph_theta = r1;
        }
        if (g == MC_POLAR) {
            sph_r = sqrt(r0 * r0 + r2 * r2);
            sph_theta = acos(r2 / sph_r);
        }
    }
    *r = sph_r;
    *theta = sph_theta;
}

/* Src/geometry.c:108-154 hydroCoordinateToMcratCoordinate */
void mc_hydro_coord_to_mcrat(const mc_oracle *o, double *out, double hydro_r0, double hydro_r1, double hydro_r2)
{
    double x = 0, y = 0, z = 0;
    const int g = o->cfg.geometry;
    if (o->cfg.dimensions == MC_TWO || o->cfg.dimensions == MC_TWO_POINT_FIVE) {
        if (g == MC_CARTESIAN || g == MC_CYLINDRICAL) {
            x = hydro_r0 * cos(hydro_r2);
            y = hydro_r0 * sin(hydro_r2);
            z = hydro_r1;
        }
        if (g == MC_SPHERICAL) {
            x = hydro_r0 * sin(hydro_r1) * cos(hydro_r2);
            y = hydro_r0 * sin(hydro_r1) * sin(hydro_r2);
            z = hydro_r0 * cos(hydro_r1);
        }
    } else {
        if (g == MC_CARTESIAN) {
            x = hydro_r0;
            y = hydro_r1;
            z = hydro_r2;
        }
        if (g == MC_SPHERICAL) {
            x = hydro_r0 * sin(hydro_r1) * cos(hydro_r2);
            y = hydro_r0 * sin(hydro_r1) * sin(hydro_r2);
            z = hydro_r0 * cos(hydro_r1);
        }
        if (g == MC_POLAR) {
            x = hydro_r0 * cos(hydro_r1);
            y = hydro_r0 * sin(hydro_r1);
            z = hydro_r2;
        }
    }
    out[0] = x;
    out[1] = y;
    out[2] = z;
}

/* Src/geometry.c:176-187 vectorMagnitude */
static double vector_magnitude(const mc_oracle *o, double v0, double v1, double v2)
{
    if (o->cfg.dimensions == MC_TWO) return sqrt(v0 * v0 + v1 * v1);
    return sqrt(v0 * v0 + v1 * v1 + v2 * v2);
}

/* Src/geometry.c:189-253 hydroVectorToCartesian */
void mc_hydro_vector_to_cartesian(const mc_oracle *o, double *out, double v0, double v1, double v2, double x0,
                                  double x1, double x2)
{
    double t0 = 0, t1 = 0, t2 = 0;
    const int g = o->cfg.geometry;
    (void)x0;
    if (o->cfg.dimensions == MC_TWO) {
        if (g == MC_CARTESIAN || g == MC_CYLINDRICAL) {
            t0 = v0 * cos(x2);
            t1 = v0 * sin(x2);
            t2 = v1;
        }
        if (g == MC_SPHERICAL) {
            v2 = 0;
            t0 = v0 * sin(x1) * cos(x2) + v1 * cos(x1) * cos(x2) - v2 * sin(x2);
            t1 = v0 * sin(x1) * sin(x2) + v1 * cos(x1) * sin(x2) + v2 * cos(x2);
            t2 = v0 * cos(x1) - v1 * sin(x1);
        }
    } else if (o->cfg.dimensions == MC_TWO_POINT_FIVE) {
        if (g == MC_CARTESIAN || g == MC_CYLINDRICAL) {
            t0 = v0 * cos(x2) - v2 * sin(x2);
            t1 = v0 * sin(x2) + v2 * cos(x2);
            t2 = v1;
        }
        if (g == MC_SPHERICAL) {
            t0 = v0 * sin(x1) * cos(x2) + v1 * cos(x1) * cos(x2) - v2 * sin(x2);
            t1 = v0 * sin(x1) * sin(x2) + v1 * cos(x1) * sin(x2) + v2 * cos(x2);
            t2 = v0 * cos(x1) - v1 * sin(x1);
        }
    } else {
        if (g == MC_CARTESIAN) {
            t0 = v0;
            t1 = v1;
            t2 = v2;
        }
        if (g == MC_SPHERICAL) {
            t0 = v0 * sin(x1) * cos(x2) + v1 * cos(x1) * cos(x2) - v2 * sin(x2);
            t1 = v0 * sin(x1) * sin(x2) + v1 * cos(x1) * sin(x2) + v2 * cos(x2);
            t2 = v0 * cos(x1) - v1 * sin(x1);
        }
        if (g == MC_POLAR) {
            t0 = v0 * cos(x1) - v1 * sin(x1);
            t1 = v0 * sin(x1) + v1 * cos(x1);
            t2 = v2;
        }
    }
    out[0] = t0;
    out[1] = t1;
    out[2] = t2;
}

/* Src/geometry.c:255-296 hydroElementVolume */
double mc_hydro_element_volume(const mc_oracle *o, const mc_hydro *h, int index)
{
    double V = 0, r0_min, r0_max, r1_min, r1_max, r2_min, r2_max;
    const int g = o->cfg.geometry;
    r0_max = h->r0[index] + 0.5 * h->r0_size[index];
    r0_min = h->r0[index] - 0.5 * h->r0_size[index];
    r1_max = h->r1[index] + 0.5 * h->r1_size[index];
    r1_min = h->r1[index] - 0.5 * h->r1_size[index];
    if (o->cfg.dimensions == MC_TWO || o->cfg.dimensions == MC_TWO_POINT_FIVE) {
        if (g == MC_CARTESIAN || g == MC_CYLINDRICAL) V = M_PI * (r0_max * r0_max - r0_min * r0_min) * h->r1_size[index];
        if (g == MC_SPHERICAL)
            V = (2.0 * M_PI / 3.0) * (r0_max * r0_max * r0_max - r0_min * r0_min * r0_min) * (cos(r1_min) - cos(r1_max));
    } else {
        r2_max = h->r2[index] + 0.5 * h->r2_size[index];
        r2_min = h->r2[index] - 0.5 * h->r2_size[index];
        if (g == MC_CARTESIAN) V = h->r0_size[index] * h->r1_size[index] * h->r2_size[index];
        if (g == MC_SPHERICAL)
            V = (1.0 / 3.0) * (r0_max * r0_max * r0_max - r0_min * r0_min * r0_min) * (cos(r1_min) - cos(r1_max)) *
                (r2_max - r2_min);
        if (g == MC_POLAR) V = 0.5 * (r0_max * r0_max - r0_min * r0_min) * h->r1_size[index] * h->r2_size[index];
    }
    return V;
}

/* Src/geometry.c:394-417 checkInBlock */
int mc_check_in_block(const mc_oracle *o, double r0, double r1, double r2, const mc_hydro *h, int i)
{
    bool in;
    if (o->cfg.dimensions == MC_TWO || o->cfg.dimensions == MC_TWO_POINT_FIVE)
        in = (2 * fabs(r0 - h->r0[i]) - h->r0_size[i] <= 0) && (2 * fabs(r1 - h->r1[i]) - h->r1_size[i] <= 0);
    else
        in = (2 * fabs(r0 - h->r0[i]) - h->r0_size[i] <= 0) && (2 * fabs(r1 - h->r1[i]) - h->r1_size[i] <= 0) &&
             (2 * fabs(r2 - h->r2[i]) - h->r2_size[i] <= 0);
    return in ? 1 : 0;
}

/* Src/geometry.c:350-391 findContainingBlock (== findContainingBlock_grid with
 * hydro_data->grid == NULL, Src/geometry.c:426-430, Src/mcrat_io.c:1985) */
int mc_find_containing_block(mc_oracle *o, double r0, double r1, double r2, const mc_hydro *h)
{
    int i, within = 0, in = 0;
    for (i = 0; i < h->num_elements; i++) {
        in = mc_check_in_block(o, r0, r1, r2, h, i);
        o->checkinblock_evals++;
        if (in) {
            within = i;
            break; /* the reference sets i=num_elements: first match wins */
        }
    }
    if (!in) {
        if (o->log) {
            if (o->cfg.dimensions == MC_THREE)
                fprintf(o->log, "MCRaT Couldn't find a block for the photon located at r0=%e r1=%e r2=%e\n", r0, r1, r2);
            else
                fprintf(o->log, "MCRaT Couldn't find a block for the photon located at r0=%e r1=%e\n", r0, r1);
        }
        within = -1;
    }
    return within;
}

/* ============================================================================ */
/* mclib.c: lorentzBoost / zeroNorm                                               */
/* ============================================================================ */

/* Src/mclib.c:409-434 zeroNorm */
void mc_zero_norm(double *p_ph)
{
    double normalizing_factor;
    if (p_ph[0] != mc_dnrm2(3, p_ph + 1)) {
        normalizing_factor = mc_dnrm2(3, p_ph + 1);
        p_ph[1] = (p_ph[1] / normalizing_factor) * p_ph[0];
        p_ph[2] = (p_ph[2] / normalizing_factor) * p_ph[0];
        p_ph[3] = (p_ph[3] / normalizing_factor) * p_ph[0];
    }
}

/* Src/mclib.c:302-407 lorentzBoost.  NOTE: in the no-boost branch the reference
 * runs zeroNorm on the *input* array in place (Src/mclib.c:390), so p_ph is
 * deliberately not const-correct here. */
void mc_lorentz_boost(const double *b, const double *p_in, double *result, char object)
{
    double beta, gamma, L[16], pp[4];
    double *p_ph = (double *)p_in;
    if (mc_dnrm2(3, b) > 0) {
        beta = mc_dnrm2(3, b);
        gamma = 1.0 / sqrt(1 - beta * beta);
        memset(L, 0, sizeof(L));
        L[0 * 4 + 0] = gamma;
        L[0 * 4 + 1] = -1 * b[0] * gamma;
        L[0 * 4 + 2] = -1 * b[1] * gamma;
        L[0 * 4 + 3] = -1 * b[2] * gamma;
        L[1 * 4 + 1] = 1 + ((gamma - 1) * (b[0] * b[0]) / (beta * beta));
        L[1 * 4 + 2] = ((gamma - 1) * (b[0] * b[1] / (beta * beta)));
        L[1 * 4 + 3] = ((gamma - 1) * (b[0] * b[2] / (beta * beta)));
        L[2 * 4 + 2] = 1 + ((gamma - 1) * (b[1] * b[1]) / (beta * beta));
        L[2 * 4 + 3] = ((gamma - 1) * (b[1] * b[2]) / (beta * beta));
        L[3 * 4 + 3] = 1 + ((gamma - 1) * (b[2] * b[2]) / (beta * beta));
        L[1 * 4 + 0] = L[0 * 4 + 1];
        L[2 * 4 + 0] = L[0 * 4 + 2];
        L[3 * 4 + 0] = L[0 * 4 + 3];
        L[2 * 4 + 1] = L[1 * 4 + 2];
        L[3 * 4 + 1] = L[1 * 4 + 3];
        L[3 * 4 + 2] = L[2 * 4 + 3];
        mc_dgemv(4, L, p_ph, pp);
        if (object == 'p') mc_zero_norm(pp);
        result[0] = pp[0];
        result[1] = pp[1];
        result[2] = pp[2];
        result[3] = pp[3];
    } else {
        if (object == 'p') mc_zero_norm(p_ph);
        result[0] = p_ph[0];
        result[1] = p_ph[1];
        result[2] = p_ph[2];
        result[3] = p_ph[3];
    }
}

/* ============================================================================ */
/* optical_depth.c / hot_x_section.c                                              */
/* ============================================================================ */

/* Src/mc_cyclosynch.c:48-52 calcDimlessTheta */
double mc_calc_dimless_theta(double temp) { return K_B * temp / (M_EL * C_LIGHT * C_LIGHT); }

/* Src/mcrat_scattering.c:597-623 kleinNishinaCrossSection */
double mc_klein_nishina_cross_section(double energy_ratio)
{
    double result;
    if (energy_ratio >= 1e-3) {
        result = (3. / 4.) * (2. / (energy_ratio * energy_ratio) +
                              (1. / (2. * energy_ratio) - (1. + energy_ratio) / (energy_ratio * energy_ratio * energy_ratio)) *
                                  log(1. + 2. * energy_ratio) +
                              (1. + energy_ratio) / ((1. + 2. * energy_ratio) * (1. + 2. * energy_ratio)));
    } else {
        result = (1. - 2. * energy_ratio);
    }
    return result;
}

/* Src/electron.c:538-560 singleMaxwellJuttner */
double mc_single_maxwell_juttner(double gamma, double theta)
{
    double normalization;
    if (theta > 1.e-2)
        normalization = mc_bessel_Kn(2, 1. / theta) * exp(1. / theta);
    else
        normalization = sqrt(M_PI * theta / 2.);
    return ((gamma * sqrt(gamma * gamma - 1.) / (theta * normalization)) * exp(-(gamma - 1.) / theta));
}

/* Src/hot_x_section.c:369-400 boostedCrossSection (diagnostic prints omitted) */
double mc_boosted_cross_section(double norm_ph_comv, double mu, double gamma)
{
    double beta = sqrt(gamma * gamma - 1.) / gamma;
    double norm_ph_e = norm_ph_comv * gamma * (1. - mu * beta);
    return mc_klein_nishina_cross_section(norm_ph_e) * (1. - mu * beta);
}

struct mj_params {
    double norm_ph_comv, theta;
};
/* Src/hot_x_section.c:358-367 thermalCrossSectionIntegrand */
static double thermal_integrand(double *x, size_t dim, void *p)
{
    struct mj_params *fp = (struct mj_params *)p;
    (void)dim;
    return mc_single_maxwell_juttner(x[0], fp->theta) * mc_boosted_cross_section(fp->norm_ph_comv, x[1], x[0]);
}

/* Src/hot_x_section.c:324-357 calculateTotalThermalCrossSection */
double mc_total_thermal_cross_section(double ph_comv, double theta, mc_rng *rng)
{
    double result = 0, error = 0;
    double xl[2] = {1, -1};
    double xu[2] = {1. + 12 * theta, 1};
    struct mj_params params = {ph_comv, theta};
    if (theta < pow(10, MC_LOG_T_MIN) && ph_comv < pow(10, MC_LOG_PH_E_MIN)) return 1;
    if (theta < pow(10, MC_LOG_T_MIN)) return mc_klein_nishina_cross_section(ph_comv);
    if (rng->kind == MC_RNG_PHILOX) {
        /* keyed generators draw the integral's samples from a stream of its own, (sample, iteration, slot, 3): what the
         * device does (total_thermal_cross_section_mc in mcrat_b200/csrc/device_math.cuh) */
        const uint32_t stream = rng->hint_stream;
        const uint64_t draw = rng->hint_draw;
        rng->hint_stream = 3;
        rng->hint_draw = 0;
        mc_monte_plain(thermal_integrand, &params, xl, xu, 2, 500000, rng, &result, &error);
        rng->hint_stream = stream;
        rng->hint_draw = draw;
    } else {
        mc_monte_plain(thermal_integrand, &params, xl, xu, 2, 500000, rng, &result, &error);
    }
    return 0.5 * result;
}

/* Src/hot_x_section.c:545-605 interpolateThermalHotCrossSection */
double mc_interpolate_thermal_hot_cross_section(mc_oracle *o, double log_e, double log_theta, mc_rng *rng)
{
    double result = NAN;
    int status = mc_bilinear_eval(o->xa, o->ya, o->za, MC_N_PH_E + 1, MC_N_T + 1, log_e, log_theta, &result);
    if (status != 0) {
        double ph_comv = pow(10.0, log_e);
        double theta = pow(10.0, log_theta);
        result = log10(mc_total_thermal_cross_section(ph_comv, theta, rng));
    }
    return result;
}

/* Src/optical_depth.c:132-149 getThermalCrossSection (via :117-130 getCrossSection) */
double mc_thermal_cross_section(mc_oracle *o, double photon_comv_e, double fluid_temp, mc_rng *rng)
{
    if (o->cfg.tau_calculation == MC_TABLE) {
        double normalized_photon_comv_e = photon_comv_e / (M_EL * C_LIGHT);
        double theta = mc_calc_dimless_theta(fluid_temp);
        return pow(10.0, mc_interpolate_thermal_hot_cross_section(o, log10(normalized_photon_comv_e), log10(theta), rng));
    }
    return 1;
}

/* Src/optical_depth.c:7-115 calculateOpticalDepth (NONTHERMAL_E_DIST == OFF) */
void mc_calculate_optical_depth(mc_oracle *o, mc_photon *ph, const mc_hydro *h, mc_rng *rng)
{
    int idx = ph->nearest_block_index;
    double ph_phi = 0, fluid_beta[3];
    double fl_v_x, fl_v_y, fl_v_z, ph_v_norm, fl_v_norm, n_cosangle, thermal_n_dens_lab, beta, fluid_factor;
    double norm_cross_section;

    if (o->cfg.dimensions == MC_THREE) {
        mc_hydro_vector_to_cartesian(o, fluid_beta, h->v0[idx], h->v1[idx], h->v2[idx], h->r0[idx], h->r1[idx], h->r2[idx]);
    } else if (o->cfg.dimensions == MC_TWO_POINT_FIVE) {
        ph_phi = atan2(ph->r1, ph->r0);
        mc_hydro_vector_to_cartesian(o, fluid_beta, h->v0[idx], h->v1[idx], h->v2[idx], h->r0[idx], h->r1[idx], ph_phi);
    } else {
        ph_phi = atan2(ph->r1, ph->r0);
        mc_hydro_vector_to_cartesian(o, fluid_beta, h->v0[idx], h->v1[idx], 0, h->r0[idx], h->r1[idx], ph_phi);
    }
    fl_v_x = fluid_beta[0];
    fl_v_y = fluid_beta[1];
    fl_v_z = fluid_beta[2];
    fl_v_norm = sqrt(fl_v_x * fl_v_x + fl_v_y * fl_v_y + fl_v_z * fl_v_z);
    ph_v_norm = sqrt((ph->p1) * (ph->p1) + (ph->p2) * (ph->p2) + (ph->p3) * (ph->p3));
    n_cosangle = ((fl_v_x * (ph->p1)) + (fl_v_y * (ph->p2)) + (fl_v_z * (ph->p3))) / (fl_v_norm * ph_v_norm);
    beta = sqrt(1.0 - 1.0 / (h->gamma[idx] * h->gamma[idx]));
    fluid_factor = (1.0 - beta * n_cosangle);
    thermal_n_dens_lab = h->dens_lab[idx] / M_P;
    norm_cross_section = mc_thermal_cross_section(o, ph->comv_p0, h->temp[idx], rng);
    ph->total_optical_depth = (thermal_n_dens_lab) * (THOM_X_SECT * norm_cross_section) * fluid_factor;
}

/* ============================================================================ */
/* mclib.c: findContainingHydroCell / calcMeanFreePath / updatePhotonPosition      */
/* ============================================================================ */

/* Src/mclib.c:436-615 findContainingHydroCell */
int mc_find_containing_hydro_cell(mc_oracle *o, mc_photon_list *l, const mc_hydro *h, int sw, mc_rng *rng)
{
    int i, min_index, ph_block_index, count = 0;
    bool is_in_block;
    double ph_phi, ph_p_comv[4], ph_p[4], fluid_beta[3], hc[3];
    const int dims = o->cfg.dimensions;

    for (i = 0; i < l->list_capacity; i++) {
        mc_photon *ph = &l->photons[i];
        bool in_domain;
        ph_block_index = (sw == 0) ? ph->nearest_block_index : 0;
        mc_coord_to_hydro(o, hc, ph->r0, ph->r1, ph->r2);
        if (dims == MC_TWO || dims == MC_TWO_POINT_FIVE)
            in_domain = ((hc[1] < h->r1_domain[1]) && (hc[1] > h->r1_domain[0]) && (hc[0] < h->r0_domain[1]) &&
                         (hc[0] > h->r0_domain[0])) &&
                        (ph->nearest_block_index != -1);
        else
            in_domain = ((hc[2] < h->r2_domain[1]) && (hc[2] > h->r2_domain[0]) && (hc[1] < h->r1_domain[1]) &&
                         (hc[1] > h->r1_domain[0]) && (hc[0] < h->r0_domain[1]) && (hc[0] > h->r0_domain[0])) &&
                        (ph->nearest_block_index != -1);
        if (in_domain) {
            is_in_block = mc_check_in_block(o, hc[0], hc[1], hc[2], h, ph_block_index);
            if (o->cfg.cyclosynch_switch) {
                if ((ph_block_index == 0) && ((ph->comv_p0) + (ph->comv_p1) + (ph->comv_p2) + (ph->comv_p3) == 0))
                    is_in_block = 0;
            }
            if (sw == 1 || !is_in_block) {
                min_index = mc_find_containing_block(o, hc[0], hc[1], hc[2], h);
                ph->nearest_block_index = min_index;
                if (min_index != -1) {
                    ph_p[0] = ph->p0;
                    ph_p[1] = ph->p1;
                    ph_p[2] = ph->p2;
                    ph_p[3] = ph->p3;
                    if (dims == MC_THREE) {
                        mc_hydro_vector_to_cartesian(o, fluid_beta, h->v0[min_index], h->v1[min_index], h->v2[min_index],
                                                     h->r0[min_index], h->r1[min_index], h->r2[min_index]);
                    } else if (dims == MC_TWO_POINT_FIVE) {
                        ph_phi = atan2(ph->r1, ph->r0);
                        mc_hydro_vector_to_cartesian(o, fluid_beta, h->v0[min_index], h->v1[min_index], h->v2[min_index],
                                                     h->r0[min_index], h->r1[min_index], ph_phi);
                    } else {
                        ph_phi = atan2(ph->r1, ph->r0);
                        mc_hydro_vector_to_cartesian(o, fluid_beta, h->v0[min_index], h->v1[min_index], 0,
                                                     h->r0[min_index], h->r1[min_index], ph_phi);
                    }
                    mc_lorentz_boost(fluid_beta, ph_p, ph_p_comv, 'p');
                    ph->comv_p0 = ph_p_comv[0];
                    ph->comv_p1 = ph_p_comv[1];
                    ph->comv_p2 = ph_p_comv[2];
                    ph->comv_p3 = ph_p_comv[3];
                    if (rng) mc_rng_hint_mfp(rng, o->iter, (uint32_t)i); /* keyed generators: (slot, iteration) of a table fall-back */
                    mc_calculate_optical_depth(o, ph, h, rng);
                    if (ph->recalc_properties == 1) ph->recalc_properties = 0;
                    count += 1;
                } else if (o->log) {
                    fprintf(o->log, "Photon number %d Hydro grid index not found, making sure it doesnt scatter.\n", i);
                }
            }
        } else {
            ph->nearest_block_index = -1;
        }
    }
    if (sw != 0) count = 0; /* Src/mclib.c:608-611 */
    return count;
}

/* Src/mclib.c:753-763 compare2 */
static int compare2(const void *a, const void *b, void *ar)
{
    int aa = *(const int *)a, bb = *(const int *)b;
    double *arr = (double *)ar;
    return ((arr[aa] > arr[bb]) - (arr[aa] < arr[bb]));
}

/* Src/mclib.c:617-714 calcMeanFreePath */
void mc_calc_mean_free_path(mc_oracle *o, mc_photon_list *l, const mc_hydro *h, mc_rng *rng)
{
    int i;
    double mfp, default_mfp = 1e12, rnd_tracker;
    double *all_time_steps = (double *)malloc((l->list_capacity > 0 ? l->list_capacity : 1) * sizeof(double));

    for (i = 0; i < l->list_capacity; i++) {
        mc_photon *ph = &l->photons[i];
        if (ph->nearest_block_index != -1) {
            mc_rng_hint_mfp(rng, o->iter, (uint32_t)i);
            if (ph->recalc_properties == 1) {
                mc_calculate_optical_depth(o, ph, h, rng);
                ph->recalc_properties = 0;
            }
            rnd_tracker = rng->uniform_pos(rng);
            mfp = (-1.0 / ph->total_optical_depth) * log(rnd_tracker);
        } else {
            mfp = default_mfp;
        }
        ph->time_to_scatter = mfp / C_LIGHT;
    }
    for (i = 0; i < l->list_capacity; i++) {
        l->sorted_indexes[i] = i;
        all_time_steps[i] = l->photons[i].time_to_scatter;
    }
    qsort_r(l->sorted_indexes, (size_t)l->list_capacity, sizeof(int), compare2, all_time_steps);
    free(all_time_steps);
}

/* Src/mclib.c:1054-1100 updatePhotonPosition */
void mc_update_photon_position(mc_photon_list *l, double t)
{
    int i;
    double divide_p0;
    for (i = 0; i < l->list_capacity; i++) {
        mc_photon *ph = &l->photons[i];
        if ((ph->type != MC_CS_POOL_PHOTON) && (ph->weight != 0)) {
            divide_p0 = 1.0 / (ph->p0);
            (ph->r0) += (ph->p1) * divide_p0 * C_LIGHT * t;
            (ph->r1) += (ph->p2) * divide_p0 * C_LIGHT * t;
            (ph->r2) += (ph->p3) * divide_p0 * C_LIGHT * t;
        }
    }
}

/* ============================================================================ */
/* mcrat_scattering.c                                                             */
/* ============================================================================ */

/* Src/mcrat_scattering.c:10-39 mullerMatrixRotation (4x4 dgemv written out with the
 * zero entries kept so that signed zeros / NaNs propagate as in the reference) */
void mc_muller_matrix_rotation(double theta, double *s)
{
    double M[16], r[4];
    memset(M, 0, sizeof(M));
    M[0] = 1;
    M[15] = 1;
    M[1 * 4 + 1] = cos(2 * theta);
    M[2 * 4 + 2] = cos(2 * theta);
    M[1 * 4 + 2] = -1 * sin(2 * theta);
    M[2 * 4 + 1] = sin(2 * theta);
    mc_dgemv(4, M, s, r);
    s[0] = r[0];
    s[1] = r[1];
    s[2] = r[2];
    s[3] = r[3];
}

/* Src/mcrat_scattering.c:41-65 findXY */
void mc_find_xy(const double *v_ph, const double *vector, double *x, double *y)
{
    double norm;
    y[0] = (v_ph[1] * vector[2] - v_ph[2] * vector[1]);
    y[1] = -1 * (v_ph[0] * vector[2] - v_ph[2] * vector[0]);
    y[2] = (v_ph[0] * vector[1] - v_ph[1] * vector[0]);
    norm = 1.0 / sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
    y[0] *= norm;
    y[1] *= norm;
    y[2] *= norm;
    x[0] = y[1] * v_ph[2] - y[2] * v_ph[1];
    x[1] = -1 * (y[0] * v_ph[2] - y[2] * v_ph[0]);
    x[2] = y[0] * v_ph[1] - y[1] * v_ph[0];
    norm = 1.0 / sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    x[0] *= norm;
    x[1] *= norm;
    x[2] *= norm;
}

/* Src/mcrat_scattering.c:67-101 findPhi */
double mc_find_phi(const double *x_old, const double *y_old, const double *x_new, const double *y_new)
{
    double factor, dot;
    (void)x_new;
    dot = mc_ddot(3, x_old, y_new);
    if (dot > 0)
        factor = 1;
    else if (dot < 0)
        factor = -1;
    else
        factor = 0;
    dot = mc_ddot(3, y_old, y_new);
    if ((dot < -1) || (dot > 1)) dot = round(dot);
    return -1 * factor * acos(dot);
}

/* Src/mcrat_scattering.c:103-149 stokesRotation */
void mc_stokes_rotation(const double *v, const double *v_ph, const double *v_ph_boosted, double *s)
{
    double z_hat[3] = {0, 0, 1};
    double x[3] = {0, 0, 0}, y[3] = {0, 0, 0}, x_new[3] = {0, 0, 0}, y_new[3] = {0, 0, 0};
    double phi;
    mc_find_xy(v_ph, z_hat, x, y);
    mc_find_xy(v_ph, v, x_new, y_new);
    phi = mc_find_phi(x, y, x_new, y_new);
    mc_muller_matrix_rotation(phi, s);
    mc_find_xy(v_ph_boosted, v, x, y);
    mc_find_xy(v_ph_boosted, z_hat, x_new, y_new);
    phi = mc_find_phi(x, y, x_new, y_new);
    mc_muller_matrix_rotation(phi, s);
}

/* Src/mcrat_scattering.c:509-595 kleinNishinaScatter */
int mc_klein_nishina_scatter(const mc_oracle *o, double *theta, double *phi, double p0, double q, double u,
                             mc_rng *rng)
{
    double phi_dum = 0, cos_theta_dum = 0, f_phi_dum = 0, f_cos_theta_dum = 0, f_theta_dum = 0, phi_y_dum = 0,
           cos_theta_y_dum = 0, kn = 0, rand_num = 0;
    double mu = 0, phi_max = 0, norm = 0;
    int will_scatter = 0;
    double energy_ratio = p0 / (M_EL * C_LIGHT);

    kn = mc_klein_nishina_cross_section(energy_ratio);
    rand_num = rng->uniform(rng);
    if (rand_num <= kn) {
        phi_y_dum = 1;
        cos_theta_y_dum = 1;
        f_cos_theta_dum = 0;
        f_phi_dum = 0;
        while ((cos_theta_y_dum > f_cos_theta_dum)) {
            cos_theta_y_dum = rng->uniform(rng) * 2;
            cos_theta_dum = rng->uniform(rng) * 2 - 1;
            f_cos_theta_dum = pow((1 + energy_ratio * (1 - cos_theta_dum)), -2) *
                              (energy_ratio * (1 - cos_theta_dum) + (1 / (1 + energy_ratio * (1 - cos_theta_dum))) +
                               cos_theta_dum * cos_theta_dum);
        }
        *theta = acos(cos_theta_dum);
        mu = 1 + energy_ratio * (1 - cos(*theta));
        f_theta_dum = (pow(mu, -1.0) + pow(mu, -3.0) - pow(mu, -2.0) * sin(*theta) * sin(*theta)) * sin(*theta);
        while ((phi_y_dum > f_phi_dum)) {
            if (!o->cfg.stokes_switch) {
                phi_dum = rng->uniform(rng) * 2 * M_PI;
                phi_y_dum = -1;
            } else {
                if (u == 0 && q == 0) {
                    phi_dum = rng->uniform(rng) * 2 * M_PI;
                    phi_y_dum = -1;
                } else {
                    phi_max = fabs(atan2(-u, q)) / 2.0;
                    norm = (f_theta_dum + pow(mu, -2.0) * sin(*theta) * sin(*theta) * sin(*theta) *
                                              (q * cos(2 * phi_max) - u * sin(2 * phi_max)));
                    phi_y_dum = rng->uniform(rng);
                    phi_dum = rng->uniform(rng) * 2 * M_PI;
                    f_phi_dum = (f_theta_dum + pow(mu, -2.0) * sin(*theta) * sin(*theta) * sin(*theta) *
                                                   (q * cos(2 * phi_dum) - u * sin(2 * phi_dum))) /
                                norm;
                }
            }
        }
        *phi = phi_dum;
        will_scatter = 1;
    } else {
        will_scatter = 0;
    }
    return will_scatter;
}

/* y = A x for a 3x3 row-major matrix with dgemv semantics */
static void dgemv3(const double *A, const double *x, double *y) { mc_dgemv(3, A, x, y); }

/* Src/mcrat_scattering.c:151-485 singleScatter */
int mc_single_scatter(const mc_oracle *o, double *el_comov, double *ph_comov, double *s, mc_rng *rng)
{
    int scattering_occured = 0;
    double z_axis[3] = {0, 0, 1};
    double el_v[3], negative_el_v[3], ph_p_prime[4], el_p_prime[4];
    double phi0 = 0, phi1 = 0, phi = 0, theta = 0;
    double x_tilde[3] = {0, 0, 0}, y_tilde[3] = {0, 0, 0}, x_tilde_new[3] = {0, 0, 0}, y_tilde_new[3] = {0, 0, 0};
    double rot0[9], rot1[9], scatt[16], scatt_result[4], result0[3], result1[3], result[4], ph_p_orig[4];
    double *ph_p = ph_p_prime + 1; /* gsl_vector_view_array((ph_p_prime+1), 3), :233 */

    el_v[0] = el_comov[1] / el_comov[0];
    el_v[1] = el_comov[2] / el_comov[0];
    el_v[2] = el_comov[3] / el_comov[0];

    mc_lorentz_boost(el_v, el_comov, el_p_prime, 'e');
    mc_lorentz_boost(el_v, ph_comov, ph_p_prime, 'p');

    if (o->cfg.stokes_switch) mc_stokes_rotation(el_v, (ph_comov + 1), (ph_p_prime + 1), s);

    ph_p_orig[0] = ph_p_prime[0];
    ph_p_orig[1] = ph_p_prime[1];
    ph_p_orig[2] = ph_p_prime[2];
    ph_p_orig[3] = ph_p_prime[3];

    phi0 = atan2(ph_p_prime[2], ph_p_prime[1]);
    memset(rot0, 0, sizeof(rot0));
    rot0[2 * 3 + 2] = 1;
    rot0[0 * 3 + 0] = cos(-phi0);
    rot0[1 * 3 + 1] = cos(-phi0);
    rot0[0 * 3 + 1] = -sin(-phi0);
    rot0[1 * 3 + 0] = sin(-phi0);
    dgemv3(rot0, ph_p, result0);

    ph_p_prime[1] = result0[0];
    ph_p_prime[2] = 0;
    ph_p_prime[3] = result0[2];

    phi1 = atan2(result0[2], result0[0]);

    memset(rot1, 0, sizeof(rot1));
    rot1[1 * 3 + 1] = 1;
    rot1[0 * 3 + 0] = cos(-phi1);
    rot1[2 * 3 + 2] = cos(-phi1);
    rot1[0 * 3 + 2] = -sin(-phi1);
    rot1[2 * 3 + 0] = sin(-phi1);
    dgemv3(rot1, ph_p, result1);

    ph_p_prime[1] = ph_p_prime[0];
    ph_p_prime[2] = result1[1];
    ph_p_prime[3] = 0;

    scattering_occured = mc_klein_nishina_scatter(o, &theta, &phi, ph_p_prime[0], s[1], s[2], rng);

    if (scattering_occured == 1) {
        result[0] = (ph_p_prime[0]) / (1 + (((ph_p_prime[0]) * (1 - cos(theta))) / (M_EL * C_LIGHT)));
        result[1] = result[0] * cos(theta);
        result[2] = result[0] * sin(theta) * sin(phi);
        result[3] = result[0] * sin(theta) * cos(phi);

        /* the electron update (:342-348) does not feed back into any output */

        ph_p_prime[0] = result[0];
        ph_p_prime[1] = result[1];
        ph_p_prime[2] = result[2];
        ph_p_prime[3] = result[3];
        memset(rot1, 0, sizeof(rot1));
        rot1[1 * 3 + 1] = 1;
        rot1[0 * 3 + 0] = cos(-phi1);
        rot1[2 * 3 + 2] = cos(-phi1);
        rot1[0 * 3 + 2] = sin(-phi1);
        rot1[2 * 3 + 0] = -sin(-phi1);
        dgemv3(rot1, ph_p, result1);

        ph_p_prime[1] = result1[0];
        ph_p_prime[2] = result1[1];
        ph_p_prime[3] = result1[2];
        memset(rot0, 0, sizeof(rot0));
        rot0[2 * 3 + 2] = 1;
        rot0[0 * 3 + 0] = cos(-phi0);
        rot0[1 * 3 + 1] = cos(-phi0);
        rot0[0 * 3 + 1] = sin(-phi0);
        rot0[1 * 3 + 0] = -sin(-phi0);
        dgemv3(rot0, ph_p, result0);

        if (o->cfg.stokes_switch) {
            mc_find_xy(ph_p_orig + 1, z_axis, x_tilde, y_tilde);
            mc_find_xy(result0, ph_p_orig + 1, x_tilde_new, y_tilde_new);
            phi = mc_find_phi(x_tilde, y_tilde, x_tilde_new, y_tilde_new);
            mc_muller_matrix_rotation(phi, s);

            theta = acos((ph_p_orig[1] * result0[0] + ph_p_orig[2] * result0[1] + ph_p_orig[3] * result0[2]) /
                         (ph_p_orig[0] * (ph_p_prime[0])));

            memset(scatt, 0, sizeof(scatt));
            scatt[0 * 4 + 0] = 1.0 + pow(cos(theta), 2.0) + ((1 - cos(theta)) * (ph_p_orig[0] - result[0]) / (M_EL * C_LIGHT));
            scatt[0 * 4 + 1] = sin(theta) * sin(theta);
            scatt[1 * 4 + 0] = sin(theta) * sin(theta);
            scatt[1 * 4 + 1] = 1.0 + cos(theta) * cos(theta);
            scatt[2 * 4 + 2] = 2.0 * cos(theta);
            scatt[3 * 4 + 3] = 2.0 * cos(theta) + ((cos(theta)) * (1 - cos(theta)) * (ph_p_orig[0] - result[0]) / (M_EL * C_LIGHT));
            mc_dgemv(4, scatt, s, scatt_result);

            s[0] = scatt_result[0] / scatt_result[0];
            s[1] = scatt_result[1] / scatt_result[0];
            s[2] = scatt_result[2] / scatt_result[0];
            s[3] = scatt_result[3] / scatt_result[0];

            mc_find_xy(result0, ph_p_orig + 1, x_tilde, y_tilde);
            mc_find_xy(result0, z_axis, x_tilde_new, y_tilde_new);
            phi = mc_find_phi(x_tilde, y_tilde, x_tilde_new, y_tilde_new);
            mc_muller_matrix_rotation(phi, s);
        }

        ph_p_prime[1] = result0[0];
        ph_p_prime[2] = result0[1];
        ph_p_prime[3] = result0[2];

        negative_el_v[0] = (-1 * el_v[0]);
        negative_el_v[1] = (-1 * el_v[1]);
        negative_el_v[2] = (-1 * el_v[2]);

        mc_lorentz_boost(negative_el_v, ph_p_prime, ph_comov, 'p');

        if (o->cfg.stokes_switch) mc_stokes_rotation(negative_el_v, (ph_p_prime + 1), (ph_comov + 1), s);
    }
    return scattering_occured;
}

/* ============================================================================ */
/* electron.c                                                                      */
/* ============================================================================ */

/* Src/electron.c:177-200 sampleElectronTheta */
double mc_sample_electron_theta(double beta, mc_rng *rng)
{
    return acos((1 - sqrt(1 + beta * beta + 2 * beta - 4 * beta * rng->uniform(rng))) / beta);
}

/* Src/electron.c:202-237 sampleThermalElectron */
double mc_sample_thermal_electron(double temp, mc_rng *rng)
{
    double gamma = 1, factor, x_dum = 0, y_dum, f_x_dum, beta_x_dum;
    if (temp >= 1e7) {
        factor = K_B * temp / (M_EL * C_LIGHT * C_LIGHT);
        y_dum = 1;
        f_x_dum = 0;
        while ((isnan(f_x_dum) != 0) || (y_dum > f_x_dum)) {
            x_dum = rng->uniform_pos(rng) * (1 + 100 * factor);
            beta_x_dum = sqrt(1 - (1 / (x_dum * x_dum)));
            y_dum = rng->uniform(rng) / 2.0;
            f_x_dum = x_dum * x_dum * (beta_x_dum / mc_bessel_Kn(2, 1.0 / factor)) * exp(-1 * x_dum / factor);
        }
        gamma = x_dum;
    } else {
        /* the three deviates sit in one expression in the reference (:233); the
         * order in which a compiler evaluates them is unspecified.  They are drawn
         * left to right here. */
        double g1, g2, g3;
        factor = sqrt(K_B * temp / M_EL);
        g1 = mc_ran_gaussian(rng, factor);
        g2 = mc_ran_gaussian(rng, factor);
        g3 = mc_ran_gaussian(rng, factor);
        gamma = 1.0 / sqrt(1 - (pow(g1 / C_LIGHT, 2) + pow(g2 / C_LIGHT, 2) + pow(g3 / C_LIGHT, 2)));
    }
    return gamma;
}

/* Src/electron.c:126-175 rotateElectron */
void mc_rotate_electron(double *el_p, const double *ph_p)
{
    double ph_theta, ph_phi, rot[9], result[3];
    double *el_p_prime = el_p + 1;
    ph_phi = atan2(ph_p[2], ph_p[3]);
    ph_theta = atan2(sqrt(pow(ph_p[2], 2) + pow(ph_p[3], 2)), ph_p[1]);
    memset(rot, 0, sizeof(rot));
    rot[1 * 3 + 1] = 1;
    rot[2 * 3 + 2] = cos(ph_theta);
    rot[0 * 3 + 0] = cos(ph_theta);
    rot[0 * 3 + 2] = -sin(ph_theta);
    rot[2 * 3 + 0] = sin(ph_theta);
    dgemv3(rot, el_p_prime, result);
    memset(rot, 0, sizeof(rot));
    rot[0 * 3 + 0] = 1;
    rot[1 * 3 + 1] = cos(-ph_phi);
    rot[2 * 3 + 2] = cos(-ph_phi);
    rot[1 * 3 + 2] = -sin(-ph_phi);
    rot[2 * 3 + 1] = sin(-ph_phi);
    dgemv3(rot, result, el_p_prime);
}

/* Src/electron.c:70-94 singleThermalElectron */
void mc_single_thermal_electron(double *el_p, double temp, const double *ph_p, mc_rng *rng)
{
    double gamma, beta, phi, theta;
    gamma = mc_sample_thermal_electron(temp, rng);
    beta = sqrt(1 - (1 / (gamma * gamma)));
    phi = rng->uniform(rng) * 2 * M_PI;
    theta = mc_sample_electron_theta(beta, rng);
    el_p[0] = gamma * (M_EL) * (C_LIGHT);
    el_p[1] = gamma * (M_EL) * (C_LIGHT)*beta * cos(theta);
    el_p[2] = gamma * (M_EL) * (C_LIGHT)*beta * sin(theta) * sin(phi);
    el_p[3] = gamma * (M_EL) * (C_LIGHT)*beta * sin(theta) * cos(phi);
    mc_rotate_electron(el_p, ph_p);
}

/* ============================================================================ */
/* mclib.c: photonEvent + statistics                                               */
/* ============================================================================ */

/* Src/mclib.c:1107-1356 photonEvent */
double mc_photon_event(mc_oracle *o, mc_photon_list *l, double dt_max, const mc_hydro *h, int *scattered_ph_index,
                       int *frame_scatt_cnt, int *frame_abs_cnt, mc_rng *rng)
{
    int i = 0, index = 0, ph_index = 0, event_did_occur = 0;
    double scatt_time = 0, old_scatt_time = 0, ph_phi = 0, fluid_temp = 0;
    double ph_p[4], el_p_comov[4], ph_p_comov[4], fluid_beta[3], negative_fluid_beta[3], s[4];
    const int dims = o->cfg.dimensions;
    (void)frame_abs_cnt;

    mc_rng_hint_event(rng, o->iter);
    while (i < l->list_capacity && event_did_occur == 0) {
        mc_photon *ph = &l->photons[l->sorted_indexes[i]];
        ph_index = l->sorted_indexes[i];
        scatt_time = ph->time_to_scatter;
        if (scatt_time < dt_max) {
            mc_update_photon_position(l, scatt_time - old_scatt_time);
            index = ph->nearest_block_index;
            fluid_temp = h->temp[index];
            ph_phi = atan2((ph->r1), ((ph->r0)));
            if (dims == MC_THREE)
                mc_hydro_vector_to_cartesian(o, fluid_beta, h->v0[index], h->v1[index], h->v2[index], h->r0[index],
                                             h->r1[index], h->r2[index]);
            else if (dims == MC_TWO_POINT_FIVE)
                mc_hydro_vector_to_cartesian(o, fluid_beta, h->v0[index], h->v1[index], h->v2[index], h->r0[index],
                                             h->r1[index], ph_phi);
            else
                mc_hydro_vector_to_cartesian(o, fluid_beta, h->v0[index], h->v1[index], 0, h->r0[index], h->r1[index],
                                             ph_phi);
            ph_p[0] = ph->p0;
            ph_p[1] = ph->p1;
            ph_p[2] = ph->p2;
            ph_p[3] = ph->p3;
            ph_p_comov[0] = ph->comv_p0;
            ph_p_comov[1] = ph->comv_p1;
            ph_p_comov[2] = ph->comv_p2;
            ph_p_comov[3] = ph->comv_p3;
            s[0] = ph->s0;
            s[1] = ph->s1;
            s[2] = ph->s2;
            s[3] = ph->s3;

            if (o->cfg.stokes_switch) mc_stokes_rotation(fluid_beta, (ph_p + 1), (ph_p_comov + 1), s);

            mc_single_thermal_electron(el_p_comov, fluid_temp, ph_p_comov, rng);

            event_did_occur = mc_single_scatter(o, el_p_comov, ph_p_comov, s, rng);

            if (event_did_occur == 1) {
                negative_fluid_beta[0] = -1 * (fluid_beta[0]);
                negative_fluid_beta[1] = -1 * (fluid_beta[1]);
                negative_fluid_beta[2] = -1 * (fluid_beta[2]);
                mc_lorentz_boost(negative_fluid_beta, ph_p_comov, ph_p, 'p');
                if (o->cfg.stokes_switch) {
                    mc_stokes_rotation(negative_fluid_beta, (ph_p_comov + 1), (ph_p + 1), s);
                    ph->s0 = s[0];
                    ph->s1 = s[1];
                    ph->s2 = s[2];
                    ph->s3 = s[3];
                }
                if (((ph_p[0]) * C_LIGHT / 1.6e-9) > 1e4) {
                    if (o->log) fprintf(o->log, "Extremely High Photon Energy!!!!!!!!\n");
                }
                ph->p0 = ph_p[0];
                ph->p1 = ph_p[1];
                ph->p2 = ph_p[2];
                ph->p3 = ph_p[3];
                ph->comv_p0 = ph_p_comov[0];
                ph->comv_p1 = ph_p_comov[1];
                ph->comv_p2 = ph_p_comov[2];
                ph->comv_p3 = ph_p_comov[3];
                ph->num_scatt += 1;
                *frame_scatt_cnt += 1;
                ph->recalc_properties = 1;
            }
        } else {
            scatt_time = dt_max;
            mc_update_photon_position(l, scatt_time - old_scatt_time);
            event_did_occur = 1;
        }
        old_scatt_time = scatt_time;
        i++;
    }
    *scattered_ph_index = ph_index;
    return scatt_time;
}

/* Src/mclib.c:1358-1383 averagePhotonEnergy */
double mc_average_photon_energy(const mc_oracle *o, const mc_photon_list *l)
{
    double e_sum = 0, w_sum = 0;
    for (int i = 0; i < l->list_capacity; i++) {
        const mc_photon *ph = &l->photons[i];
        if (!o->cfg.cyclosynch_switch || (ph->weight != 0)) {
            e_sum += ((ph->p0) * (ph->weight));
            w_sum += (ph->weight);
        }
    }
    return (e_sum * C_LIGHT) / w_sum;
}

/* Src/mclib.c:1385-1462 phScattStats (return values only; the per-type r averages are log lines) */
void mc_ph_scatt_stats(const mc_oracle *o, const mc_photon_list *l, int *max, int *min, double *avg, double *r_avg)
{
    int temp_max = 0, temp_min = INT_MAX, count = 0;
    double sum = 0, avg_r_sum = 0;
    for (int i = 0; i < l->list_capacity; i++) {
        const mc_photon *ph = &l->photons[i];
        if (!o->cfg.cyclosynch_switch || (ph->weight != 0)) {
            sum += (ph->num_scatt);
            avg_r_sum += sqrt((ph->r0) * (ph->r0) + (ph->r1) * (ph->r1) + (ph->r2) * (ph->r2));
            if ((ph->num_scatt) > temp_max) temp_max = (int)(ph->num_scatt);
            if ((ph->num_scatt) < temp_min) temp_min = (int)(ph->num_scatt);
            count++;
        }
    }
    *avg = sum / count;
    *r_avg = avg_r_sum / count;
    *max = temp_max;
    *min = temp_min;
}

/* Src/mclib.c:1465-1515 phMinMax */
void mc_ph_min_max(const mc_photon_list *l, double *min, double *max, double *min_theta, double *max_theta)
{
    double temp_r_max = 0, temp_r_min = DBL_MAX, temp_theta_max = 0, temp_theta_min = DBL_MAX;
    for (int i = 0; i < l->list_capacity; i++) {
        const mc_photon *ph = &l->photons[i];
        if (ph->weight != 0) {
            double ph_r = sqrt((ph->r0) * (ph->r0) + (ph->r1) * (ph->r1) + (ph->r2) * (ph->r2));
            double ph_theta = acos((ph->r2) / ph_r);
            if (ph_r > temp_r_max) temp_r_max = ph_r;
            if (ph_r < temp_r_min) temp_r_min = ph_r;
            if (ph_theta > temp_theta_max) temp_theta_max = ph_theta;
            if (ph_theta < temp_theta_min) temp_theta_min = ph_theta;
        }
    }
    *max = temp_r_max;
    *min = temp_r_min;
    *max_theta = temp_theta_max;
    *min_theta = temp_theta_min;
}

/* ============================================================================ */
/* photons.c                                                                        */
/* ============================================================================ */

static void verify_photon_num(mc_photon_list *l)
{
    if (l->num_photons + l->num_null_photons != l->list_capacity) {
        printf("Error with incremenitng real or null photon in the photonList. conservation of photon error\n");
        exit(1);
    }
}

/* Src/photons.c:3-14 */
void mc_list_init(mc_photon_list *l)
{
    l->photons = NULL;
    l->sorted_indexes = NULL;
    l->num_photons = 0;
    l->num_null_photons = 0;
    l->list_capacity = 0;
}

/* Src/photons.c:16-26 */
void mc_list_free(mc_photon_list *l)
{
    free(l->photons);
    free(l->sorted_indexes);
    mc_list_init(l);
}

/* Src/photons.c:208-251 setNullPhoton */
void mc_list_set_null(mc_photon_list *l, int index)
{
    mc_photon *p = &l->photons[index];
    p->type = MC_NULL_PHOTON;
    p->weight = 0;
    p->nearest_block_index = -1;
    p->recalc_properties = 0;
    p->p0 = p->p1 = p->p2 = p->p3 = 0;
    p->comv_p0 = p->comv_p1 = p->comv_p2 = p->comv_p3 = 0;
    p->r0 = p->r1 = p->r2 = 0;
    p->s0 = p->s1 = p->s2 = p->s3 = 0;
    p->num_scatt = 0;
    p->total_optical_depth = 0;
    l->num_photons -= 1;
    l->num_null_photons += 1;
    verify_photon_num(l);
}

/* Src/photons.c:82-108 setPhotonList */
void mc_list_set(mc_photon_list *l, const mc_photon *arr, int n)
{
    int nulls = 0;
    if (l->photons != NULL) mc_list_free(l);
    l->photons = (mc_photon *)malloc((n > 0 ? n : 1) * sizeof(mc_photon));
    l->sorted_indexes = (int *)malloc((n > 0 ? n : 1) * sizeof(int));
    memcpy(l->photons, arr, (size_t)n * sizeof(mc_photon));
    l->list_capacity = n;
    l->num_photons = n;
    for (int i = 0; i < n; i++)
        if (l->photons[i].type == MC_NULL_PHOTON) nulls++;
    l->num_null_photons = nulls;
}

/* Src/photons.c:37-80 reallocatePhotonListMemory */
static void list_realloc(mc_photon_list *l, int new_capacity)
{
    int old = l->list_capacity;
    l->photons = (mc_photon *)realloc(l->photons, (size_t)new_capacity * sizeof(mc_photon));
    l->sorted_indexes = (int *)realloc(l->sorted_indexes, (size_t)new_capacity * sizeof(int));
    if (!l->photons || !l->sorted_indexes) {
        printf("Error with reserving space to hold new photons\n");
        exit(1);
    }
    l->list_capacity = new_capacity;
    l->num_photons += (new_capacity - old);
    for (int i = old; i < new_capacity; i++) mc_list_set_null(l, i);
}

/* Src/photons.c:110-206 addToPhotonList */
void mc_list_add(mc_photon_list *l, const mc_photon *ph, size_t num_photons)
{
    int idx = 0, i, j = 0, new_capacity;
    if ((l->num_photons >= l->list_capacity) && ((size_t)l->num_null_photons <= num_photons)) {
        if ((size_t)l->list_capacity * 2 > (size_t)l->list_capacity + num_photons)
            new_capacity = l->list_capacity * 2;
        else
            new_capacity = (int)(l->list_capacity * (num_photons / l->list_capacity));
        list_realloc(l, new_capacity);
    }
    if (num_photons == 1) {
        if (l->num_null_photons == 0) {
            idx = l->num_photons;
        } else {
            for (i = 0; i < l->list_capacity; i++) {
                if (l->photons[i].type == MC_NULL_PHOTON) {
                    idx = i;
                    break;
                }
            }
        }
        memcpy(&l->photons[idx], ph, sizeof(mc_photon));
        l->num_photons += 1;
        l->num_null_photons -= 1;
        verify_photon_num(l);
    } else {
        int *null_idx = (int *)malloc((l->num_null_photons > 0 ? l->num_null_photons : 1) * sizeof(int));
        if (num_photons > (size_t)l->num_null_photons) {
            printf("Adding to the photon list has failed. the number of null photons in the list is less than the "
                   "number of photons to add to the list. %d vs %zu",
                   l->num_null_photons, num_photons);
            exit(1);
        }
        for (i = 0; i < l->list_capacity; i++)
            if (l->photons[i].type == MC_NULL_PHOTON) null_idx[j++] = i;
        for (i = 0; i < (int)num_photons; i++) {
            if (ph[i].type != MC_NULL_PHOTON) {
                idx = null_idx[i];
                memcpy(&l->photons[idx], &ph[i], sizeof(mc_photon));
                l->num_photons += 1;
                l->num_null_photons -= 1;
                verify_photon_num(l);
            }
        }
        free(null_idx);
    }
}

/* ============================================================================ */
/* mc_cyclosynch.c                                                                  */
/* ============================================================================ */

/* Src/mc_cyclosynch.c:30-34 calcCyclotronFreq */
double mc_calc_cyclotron_freq(double b) { return CHARGE_EL * b / (2 * M_PI * M_EL * C_LIGHT); }

/* Src/mc_cyclosynch.c:54-76 calcB */
double mc_calc_b(const mc_oracle *o, double el_dens, double temp)
{
    if (o->cfg.b_field_calc == MC_INTERNAL_E) return sqrt(o->cfg.epsilon_b * 8 * M_PI * 3 * el_dens * K_B * temp / 2);
    if (o->cfg.b_field_calc == MC_TOTAL_E)
        return sqrt(8 * M_PI * o->cfg.epsilon_b * (el_dens * M_P * C_LIGHT * C_LIGHT + 4 * A_RAD * temp * temp * temp * temp / 3));
    return 0;
}

/* Src/mc_cyclosynch.c:78-92 getMagneticFieldMagnitude */
double mc_magnetic_field_magnitude(const mc_oracle *o, const mc_hydro *h, int i)
{
    if (o->cfg.b_field_calc == MC_TOTAL_E || o->cfg.b_field_calc == MC_INTERNAL_E) {
        double el_dens = h->dens[i] / M_P;
        return mc_calc_b(o, el_dens, h->temp[i]);
    }
    if (o->cfg.dimensions == MC_TWO) return vector_magnitude(o, h->B0[i], h->B1[i], 0);
    return vector_magnitude(o, h->B0[i], h->B1[i], h->B2[i]);
}

/* Src/mc_cyclosynch.c:225-242 calcCyclosynchRLimits */
double mc_cyclosynch_r_limits(int frame_scatt, int frame_inj, double fps, double r_inj, const char *min_or_max)
{
    double val = r_inj;
    if (strcmp(min_or_max, "min") == 0)
        val += (C_LIGHT * (frame_scatt - frame_inj) / fps - 0.5 * C_LIGHT / fps);
    else
        val += (C_LIGHT * (frame_scatt - frame_inj) / fps + 0.5 * C_LIGHT / fps);
    return val;
}

/* Src/mc_cyclosynch.c:185-195 blackbody_ph_spect */
static double blackbody_ph_spect(double nu, void *p)
{
    double temp = ((double *)p)[0];
    return (8 * M_PI * nu * nu) / (exp(PL_CONST * nu / (K_B * temp)) - 1) / (C_LIGHT * C_LIGHT * C_LIGHT);
}

static void cell_corners(const mc_oracle *o, const mc_hydro *h, int i, double *ri, double *ti, double *ro, double *to)
{
    if (o->cfg.dimensions == MC_THREE) {
        mc_hydro_coord_to_spherical(o, ri, ti, fabs(h->r0[i]) - 0.5 * h->r0_size[i], fabs(h->r1[i]) - 0.5 * h->r1_size[i],
                                    fabs(h->r2[i]) - 0.5 * h->r2_size[i]);
        mc_hydro_coord_to_spherical(o, ro, to, fabs(h->r0[i]) + 0.5 * h->r0_size[i], fabs(h->r1[i]) + 0.5 * h->r1_size[i],
                                    fabs(h->r2[i]) + 0.5 * h->r2_size[i]);
    } else {
        mc_hydro_coord_to_spherical(o, ri, ti, h->r0[i] - 0.5 * h->r0_size[i], h->r1[i] - 0.5 * h->r1_size[i], 0);
        mc_hydro_coord_to_spherical(o, ro, to, h->r0[i] + 0.5 * h->r0_size[i], h->r1[i] + 0.5 * h->r1_size[i], 0);
    }
}

static void fill_emitted_photon(const mc_oracle *o, mc_photon *e, const mc_hydro *h, int i, double nu_c,
                                double weight, int block_index, mc_rng *rng, double *position_phi_out)
{
    double position_phi, com_v_phi, com_v_theta, fr_dum = nu_c, p_comv[4], boost[3], l_boost[4], pos[3];
    const int dims = o->cfg.dimensions;
    if (dims == MC_TWO || dims == MC_TWO_POINT_FIVE)
        position_phi = rng->uniform(rng) * 2 * M_PI;
    else
        position_phi = 0;
    com_v_phi = rng->uniform(rng) * 2 * M_PI;
    com_v_theta = rng->uniform(rng) * M_PI;
    p_comv[0] = PL_CONST * fr_dum / C_LIGHT;
    p_comv[1] = (PL_CONST * fr_dum / C_LIGHT) * sin(com_v_theta) * cos(com_v_phi);
    p_comv[2] = (PL_CONST * fr_dum / C_LIGHT) * sin(com_v_theta) * sin(com_v_phi);
    p_comv[3] = (PL_CONST * fr_dum / C_LIGHT) * cos(com_v_theta);
    if (dims == MC_THREE)
        mc_hydro_vector_to_cartesian(o, boost, h->v0[i], h->v1[i], h->v2[i], h->r0[i], h->r1[i], h->r2[i]);
    else if (dims == MC_TWO_POINT_FIVE)
        mc_hydro_vector_to_cartesian(o, boost, h->v0[i], h->v1[i], h->v2[i], h->r0[i], h->r1[i], position_phi);
    else
        mc_hydro_vector_to_cartesian(o, boost, h->v0[i], h->v1[i], 0, h->r0[i], h->r1[i], position_phi);
    boost[0] *= -1;
    boost[1] *= -1;
    boost[2] *= -1;
    mc_lorentz_boost(boost, p_comv, l_boost, 'p');
    memset(e, 0, sizeof(*e));
    e->p0 = l_boost[0];
    e->p1 = l_boost[1];
    e->p2 = l_boost[2];
    e->p3 = l_boost[3];
    e->comv_p0 = p_comv[0];
    e->comv_p1 = p_comv[1];
    e->comv_p2 = p_comv[2];
    e->comv_p3 = p_comv[3];
    if (dims == MC_THREE)
        mc_hydro_coord_to_mcrat(o, pos, h->r0[i], h->r1[i], h->r2[i]);
    else
        mc_hydro_coord_to_mcrat(o, pos, h->r0[i], h->r1[i], position_phi);
    e->r0 = pos[0];
    e->r1 = pos[1];
    e->r2 = pos[2];
    e->s0 = 1;
    e->s1 = 0;
    e->s2 = 0;
    e->s3 = 0;
    e->num_scatt = 0;
    e->weight = weight;
    e->nearest_block_index = block_index;
    e->type = MC_CS_POOL_PHOTON;
    e->recalc_properties = 1;
    *position_phi_out = position_phi;
}

/* Src/mc_cyclosynch.c:1176-1569 photonEmitCyclosynch.
 * The uninitialised members of freshly malloc'ed emitted photons
 * (time_to_scatter, total_optical_depth) are zeroed here. */
int mc_photon_emit_cyclosynch(mc_oracle *o, mc_photon_list *l, double r_inj, double ph_weight, int maximum_photons,
                              double theta_min, double theta_max, const mc_hydro *h, mc_rng *rng,
                              int inject_single_switch, int scatt_ph_index)
{
    double rmin = 0, rmax = 0, max_photons = o->cfg.cs_rebin_e_perc * maximum_photons;
    double ph_weight_adjusted = 0, position_phi = 0, nu_c = 0, error = 0, ph_dens_calc = 0, b_field = 0;
    double ri = 0, ro = 0, ti = 0, to = 0, params[3];
    int block_cnt = 0, i, j = 0, k = 0, *ph_dens = NULL, ph_tot = 0, net_ph = 0, min_photons = 1, idx = 0;
    unsigned int search_pass = 0;
    mc_photon *ph_emit = NULL;
    const int dims = o->cfg.dimensions;

    if (inject_single_switch == 0) {
        rmin = mc_cyclosynch_r_limits(h->scatt_frame_number, h->inj_frame_number, h->fps, r_inj, "min");
        rmax = mc_cyclosynch_r_limits(h->scatt_frame_number, h->inj_frame_number, h->fps, r_inj, "max");
        for (i = 0; i < h->num_elements; i++) {
            cell_corners(o, h, i, &ri, &ti, &ro, &to);
            if ((rmin <= ro) && (ri < rmax) && (to >= theta_min) && (ti < theta_max)) block_cnt += 1;
        }
        if (block_cnt == 0) min_photons = block_cnt;
        ph_dens = (int *)malloc((block_cnt > 0 ? block_cnt : 1) * sizeof(int));
        j = 0;
        ph_tot = -1;
        ph_weight_adjusted = ph_weight;
        o->emit_epoch += 1;
        while ((ph_tot > max_photons) || (ph_tot < min_photons)) {
            j = 0;
            ph_tot = 0;
            search_pass += 1;
            for (i = 0; i < h->num_elements; i++) {
                cell_corners(o, h, i, &ri, &ti, &ro, &to);
                if ((rmin <= ro) && (ri < rmax) && (to >= theta_min) && (ti < theta_max)) {
                    b_field = mc_magnetic_field_magnitude(o, h, i);
                    nu_c = mc_calc_cyclotron_freq(b_field);
                    params[0] = h->temp[i];
                    params[1] = mc_calc_dimless_theta(h->temp[i]);
                    params[2] = h->dens[i] / M_P;
                    mc_integrate_adaptive(blackbody_ph_spect, params, 10, nu_c, 0, 1e-2, 10000, &ph_dens_calc, &error);
                    ph_dens_calc *= mc_hydro_element_volume(o, h, i) / (ph_weight_adjusted);
                    /* keyed generators: one stream per selected cell and weight-search pass (the device draws them in parallel) */
                    if (rng->kind == MC_RNG_PHILOX)
                        mc_rng_hint_keyed(rng, 4, (uint32_t)j, ((uint64_t)o->emit_epoch << 32) | (uint64_t)(search_pass - 1));
                    ph_dens[j] = (int)mc_ran_poisson(rng, ph_dens_calc);
                    ph_tot += ph_dens[j];
                    j++;
                }
            }
            if (ph_tot > max_photons)
                ph_weight_adjusted *= 10;
            else if (ph_tot < min_photons)
                ph_weight_adjusted *= 0.5;
        }
    } else {
        ph_tot = 1;
    }

    ph_emit = (mc_photon *)malloc((ph_tot > 0 ? ph_tot : 1) * sizeof(mc_photon));

    if (inject_single_switch == 0) {
        net_ph = ph_tot;
        ph_tot = 0;
        for (i = 0; i < h->num_elements; i++) {
            cell_corners(o, h, i, &ri, &ti, &ro, &to);
            if ((rmin <= ro) && (ri < rmax) && (to >= theta_min) && (ti < theta_max)) {
                b_field = mc_magnetic_field_magnitude(o, h, i);
                nu_c = mc_calc_cyclotron_freq(b_field);
                for (j = 0; j < ph_dens[k]; j++) {
                    if (rng->kind == MC_RNG_PHILOX) mc_rng_hint_keyed(rng, 5, (uint32_t)ph_tot, (uint64_t)o->emit_epoch << 32);
                    fill_emitted_photon(o, &ph_emit[ph_tot], h, i, nu_c, ph_weight_adjusted, 0, rng, &position_phi);
                    ph_tot++;
                    if (net_ph == ph_tot) i = h->num_elements;
                }
                k++;
            }
        }
    } else {
        mc_photon *tmp = &l->photons[scatt_ph_index];
        double pr, pr2, pr3, pos[3];
        i = tmp->nearest_block_index;
        b_field = mc_magnetic_field_magnitude(o, h, i);
        nu_c = mc_calc_cyclotron_freq(b_field);
        fill_emitted_photon(o, &ph_emit[0], h, i, nu_c, tmp->weight, i, rng, &position_phi);
        pr = rng->uniform_pos(rng) * (h->r0_size[i]) - (h->r0_size[i]) / 2.0;
        pr2 = rng->uniform_pos(rng) * (h->r1_size[i]) - (h->r1_size[i]) / 2.0;
        if (dims == MC_THREE) {
            pr3 = rng->uniform_pos(rng) * (h->r2_size[i]) - (h->r2_size[i]) / 2.0;
            mc_hydro_coord_to_mcrat(o, pos, h->r0[i] + pr, h->r1[i] + pr2, h->r2[i] + pr3);
        } else {
            mc_hydro_coord_to_mcrat(o, pos, h->r0[i] + pr, h->r1[i] + pr2, position_phi);
        }
        tmp->r0 = pos[0];
        tmp->r1 = pos[1];
        tmp->r2 = pos[2];
        idx = 0;
        (void)idx;
    }
    o->last_emit_weight = ph_weight_adjusted;
    mc_list_add(l, ph_emit, (size_t)ph_tot);
    free(ph_dens);
    free(ph_emit);
    return ph_tot;
}

/* Src/mc_cyclosynch.c:1571-1644 phAbsCyclosynch */
double mc_ph_abs_cyclosynch(mc_oracle *o, mc_photon_list *l, int *num_abs_ph, int *scatt_cyclosynch_num_ph,
                            const mc_hydro *h)
{
    int i, abs_ph_count = 0, synch_ph_count = 0;
    double nu_c, abs_count = 0, b_field;
    *scatt_cyclosynch_num_ph = 0;
    for (i = 0; i < l->list_capacity; i++) {
        mc_photon *ph = &l->photons[i];
        if ((ph->weight != 0) && (ph->nearest_block_index != -1)) {
            b_field = mc_magnetic_field_magnitude(o, h, ph->nearest_block_index);
            nu_c = mc_calc_cyclotron_freq(b_field);
            if ((ph->comv_p0 * C_LIGHT / PL_CONST <= nu_c) || (ph->type == MC_CS_POOL_PHOTON)) {
                abs_ph_count++;
                if ((ph->type != MC_INJECTED_PHOTON) && (ph->type != MC_UNABSORBED_CS_PHOTON)) {
                    if (ph->type == MC_CS_POOL_PHOTON) synch_ph_count++;
                } else {
                    abs_count += ph->weight;
                    ph->p0 = -1;
                }
                mc_list_set_null(l, i);
            } else {
                if ((ph->type == MC_COMPTONIZED_PHOTON) || (ph->type == MC_UNABSORBED_CS_PHOTON))
                    *scatt_cyclosynch_num_ph += 1;
            }
        }
    }
    (void)synch_ph_count;
    *num_abs_ph = abs_ph_count;
    return abs_count;
}

/* ============================================================================ */
/* rebinCyclosynchCompPhotons, Src/mc_cyclosynch.c:244-710                          */
/* ============================================================================ */
#define MC_RAD_TO_DEG (180.0 / M_PI) /* Src/mcrat.h:80-81 */
#define MC_DEG_TO_RAD (M_PI / 180.0)
#define MC_REBIN_ANG 0.5             /* CYCLOSYNCHROTRON_REBIN_ANG, Src/mcrat.h:313-316 */
#define MC_REBIN_ANG_PHI 10.0        /* CYCLOSYNCHROTRON_REBIN_ANG_PHI, :318-321 */

/* calculate_photon_position, :246-270 */
static void rebin_position(const mc_oracle *o, const mc_photon *ph, double *r, double *theta, double *phi)
{
    double x = ph->r0, y = ph->r1, z = ph->r2;
    *r = sqrt(x * x + y * y + z * z);
    if (*r < DBL_MIN) {
        *theta = 0.0;
        *phi = 0.0;
    } else {
        *theta = acos(z / *r);
        if (o->cfg.dimensions == MC_THREE) {
            double phi_rad = atan2(y, x);
            *phi = fmod(phi_rad * MC_RAD_TO_DEG + 360.0, 360.0);
        } else {
            *phi = 0;
        }
    }
}

static int rebin_eligible(const mc_photon *ph)
{
    return (ph->type != MC_NULL_PHOTON) && (ph->type != MC_CS_POOL_PHOTON) && (ph->type != MC_INJECTED_PHOTON);
}

int mc_rebin_cyclosynch_comp_photons(mc_oracle *o, mc_photon_list *l, int *num_cyclosynch_ph_emit,
                                     int *scatt_cyclosynch_num_ph, int max_photons)
{
    const int three = (o->cfg.dimensions == MC_THREE);
    int i, valid = 0, synch = 0, null_count = 0;
    double p0_min = DBL_MAX, p0_max = 0.0, th_min = DBL_MAX, th_max = 0.0, ph_min = DBL_MAX, ph_max = 0.0;
    /* collect_photon_statistics, :273-322 */
    for (i = 0; i < l->list_capacity; i++) {
        const mc_photon *ph = &l->photons[i];
        if (rebin_eligible(ph)) {
            double r, theta, phi = 0.0;
            if (ph->p0 > 0) {
                p0_min = fmin(p0_min, ph->p0);
                p0_max = fmax(p0_max, ph->p0);
                valid++;
            }
            rebin_position(o, ph, &r, &theta, &phi);
            th_min = fmin(th_min, theta);
            th_max = fmax(th_max, theta);
            if (three) {
                ph_min = fmin(ph_min, phi);
                ph_max = fmax(ph_max, phi);
            }
        }
        if (ph->type == MC_CS_POOL_PHOTON) synch++;
    }
    if (valid <= 0) return -1;
    {
        const double log_p0_min = (p0_min > 0 && p0_max > 0) ? log10(p0_min) : 0.0;
        const double log_p0_max = (p0_min > 0 && p0_max > 0) ? log10(p0_max) : 1.0;
        /* calculate_binning_params, :325-347 */
        const int num_bins = (int)(o->cfg.cs_rebin_e_perc * max_photons);
        const int num_bins_theta = (int)ceil((th_max - th_min) / (MC_REBIN_ANG * MC_DEG_TO_RAD));
        const int num_bins_phi = three ? (int)ceil((ph_max - ph_min) / MC_REBIN_ANG_PHI) : 1;
        int total_bins = num_bins_theta * num_bins;
        double *re, *rt, *rp;
        double (*st)[14];
        mc_photon *rebin_ph;
        if (three) total_bins *= num_bins_phi;
        if (total_bins > max_photons) return -1; /* :637-642 */
        if (num_bins <= 0 || num_bins_theta <= 0 || num_bins_phi <= 0) return -1; /* :352-355 */
        /* allocate_histograms, :350-392: uniform ranges with the upper edge nudged outward */
        re = (double *)malloc((size_t)(num_bins + 1) * sizeof(double));
        rt = (double *)malloc((size_t)(num_bins_theta + 1) * sizeof(double));
        rp = (double *)malloc((size_t)(num_bins_phi + 1) * sizeof(double));
        mc_hist_uniform_ranges(re, (size_t)num_bins, log_p0_min, log_p0_max + (log_p0_max - log_p0_min) * 1e-6);
        mc_hist_uniform_ranges(rt, (size_t)num_bins_theta, th_min, th_max + (th_max - th_min) * 1e-6);
        if (three) mc_hist_uniform_ranges(rp, (size_t)num_bins_phi, ph_min, ph_max + (ph_max - ph_min) * 1e-6);
        /* accumulate_bin_statistics, :448-500.  Columns: r, theta, phi_offset, s0..s3, scatt, weight, phi_dir,
         * theta_dir, energy, phi_pos */
        st = (double (*)[14])calloc((size_t)total_bins, sizeof(*st));
        for (i = 0; i < l->list_capacity; i++) {
            const mc_photon *ph = &l->photons[i];
            double r, theta, phi = 0.0, phi_dir, theta_dir, le;
            size_t ix = 0, iy = 0, iz = 0;
            long bin;
            if (!rebin_eligible(ph)) continue;
            rebin_position(o, ph, &r, &theta, &phi);
            le = log10(ph->p0);
            if (mc_hist_find((size_t)num_bins, re, le, &ix) == 0) mc_hist_find((size_t)num_bins_theta, rt, theta, &iy);
            if (three) {
                if (mc_hist_find((size_t)num_bins, re, le, &ix) == 0) mc_hist_find((size_t)num_bins_phi, rp, phi, &iz);
                if (mc_hist_find((size_t)num_bins_theta, rt, theta, &iy) == 0) mc_hist_find((size_t)num_bins_phi, rp, phi, &iz);
            }
            /* calculate_bin_index, :432-446 */
            if ((int)ix >= num_bins || (int)iy >= num_bins_theta)
                bin = -1;
            else if (three)
                bin = ((int)iz >= num_bins_phi) ? -1 : (long)iz * num_bins * num_bins_theta + (long)ix * num_bins_theta + (long)iy;
            else
                bin = (long)ix * num_bins_theta + (long)iy;
            if (bin < 0 || bin >= total_bins) {
                fprintf(stderr, "oracle rebin: photon %d maps to invalid bin index %ld\n", i, bin);
                exit(1); /* :469-472 */
            }
            st[bin][0] += r * ph->weight;
            st[bin][1] += theta * ph->weight;
            st[bin][2] += (atan2(ph->p2, ph->p1) - atan2(ph->r1, ph->r0)) * MC_RAD_TO_DEG * ph->weight;
            st[bin][3] += ph->s0 * ph->weight;
            st[bin][4] += ph->s1 * ph->weight;
            st[bin][5] += ph->s2 * ph->weight;
            st[bin][6] += ph->s3 * ph->weight;
            st[bin][7] += ph->num_scatt * ph->weight;
            st[bin][8] += ph->weight;
            phi_dir = fmod(atan2(ph->p2, ph->p1) * MC_RAD_TO_DEG + 360.0, 360.0);
            theta_dir = acos(ph->p3 / ph->p0) * MC_RAD_TO_DEG;
            st[bin][9] += phi_dir * ph->weight;
            st[bin][10] += theta_dir * ph->weight;
            st[bin][11] += ph->p0 * ph->weight;
            if (three) st[bin][12] += phi * ph->weight;
        }
        /* create_rebinned_photons, :503-598 */
        rebin_ph = (mc_photon *)calloc((size_t)total_bins, sizeof(mc_photon));
        for (i = 0; i < total_bins; i++) {
            const double *s = st[i];
            mc_photon *q = &rebin_ph[i];
            if (s[8] <= 0) {
                q->type = MC_NULL_PHOTON;
                q->weight = 0;
                q->nearest_block_index = -1;
                q->recalc_properties = 0;
                null_count++;
            } else {
                double avg_energy = s[11] / s[8], avg_phi_dir = s[9] / s[8], avg_theta_dir = s[10] / s[8];
                double avg_r = s[0] / s[8], avg_theta_pos = s[1] / s[8], pos_phi;
                q->type = MC_COMPTONIZED_PHOTON;
                q->weight = s[8];
                q->p0 = avg_energy;
                q->p1 = avg_energy * sin(avg_theta_dir * MC_DEG_TO_RAD) * cos(avg_phi_dir * MC_DEG_TO_RAD);
                q->p2 = avg_energy * sin(avg_theta_dir * MC_DEG_TO_RAD) * sin(avg_phi_dir * MC_DEG_TO_RAD);
                q->p3 = avg_energy * cos(avg_theta_dir * MC_DEG_TO_RAD);
                if (three) {
                    double avg_phi_pos = s[12] / s[8];
                    pos_phi = avg_phi_pos * MC_DEG_TO_RAD;
                } else {
                    double avg_phi_offset = s[2] / s[8];
                    pos_phi = (avg_phi_dir - avg_phi_offset) * MC_DEG_TO_RAD;
                }
                q->r0 = avg_r * sin(avg_theta_pos) * cos(pos_phi);
                q->r1 = avg_r * sin(avg_theta_pos) * sin(pos_phi);
                q->r2 = avg_r * cos(avg_theta_pos);
                q->s0 = s[3] / s[8];
                q->s1 = s[4] / s[8];
                q->s2 = s[5] / s[8];
                q->s3 = s[6] / s[8];
                q->num_scatt = (int)(s[7] / s[8] + 0.5);
                q->nearest_block_index = 0;
                q->recalc_properties = 1;
            }
        }
        for (i = 0; i < l->list_capacity; i++) {
            mc_photon *ph = &l->photons[i];
            if (ph->type == MC_UNABSORBED_CS_PHOTON || ph->type == MC_COMPTONIZED_PHOTON) mc_list_set_null(l, i);
        }
        mc_list_add(l, rebin_ph, (size_t)total_bins);
        free(rebin_ph);
        free(st);
        free(re);
        free(rt);
        free(rp);
        /* :680-684 */
        *scatt_cyclosynch_num_ph = total_bins - null_count;
        *num_cyclosynch_ph_emit = total_bins + synch - null_count;
    }
    return null_count;
}

/* ============================================================================ */
/* the scatter-frame while-loop, Src/mcrat.c:754-851                               */
/* (the rebin of :820-830 is mc_rebin_cyclosynch_comp_photons above; the loop stops */
/*  with st->iterations < 0 where the driver would call it)                        */
/* ============================================================================ */
void mc_run_frame(mc_oracle *o, mc_photon_list *l, const mc_hydro *h, mc_rng *rng, double time_now,
                  double remaining_time, long long max_iters, int find_nearest_grid_switch, double cs_r_inj,
                  double cs_ph_weight, int cs_max_photons, double cs_theta_min, double cs_theta_max,
                  mc_frame_stats *st)
{
    int frame_scatt_cnt = 0, frame_abs_cnt = 0, ph_scatt_index = 0, num_relocate = 0;
    int num_cs_emit = 0, scatt_cs_num = st->scatt_cyclosynch_num_ph;
    double time_step = 0;
    long long iters = 0, slots = 0;
    int need_rebin = 0;

    while (remaining_time > 0 && (max_iters < 0 || iters < max_iters)) {
        num_relocate += mc_find_containing_hydro_cell(o, l, h, find_nearest_grid_switch, rng);
        mc_calc_mean_free_path(o, l, h, rng);
        find_nearest_grid_switch = 0;
        slots += l->list_capacity;
        if (l->photons[l->sorted_indexes[0]].time_to_scatter < remaining_time) {
            time_step = mc_photon_event(o, l, remaining_time, h, &ph_scatt_index, &frame_scatt_cnt, &frame_abs_cnt, rng);
            time_now += time_step;
            remaining_time -= time_step;
            if (o->cfg.cyclosynch_switch) {
                mc_photon *sp = &l->photons[ph_scatt_index];
                if (sp->type == MC_CS_POOL_PHOTON) {
                    sp->type = MC_COMPTONIZED_PHOTON;
                    num_cs_emit += mc_photon_emit_cyclosynch(o, l, cs_r_inj, cs_ph_weight, cs_max_photons, cs_theta_min,
                                                             cs_theta_max, h, rng, 1, ph_scatt_index);
                    scatt_cs_num++;
                }
                if ((frame_scatt_cnt % 1000 == 0) && (frame_scatt_cnt != 0) && scatt_cs_num > cs_max_photons) {
                    need_rebin = 1;
                    o->iter++;
                    iters++;
                    break;
                }
            }
        } else {
            time_now += remaining_time;
            mc_update_photon_position(l, remaining_time);
            time_step = remaining_time;
            remaining_time = 0;
        }
        o->iter++;
        iters++;
    }
    st->iterations = need_rebin ? -iters : iters;
    st->scatterings = frame_scatt_cnt;
    st->relocations = num_relocate;
    st->photon_slots = slots;
    st->time_now = time_now;
    st->last_time_step = time_step;
    st->cs_emitted = num_cs_emit;
    st->scatt_cyclosynch_num_ph = scatt_cs_num;
}
