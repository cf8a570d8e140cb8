/* mcrat_io.c -- host-side configuration and output surface (include/mcrat_b200_io.h).
 *
 * mc.par and mcrat_input.h readers, and a self-contained writer / reader for the subset of the
 * HDF5 file format MCRaT's photon output uses.  Structures follow the HDF5 File Format
 * Specification, version 1.1 ("Disk Format: Level 0 / 1 / 2"):
 *   superblock version 0, symbol-table groups (B-tree v1 node + local heap + symbol table node),
 *   version-1 object headers with dataspace (0x0001), datatype (0x0003), fill value (0x0005), layout
 *   (0x0008) and symbol table (0x0011) messages.
 * The reference writes these files through libhdf5 (printPhotons, Src/mcrat_io.c:113-530;
 * dirFileMerge, :1239-1770); no libhdf5 exists in this build environment.
 */
#define _GNU_SOURCE
#include "../../include/mcrat_b200_io.h"

#include <ctype.h>
#include <errno.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

static __thread char g_err[512];

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

API const char *mcrat_b200_io_last_error(void) { return g_err; }

/* ============================================================================================
 * mc.par
 * ============================================================================================ */
/* The reference reads the file positionally (fgets / fscanf, Src/mcrat_io.c:1150-1232): block
 * header, blank line, then one value (or a row of values) per line with a trailing # comment.
 * This reader takes the non-blank, non-header lines in order, which accepts the same files. */
static int next_value_line(FILE *f, char *buf, size_t n)
{
    while (fgets(buf, (int)n, f)) {
        char *p = buf;
        while (*p && isspace((unsigned char)*p)) p++;
        if (*p == 0 || *p == '[') continue; /* blank line or [Block header] */
        char *hash = strchr(p, '#');
        if (hash) *hash = 0;
        memmove(buf, p, strlen(p) + 1);
        return 1;
    }
    return 0;
}

API int mcrat_b200_read_mc_par(const char *path, mcrat_b200_mc_par *out)
{
    if (!path || !out) return fail(MCRAT_IO_ERR_ARG, "read_mc_par: null argument");
    FILE *f = fopen(path, "r");
    if (!f) return fail(MCRAT_IO_ERR_OPEN, "read_mc_par: cannot open %s: %s", path, strerror(errno));
    char buf[2048];
    memset(out, 0, sizeof(*out));
    int rc = MCRAT_IO_OK;
#define NEED_LINE(what)                                                                     \
    if (!next_value_line(f, buf, sizeof(buf))) {                                            \
        rc = fail(MCRAT_IO_ERR_FORMAT, "read_mc_par: %s: missing line for %s", path, what); \
        goto done;                                                                          \
    }
#define NEED(cond, what)                                                                  \
    if (!(cond)) {                                                                        \
        rc = fail(MCRAT_IO_ERR_FORMAT, "read_mc_par: %s: cannot parse %s", path, what);   \
        goto done;                                                                        \
    }
    NEED_LINE("fps");
    NEED(sscanf(buf, "%lf", &out->fps) == 1, "fps");
    NEED_LINE("last frame");
    NEED(sscanf(buf, "%d", &out->last_frame) == 1, "last frame");
    NEED_LINE("r0 domain");
    NEED(sscanf(buf, "%lf %lf", &out->r0_domain[0], &out->r0_domain[1]) == 2, "r0 domain");
    NEED_LINE("r1 domain");
    NEED(sscanf(buf, "%lf %lf", &out->r1_domain[0], &out->r1_domain[1]) == 2, "r1 domain");
    NEED_LINE("r2 domain");
    NEED(sscanf(buf, "%lf %lf", &out->r2_domain[0], &out->r2_domain[1]) == 2, "r2 domain");
    NEED_LINE("theta_jmin");
    NEED(sscanf(buf, "%lf", &out->theta_jmin) == 1, "theta_jmin");
    NEED_LINE("theta_j");
    NEED(sscanf(buf, "%lf", &out->theta_j) == 1, "theta_j");
    NEED_LINE("number of angle bins");
    {
        double nb = 0;
        NEED(sscanf(buf, "%lf", &nb) == 1, "number of angle bins"); /* the reference reads it as a double (:1172) */
        out->n_theta_j = (int)nb;
        NEED(out->n_theta_j >= 1 && out->n_theta_j <= MCRAT_IO_MAX_ANGLE_BINS, "number of angle bins (1..64)");
    }
    for (int row = 0; row < 3; ++row) {
        static const char *names[3] = {"injection start frames", "numbers of injection frames", "injection radii"};
        NEED_LINE(names[row]);
        char *save = NULL, *tok = strtok_r(buf, " \t\r\n", &save);
        for (int i = 0; i < out->n_theta_j; ++i) {
            NEED(tok != NULL, names[row]);
            if (row == 0)
                out->frm0[i] = (int)strtol(tok, NULL, 10);
            else if (row == 1)
                out->frm2[i] = (int)strtol(tok, NULL, 10) + out->frm0[i]; /* Src/mcrat_io.c:1201 */
            else
                out->inj_radius[i] = (double)strtof(tok, NULL); /* strtof, as the reference (:1212) */
            tok = strtok_r(NULL, " \t\r\n", &save);
        }
    }
    NEED_LINE("spectrum type");
    out->spect = buf[0];
    NEED(out->spect == 'w' || out->spect == 'b', "spectrum type (w or b)");
    NEED_LINE("min photons");
    NEED(sscanf(buf, "%d", &out->min_photons) == 1, "min photons");
    NEED_LINE("max photons");
    NEED(sscanf(buf, "%d", &out->max_photons) == 1, "max photons");
    NEED_LINE("initialise / continue");
    out->restart = buf[0];
    NEED(out->restart == 'i' || out->restart == 'c', "restart flag (i or c)");
done:
    fclose(f);
    return rc;
#undef NEED
#undef NEED_LINE
}

/* ============================================================================================
 * mcrat_input.h
 * ============================================================================================ */
typedef struct {
    const char *name;
    int value;
} sym_t;

static const sym_t SYMS[] = {
    {"ON", 1}, {"OFF", 0},
    {"FLASH", 0}, {"PLUTO_CHOMBO", 1}, {"PLUTO", 2}, {"RIKEN", 3},
    {"SCIENCE", 0}, {"CYLINDRICAL_OUTFLOW", 1}, {"SPHERICAL_OUTFLOW", 2}, {"STRUCTURED_SPHERICAL_OUTFLOW", 3},
    {"CARTESIAN", 0}, {"SPHERICAL", 1}, {"CYLINDRICAL", 2}, {"POLAR", 3},
    {"TWO", 0}, {"TWO_POINT_FIVE", 1}, {"THREE", 2},
    {"INTERNAL_E", 0}, {"TOTAL_E", 1}, {"SIMULATION", 2},
    {"DIRECT", 1}, {"TABLE", 2},
    {"POWERLAW", 1}, {"BROKENPOWERLAW", 2},
    {NULL, 0}};

static int sym_value(const char *tok, int *out)
{
    for (const sym_t *s = SYMS; s->name; ++s)
        if (strcmp(tok, s->name) == 0) {
            *out = s->value;
            return 1;
        }
    char *end = NULL;
    long v = strtol(tok, &end, 10);
    if (end && *end == 0 && end != tok) {
        *out = (int)v;
        return 1;
    }
    return 0;
}

static void copy_string_value(char *dst, size_t n, const char *val)
{
    const char *a = strchr(val, '"');
    const char *b = a ? strchr(a + 1, '"') : NULL;
    if (a && b) {
        size_t len = (size_t)(b - a - 1);
        if (len >= n) len = n - 1;
        memcpy(dst, a + 1, len);
        dst[len] = 0;
    } else {
        snprintf(dst, n, "%s", val);
    }
}

API int mcrat_b200_config_from_input_header(const char *path, mcrat_b200_config *cfg, mcrat_b200_io_switches *sw)
{
    if (!path || !cfg) return fail(MCRAT_IO_ERR_ARG, "config_from_input_header: null argument");
    FILE *f = fopen(path, "r");
    if (!f) return fail(MCRAT_IO_ERR_OPEN, "config_from_input_header: cannot open %s: %s", path, strerror(errno));
    mcrat_b200_io_switches local;
    if (!sw) sw = &local;
    memset(sw, 0, sizeof(*sw));
    int have_sim = 0, have_dim = 0, have_geo = 0, have_l = 0, have_d = 0, have_par = 0, have_b = 0, have_eps = 0;
    int have_tau = 0, nonthermal = 0, in_block_comment = 0;
    int stokes = 0, cs = 0, tau = 1, bcalc = 1, dims = -1, geom = -1;
    double eps = 0.5;
    char line[2048];
    while (fgets(line, sizeof(line), f)) {
        /* strip comments: block comments may span lines, // runs to the end of the line */
        char clean[2048];
        size_t o = 0;
        for (size_t i = 0; line[i] && o + 1 < sizeof(clean); ++i) {
            if (in_block_comment) {
                if (line[i] == '*' && line[i + 1] == '/') {
                    in_block_comment = 0;
                    i++;
                }
                continue;
            }
            if (line[i] == '/' && line[i + 1] == '*') {
                in_block_comment = 1;
                i++;
                continue;
            }
            if (line[i] == '/' && line[i + 1] == '/') break;
            clean[o++] = line[i];
        }
        clean[o] = 0;
        char *p = clean;
        while (*p && isspace((unsigned char)*p)) p++;
        if (*p != '#') continue;
        p++;
        while (*p && isspace((unsigned char)*p)) p++;
        if (strncmp(p, "define", 6) != 0 || !isspace((unsigned char)p[6])) continue;
        p += 6;
        char name[128] = "", val[1024] = "";
        if (sscanf(p, " %127s %1023[^\n]", name, val) < 1) continue;
        size_t vl = strlen(val);
        while (vl && isspace((unsigned char)val[vl - 1])) val[--vl] = 0;
        int iv = 0;
        const int is_sym = sym_value(val, &iv);
#define SWITCH(NAME, target, flag)                                                                           \
    if (strcmp(name, NAME) == 0) {                                                                           \
        if (!is_sym) {                                                                                       \
            fclose(f);                                                                                       \
            return fail(MCRAT_IO_ERR_FORMAT, "config_from_input_header: %s: unknown value '%s' for %s", path, val, NAME); \
        }                                                                                                    \
        target = iv;                                                                                         \
        flag = 1;                                                                                            \
        continue;                                                                                            \
    }
        int dummy = 0;
        SWITCH("SIM_SWITCH", sw->sim_switch, have_sim)
        SWITCH("SIMULATION_TYPE", sw->simulation_type, dummy)
        SWITCH("DIMENSIONS", dims, have_dim)
        SWITCH("GEOMETRY", geom, have_geo)
        SWITCH("STOKES_SWITCH", stokes, dummy)
        SWITCH("COMV_SWITCH", sw->comv_switch, dummy)
        SWITCH("SAVE_TYPE", sw->save_type, dummy)
        SWITCH("CYCLOSYNCHROTRON_SWITCH", cs, dummy)
        SWITCH("TAU_CALCULATION", tau, have_tau)
        SWITCH("B_FIELD_CALC", bcalc, have_b)
        SWITCH("NONTHERMAL_E_DIST", nonthermal, dummy)
#undef SWITCH
        (void)dummy;
        if (strcmp(name, "EPSILON_B") == 0) {
            eps = strtod(val, NULL);
            have_eps = 1;
        } else if (strcmp(name, "HYDRO_L_SCALE") == 0) {
            have_l = 1;
        } else if (strcmp(name, "HYDRO_D_SCALE") == 0) {
            have_d = 1;
        } else if (strcmp(name, "MCPAR") == 0) {
            copy_string_value(sw->mcpar, sizeof(sw->mcpar), val);
            have_par = 1;
        } else if (strcmp(name, "MC_PATH") == 0) {
            copy_string_value(sw->mc_path, sizeof(sw->mc_path), val);
        } else if (strcmp(name, "FILEPATH") == 0) {
            copy_string_value(sw->filepath, sizeof(sw->filepath), val);
        } else if (strcmp(name, "FILEROOT") == 0) {
            copy_string_value(sw->fileroot, sizeof(sw->fileroot), val);
        }
    }
    fclose(f);
    /* the #error checks of Src/mcrat.h:404-427 */
    if (!have_sim) return fail(MCRAT_IO_ERR_FORMAT, "Need to define hydro simulation type in mcrat_input.h file using SIM_SWITCH");
    if (!have_dim) return fail(MCRAT_IO_ERR_FORMAT, "Need to define hydro simulation dimensions in mcrat_input.h file using DIMENSIONS");
    if (!have_geo) return fail(MCRAT_IO_ERR_FORMAT, "Need to define hydro simulation geometry in mcrat_input.h file using GEOMETRY");
    if (!have_l) return fail(MCRAT_IO_ERR_FORMAT, "Need to define hydro simulation length scaling in mcrat_input.h file using HYDRO_L_SCALE");
    if (!have_d) return fail(MCRAT_IO_ERR_FORMAT, "Need to define hydro simulation density scaling in mcrat_input.h file using HYDRO_D_SCALE");
    if (!have_par) return fail(MCRAT_IO_ERR_FORMAT, "Need to define name of MCRaT parameter file in mcrat_input.h file using MCPAR");
    /* Src/mcrat.h:262-280 */
    if (nonthermal) {
        if (!have_tau) tau = 2;
        if (tau == 1) return fail(MCRAT_IO_ERR_FORMAT, "NONTHERMAL_E_DIST cannot be set while TAU_CALCULATION = DIRECT.");
        return fail(MCRAT_IO_ERR_FORMAT, "NONTHERMAL_E_DIST is outside the scope of the B200 hot path (DESIGN.md section 7)");
    }
    /* Src/mcrat.h:319-333: B_FIELD_CALC defaults to TOTAL_E, EPSILON_B to 0.5 */
    if (!have_b) bcalc = 1;
    if (!have_eps) eps = 0.5;
    if (dims < 0 || dims > 2 || geom < 0 || geom > 3) return fail(MCRAT_IO_ERR_FORMAT, "bad DIMENSIONS / GEOMETRY");
    if (dims != 2 && geom == 3) return fail(MCRAT_IO_ERR_FORMAT, "POLAR geometry exists only in 3-D (Src/geometry.c:20-58)");
    if (dims == 2 && geom == 2) return fail(MCRAT_IO_ERR_FORMAT, "CYLINDRICAL geometry exists only in 2-D / 2.5-D (Src/geometry.c:20-58)");
    cfg->abi_version = MCRAT_B200_ABI_VERSION;
    cfg->dimensions = dims;
    cfg->geometry = geom;
    cfg->stokes_switch = stokes;
    cfg->tau_calculation = tau;
    cfg->cyclosynch_switch = cs;
    cfg->b_field_calc = bcalc;
    cfg->epsilon_b = eps;
    sw->stokes_switch = stokes;
    return MCRAT_IO_OK;
}

/* ============================================================================================
 * HDF5 subset: in-memory model
 * ============================================================================================ */
typedef struct {
    char name[64];
    int is_i8;      /* 0: IEEE F64LE, 1: STD_I8LE (H5T_NATIVE_CHAR on the reference's platforms) */
    size_t n;       /* 1-D length */
    void *data;     /* n doubles or n signed chars */
    size_t chunk;   /* 0: contiguous, fixed size (mcdata files, Src/mcrat_io.c:1649); > 0: chunked with this many elements per
                     * chunk and an unlimited maximum dimension (the per-rank files, Src/mcrat_io.c:140, 254-263) */
} h5_dset;

typedef struct {
    char name[64];  /* "" = the root group itself */
    int nd, capd;
    h5_dset *d;
} h5_group;

typedef struct {
    int ng, capg;
    h5_group *g;    /* g[0] is always the root group */
} h5_file;

static void h5_free(h5_file *f)
{
    for (int i = 0; i < f->ng; ++i) {
        for (int k = 0; k < f->g[i].nd; ++k) free(f->g[i].d[k].data);
        free(f->g[i].d);
    }
    free(f->g);
    memset(f, 0, sizeof(*f));
}

static h5_group *h5_group_get(h5_file *f, const char *name, int create)
{
    for (int i = 0; i < f->ng; ++i)
        if (strcmp(f->g[i].name, name) == 0) return &f->g[i];
    if (!create) return NULL;
    if (f->ng == f->capg) {
        int nc = f->capg ? 2 * f->capg : 8;
        h5_group *ng = (h5_group *)realloc(f->g, (size_t)nc * sizeof(h5_group));
        if (!ng) return NULL;
        f->g = ng;
        f->capg = nc;
    }
    h5_group *g = &f->g[f->ng++];
    memset(g, 0, sizeof(*g));
    snprintf(g->name, sizeof(g->name), "%s", name);
    return g;
}

static h5_dset *h5_dset_get(h5_group *g, const char *name, int create)
{
    for (int k = 0; k < g->nd; ++k)
        if (strcmp(g->d[k].name, name) == 0) return &g->d[k];
    if (!create) return NULL;
    if (g->nd == g->capd) {
        int nc = g->capd ? 2 * g->capd : 24;
        h5_dset *nd = (h5_dset *)realloc(g->d, (size_t)nc * sizeof(h5_dset));
        if (!nd) return NULL;
        g->d = nd;
        g->capd = nc;
    }
    h5_dset *d = &g->d[g->nd++];
    memset(d, 0, sizeof(*d));
    snprintf(d->name, sizeof(d->name), "%s", name);
    return d;
}

/* append n elements to a dataset (creating it); the element type is fixed by the first append */
static int h5_dset_append(h5_group *g, const char *name, int is_i8, const void *src, size_t n)
{
    h5_dset *d = h5_dset_get(g, name, 1);
    if (!d) return MCRAT_IO_ERR_NOMEM;
    if (d->n == 0 && d->data == NULL) d->is_i8 = is_i8;
    if (d->is_i8 != is_i8) return fail(MCRAT_IO_ERR_FORMAT, "dataset %s: element type mismatch on append", name);
    const size_t es = is_i8 ? 1 : 8;
    void *nd = realloc(d->data, (d->n + n) * es + 8);
    if (!nd) return MCRAT_IO_ERR_NOMEM;
    d->data = nd;
    if (n) memcpy((char *)d->data + d->n * es, src, n * es);
    d->n += n;
    return MCRAT_IO_OK;
}

/* ============================================================================================
 * HDF5 subset: writer
 * ============================================================================================ */
typedef struct {
    unsigned char *b;
    size_t n, cap;
} wbuf;

static int wb_reserve(wbuf *w, size_t upto)
{
    if (upto <= w->cap) return 1;
    size_t nc = w->cap ? w->cap : 4096;
    while (nc < upto) nc *= 2;
    unsigned char *nb = (unsigned char *)realloc(w->b, nc);
    if (!nb) return 0;
    memset(nb + w->cap, 0, nc - w->cap);
    w->b = nb;
    w->cap = nc;
    return 1;
}

static void put_le(unsigned char *p, uint64_t v, int bytes)
{
    for (int i = 0; i < bytes; ++i) p[i] = (unsigned char)(v >> (8 * i));
}

static size_t align8(size_t x) { return (x + 7) & ~(size_t)7; }

#define H5_UNDEF 0xffffffffffffffffull
#define H5_INTERNAL_K 16

/* bytes of the structures of one group with `n` links, leaf-node rank K */
static size_t snod_size(int K) { return 8 + (size_t)(2 * K) * 40; }
static size_t btree_size(void) { return 24 + (size_t)(2 * H5_INTERNAL_K + 1) * 8 + (size_t)(2 * H5_INTERNAL_K) * 8; }

typedef struct {
    const char *name;
    uint64_t ohdr;      /* object header address */
    int is_group;
    uint64_t btree, heap; /* scratch-pad of a group entry */
} link_t;

static int link_cmp(const void *a, const void *b) { return strcmp(((const link_t *)a)->name, ((const link_t *)b)->name); }

/* writes object header + heap + B-tree + SNOD of one group at `at`; returns the end offset.
 * *ohdr, *btree, *heap receive the addresses the parent's symbol-table entry caches. */
static size_t write_group(wbuf *w, size_t at, link_t *links, int n, int K, uint64_t *ohdr_out, uint64_t *btree_out, uint64_t *heap_out)
{
    qsort(links, (size_t)n, sizeof(link_t), link_cmp);
    /* local heap data segment: "" at offset 0, then the names, then one free block */
    size_t names = 8;
    for (int i = 0; i < n; ++i) names += align8(strlen(links[i].name) + 1);
    const size_t heap_data_size = names + 32;
    const size_t a_ohdr = at;
    const size_t a_heap = a_ohdr + 16 + 24 + 8; /* prefix + symbol-table message + NIL message header */
    const size_t a_heap_data = a_heap + 32;
    const size_t a_btree = a_heap_data + heap_data_size;
    const size_t a_snod = a_btree + btree_size();
    const size_t end = a_snod + snod_size(K);
    if (!wb_reserve(w, end)) return 0;
    unsigned char *b = w->b;
    /* object header, version 1 */
    b[a_ohdr] = 1;
    put_le(b + a_ohdr + 2, 2, 2);  /* two messages (symbol table + NIL), like libhdf5 */
    put_le(b + a_ohdr + 4, 1, 4);  /* reference count */
    put_le(b + a_ohdr + 8, 32, 4); /* header data size */
    unsigned char *m = b + a_ohdr + 16;
    put_le(m, 0x0011, 2);
    put_le(m + 2, 16, 2);
    m[4] = 0;
    put_le(m + 8, a_btree, 8);
    put_le(m + 16, a_heap, 8);
    put_le(m + 24, 0x0000, 2); /* NIL message, size 0 */
    /* local heap */
    memcpy(b + a_heap, "HEAP", 4);
    put_le(b + a_heap + 8, heap_data_size, 8);
    put_le(b + a_heap + 16, names, 8); /* head of the free list */
    put_le(b + a_heap + 24, a_heap_data, 8);
    size_t off = 8;
    uint64_t *name_off = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n ? n : 1));
    if (!name_off) return 0;
    for (int i = 0; i < n; ++i) {
        name_off[i] = off;
        memcpy(b + a_heap_data + off, links[i].name, strlen(links[i].name));
        off += align8(strlen(links[i].name) + 1);
    }
    put_le(b + a_heap_data + names, 1, 8);       /* next free block: none (H5HL_FREE_NULL) */
    put_le(b + a_heap_data + names + 8, 32, 8);  /* size of this free block */
    /* B-tree v1 node, type 0 (group), leaf */
    memcpy(b + a_btree, "TREE", 4);
    b[a_btree + 4] = 0;
    b[a_btree + 5] = 0;
    put_le(b + a_btree + 6, n ? 1 : 0, 2);
    put_le(b + a_btree + 8, H5_UNDEF, 8);
    put_le(b + a_btree + 16, H5_UNDEF, 8);
    put_le(b + a_btree + 24, 0, 8);                       /* key 0: "" */
    put_le(b + a_btree + 32, a_snod, 8);                  /* child 0 */
    put_le(b + a_btree + 40, n ? name_off[n - 1] : 0, 8); /* key 1: the largest name */
    /* symbol table node */
    memcpy(b + a_snod, "SNOD", 4);
    b[a_snod + 4] = 1;
    put_le(b + a_snod + 6, (uint64_t)n, 2);
    for (int i = 0; i < n; ++i) {
        unsigned char *e = b + a_snod + 8 + 40 * (size_t)i;
        put_le(e, name_off[i], 8);
        put_le(e + 8, links[i].ohdr, 8);
        put_le(e + 16, links[i].is_group ? 1 : 0, 4);
        if (links[i].is_group) {
            put_le(e + 24, links[i].btree, 8);
            put_le(e + 32, links[i].heap, 8);
        }
    }
    free(name_off);
    *ohdr_out = a_ohdr;
    *btree_out = a_btree;
    *heap_out = a_heap;
    if (end > w->n) w->n = end;
    return end;
}

#define DSET_OHDR_SIZE (16 + (8 + 8) + (8 + 24) + (8 + 16) + (8 + 24))
#define DSET_OHDR_SIZE_CHUNKED (16 + (8 + 8) + (8 + 24) + (8 + 24) + (8 + 24))
#define H5_ISTORE_K 32 /* chunk B-tree: superblock version 0 has no field for it, the library's default applies */
#define CHUNK_KEY_SIZE 24 /* chunk bytes (4), filter mask (4), offsets of a 1-D dataset + the element dimension (2 x 8) */

static size_t dset_ohdr_size(const h5_dset *d) { return d->chunk ? DSET_OHDR_SIZE_CHUNKED : DSET_OHDR_SIZE; }
static size_t chunk_node_size(void) { return 24 + (size_t)(2 * H5_ISTORE_K + 1) * CHUNK_KEY_SIZE + (size_t)(2 * H5_ISTORE_K) * 8; }

/* dataset object header at `at` pointing at raw data (addr, bytes) */
static size_t write_dataset_header(wbuf *w, size_t at, const h5_dset *d, uint64_t data_addr)
{
    const size_t end = at + dset_ohdr_size(d);
    if (!wb_reserve(w, end)) return 0;
    unsigned char *b = w->b + at;
    b[0] = 1;
    put_le(b + 2, 4, 2);
    put_le(b + 4, 1, 4);
    put_le(b + 8, dset_ohdr_size(d) - 16, 4);
    unsigned char *m = b + 16;
    /* fill value (old libraries' version-1 form: allocation late, fill time if-set, default value) */
    put_le(m, 0x0005, 2);
    put_le(m + 2, 8, 2);
    m[4] = 1;
    m[8] = 1; m[9] = 2; m[10] = 2; m[11] = 1;
    m += 16;
    /* datatype */
    put_le(m, 0x0003, 2);
    put_le(m + 2, 24, 2);
    m[4] = 1;
    if (!d->is_i8) {
        m[8] = 0x11; m[9] = 0x20; m[10] = 0x3f; m[11] = 0;  /* class 1 (float) v1; LE, msb-implied mantissa; sign bit 63 */
        put_le(m + 12, 8, 4);
        put_le(m + 16, 0, 2);   /* bit offset */
        put_le(m + 18, 64, 2);  /* precision */
        m[20] = 52;             /* exponent location */
        m[21] = 11;             /* exponent size */
        m[22] = 0;              /* mantissa location */
        m[23] = 52;             /* mantissa size */
        put_le(m + 24, 1023, 4);
    } else {
        m[8] = 0x10; m[9] = 0x08; m[10] = 0; m[11] = 0;     /* class 0 (fixed point) v1; LE, signed */
        put_le(m + 12, 1, 4);
        put_le(m + 16, 0, 2);
        put_le(m + 18, 8, 2);
    }
    m += 32;
    if (d->chunk) {
        /* dataspace, version 1, rank 1, maximum dimension present and unlimited (H5S_UNLIMITED, Src/mcrat_io.c:140) */
        put_le(m, 0x0001, 2);
        put_le(m + 2, 24, 2);
        m[8] = 1; m[9] = 1; m[10] = 1;
        put_le(m + 16, d->n, 8);
        put_le(m + 24, H5_UNDEF, 8);
        m += 32;
        /* data layout, version 3, chunked: dimensionality rank + 1, chunk B-tree, chunk dimensions (elements, element size) */
        put_le(m, 0x0008, 2);
        put_le(m + 2, 24, 2);
        m[8] = 3; m[9] = 2; m[10] = 2;
        put_le(m + 11, d->n ? data_addr : H5_UNDEF, 8); /* data_addr: the root node of the chunk B-tree */
        put_le(m + 19, (uint64_t)d->chunk, 4);
        put_le(m + 23, (uint64_t)(d->is_i8 ? 1 : 8), 4);
    } else {
        /* dataspace, version 1, rank 1, no maximum dimensions */
        put_le(m, 0x0001, 2);
        put_le(m + 2, 16, 2);
        m[8] = 1; m[9] = 1; m[10] = 0;
        put_le(m + 16, d->n, 8);
        m += 24;
        /* data layout, version 3, contiguous */
        put_le(m, 0x0008, 2);
        put_le(m + 2, 24, 2);
        m[8] = 3; m[9] = 1;
        put_le(m + 10, d->n ? data_addr : H5_UNDEF, 8);
        put_le(m + 18, (uint64_t)d->n * (d->is_i8 ? 1 : 8), 8);
    }
    if (end > w->n) w->n = end;
    return end;
}

/* Raw data of a chunked dataset at `at`: the chunks one after the other (the last one zero-filled to full size, as the
 * library allocates it), then the version-1 B-tree (node type 1) that indexes them -- leaves of up to 2K entries linked
 * left to right, internal levels above them while more than one node remains.  Returns the new cursor, *root_out = address
 * of the root node. */
static size_t write_chunks(wbuf *w, size_t at, const h5_dset *d, uint64_t *root_out)
{
    const size_t es = d->is_i8 ? 1 : 8, cbytes = d->chunk * es;
    const size_t nchunks = (d->n + d->chunk - 1) / d->chunk;
    const size_t data_at = at;
    if (!wb_reserve(w, at + nchunks * cbytes + 8)) return 0;
    memset(w->b + at, 0, nchunks * cbytes);
    memcpy(w->b + at, d->data, d->n * es);
    at = align8(at + nchunks * cbytes);
    /* level by level: entry e of a level = (first chunk index it covers, address) */
    size_t nent = nchunks;
    uint64_t *first = (uint64_t *)malloc(sizeof(uint64_t) * nent), *addr = (uint64_t *)malloc(sizeof(uint64_t) * nent);
    if (!first || !addr) {
        free(first);
        free(addr);
        return 0;
    }
    for (size_t c = 0; c < nchunks; ++c) {
        first[c] = c;
        addr[c] = data_at + c * cbytes;
    }
    const size_t per = 2 * H5_ISTORE_K, nsz = chunk_node_size();
    for (int level = 0;; ++level) {
        const size_t nnodes = (nent + per - 1) / per;
        if (!wb_reserve(w, at + nnodes * nsz)) {
            free(first);
            free(addr);
            return 0;
        }
        memset(w->b + at, 0, nnodes * nsz);
        for (size_t q = 0; q < nnodes; ++q) {
            unsigned char *b = w->b + at + q * nsz;
            const size_t e0 = q * per, used = (nent - e0 < per) ? nent - e0 : per;
            memcpy(b, "TREE", 4);
            b[4] = 1;
            b[5] = (unsigned char)level;
            put_le(b + 6, used, 2);
            put_le(b + 8, q ? at + (q - 1) * nsz : H5_UNDEF, 8);
            put_le(b + 16, (q + 1 < nnodes) ? at + (q + 1) * nsz : H5_UNDEF, 8);
            unsigned char *k = b + 24;
            for (size_t e = 0; e < used; ++e, k += CHUNK_KEY_SIZE + 8) {
                put_le(k, cbytes, 4);                         /* chunk size in bytes (no filters) */
                put_le(k + 4, 0, 4);                          /* filter mask */
                put_le(k + 8, first[e0 + e] * d->chunk, 8);   /* element offset along the dataset's dimension */
                put_le(k + 16, 0, 8);                         /* ... along the element dimension */
                put_le(k + CHUNK_KEY_SIZE, addr[e0 + e], 8);
            }
            /* the key after the last child: the far corner of the last chunk this node covers */
            const size_t next_first = (e0 + used < nent) ? first[e0 + used] : nchunks;
            put_le(k, 0, 4);
            put_le(k + 4, 0, 4);
            put_le(k + 8, next_first * d->chunk, 8);
            put_le(k + 16, es, 8);
        }
        for (size_t q = 0; q < nnodes; ++q) { /* the next level's entries: one per node of this one */
            first[q] = first[q * per];
            addr[q] = at + q * nsz;
        }
        const size_t level_at = at;
        at += nnodes * nsz;
        if (at > w->n) w->n = at;
        nent = nnodes;
        if (nnodes == 1) {
            *root_out = level_at;
            break;
        }
    }
    free(first);
    free(addr);
    return at;
}

static int h5_write(const char *path, h5_file *f)
{
    wbuf w = {0};
    int maxlinks = 0;
    for (int i = 0; i < f->ng; ++i) {
        int nl = f->g[i].nd + (i == 0 ? f->ng - 1 : 0);
        if (nl > maxlinks) maxlinks = nl;
    }
    int K = (maxlinks + 1) / 2;
    if (K < 4) K = 4;
    if (K > 32767) return fail(MCRAT_IO_ERR_FORMAT, "too many links in one group (%d)", maxlinks);
    size_t at = 96; /* superblock */
    int rc = MCRAT_IO_OK;
    /* pass 1: addresses of dataset headers and raw data.  Order in the file: dataset headers, raw
     * data, then the groups (their structures need the addresses of what they link to). */
    size_t nd_total = 0;
    for (int i = 0; i < f->ng; ++i) nd_total += (size_t)f->g[i].nd;
    uint64_t *hdr_addr = (uint64_t *)malloc(sizeof(uint64_t) * (nd_total ? nd_total : 1));
    if (!hdr_addr) return MCRAT_IO_ERR_NOMEM;
    size_t k = 0, cursor = at;
    for (int i = 0; i < f->ng; ++i)
        for (int j = 0; j < f->g[i].nd; ++j) {
            hdr_addr[k++] = cursor;
            cursor += dset_ohdr_size(&f->g[i].d[j]);
        }
    cursor = align8(cursor);
    k = 0;
    for (int i = 0; i < f->ng && rc == MCRAT_IO_OK; ++i)
        for (int j = 0; j < f->g[i].nd; ++j) {
            const h5_dset *d = &f->g[i].d[j];
            const size_t bytes = d->n * (d->is_i8 ? 1 : 8);
            if (d->chunk && d->n) {
                uint64_t root = 0;
                const size_t next = write_chunks(&w, cursor, d, &root);
                if (!next || !write_dataset_header(&w, hdr_addr[k], d, root)) {
                    rc = MCRAT_IO_ERR_NOMEM;
                    break;
                }
                cursor = align8(next);
                if (cursor > w.n) w.n = cursor;
                k++;
                continue;
            }
            if (!write_dataset_header(&w, hdr_addr[k], d, cursor) || !wb_reserve(&w, cursor + bytes + 8)) {
                rc = MCRAT_IO_ERR_NOMEM;
                break;
            }
            if (bytes) memcpy(w.b + cursor, d->data, bytes);
            cursor = align8(cursor + bytes);
            if (cursor > w.n) w.n = cursor;
            k++;
        }
    /* sub-groups, then the root group */
    uint64_t root_ohdr = 0, root_btree = 0, root_heap = 0;
    link_t *links = (link_t *)malloc(sizeof(link_t) * (size_t)(maxlinks ? maxlinks : 1));
    uint64_t *g_ohdr = (uint64_t *)calloc((size_t)f->ng, sizeof(uint64_t));
    uint64_t *g_bt = (uint64_t *)calloc((size_t)f->ng, sizeof(uint64_t));
    uint64_t *g_hp = (uint64_t *)calloc((size_t)f->ng, sizeof(uint64_t));
    if (!links || !g_ohdr || !g_bt || !g_hp) rc = MCRAT_IO_ERR_NOMEM;
    if (rc == MCRAT_IO_OK) {
        size_t base = 0;
        for (int i = 0; i < f->ng; ++i) {
            if (i == 0) {
                base += (size_t)f->g[0].nd;
                continue;
            }
            for (int j = 0; j < f->g[i].nd; ++j) {
                links[j].name = f->g[i].d[j].name;
                links[j].ohdr = hdr_addr[base + (size_t)j];
                links[j].is_group = 0;
                links[j].btree = links[j].heap = 0;
            }
            base += (size_t)f->g[i].nd;
            cursor = write_group(&w, cursor, links, f->g[i].nd, K, &g_ohdr[i], &g_bt[i], &g_hp[i]);
            if (!cursor) {
                rc = MCRAT_IO_ERR_NOMEM;
                break;
            }
        }
    }
    if (rc == MCRAT_IO_OK) {
        int n = 0;
        for (int j = 0; j < f->g[0].nd; ++j, ++n) {
            links[n].name = f->g[0].d[j].name;
            links[n].ohdr = hdr_addr[j];
            links[n].is_group = 0;
            links[n].btree = links[n].heap = 0;
        }
        for (int i = 1; i < f->ng; ++i, ++n) {
            links[n].name = f->g[i].name;
            links[n].ohdr = g_ohdr[i];
            links[n].is_group = 1;
            links[n].btree = g_bt[i];
            links[n].heap = g_hp[i];
        }
        cursor = write_group(&w, cursor, links, n, K, &root_ohdr, &root_btree, &root_heap);
        if (!cursor) rc = MCRAT_IO_ERR_NOMEM;
    }
    if (rc == MCRAT_IO_OK) {
        /* superblock, version 0 */
        static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        unsigned char *b = w.b;
        memcpy(b, sig, 8);
        b[13] = 8; /* size of offsets */
        b[14] = 8; /* size of lengths */
        put_le(b + 16, (uint64_t)K, 2);
        put_le(b + 18, H5_INTERNAL_K, 2);
        put_le(b + 24, 0, 8);        /* base address */
        put_le(b + 32, H5_UNDEF, 8); /* free-space info */
        put_le(b + 40, w.n, 8);      /* end of file */
        put_le(b + 48, H5_UNDEF, 8); /* driver info */
        put_le(b + 56, 0, 8);        /* root entry: link name offset */
        put_le(b + 64, root_ohdr, 8);
        put_le(b + 72, 1, 4);        /* cache type 1: group */
        put_le(b + 80, root_btree, 8);
        put_le(b + 88, root_heap, 8);
        char tmp[1024];
        snprintf(tmp, sizeof(tmp), "%s.tmp", path);
        FILE *fp = fopen(tmp, "wb");
        if (!fp) {
            rc = fail(MCRAT_IO_ERR_OPEN, "cannot create %s: %s", tmp, strerror(errno));
        } else {
            if (fwrite(w.b, 1, w.n, fp) != w.n) rc = fail(MCRAT_IO_ERR_OPEN, "short write on %s", tmp);
            if (fclose(fp) != 0 && rc == MCRAT_IO_OK) rc = fail(MCRAT_IO_ERR_OPEN, "cannot close %s", tmp);
            if (rc == MCRAT_IO_OK && rename(tmp, path) != 0) rc = fail(MCRAT_IO_ERR_OPEN, "cannot rename %s: %s", tmp, strerror(errno));
        }
    }
    free(links);
    free(g_ohdr);
    free(g_bt);
    free(g_hp);
    free(hdr_addr);
    free(w.b);
    return rc;
}

/* ============================================================================================
 * HDF5 subset: reader
 * ============================================================================================ */
typedef struct {
    unsigned char *b;
    size_t n;
    uint64_t base;
    int leafK, intK;
} rfile;

static uint64_t get_le(const unsigned char *p, int bytes)
{
    uint64_t v = 0;
    for (int i = 0; i < bytes; ++i) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

static const unsigned char *rptr(const rfile *r, uint64_t addr, size_t len)
{
    if (addr == H5_UNDEF) return NULL;
    uint64_t a = addr + r->base;
    if (a > r->n || len > r->n - a) return NULL;
    return r->b + a;
}

typedef struct {
    int have_symtab, have_space, have_type, have_layout;
    uint64_t btree, heap;
    int rank;
    uint64_t dims[4];
    int type_class, type_size;
    int layout_class; /* 1 contiguous, 2 chunked */
    uint64_t data_addr, data_size;
    uint64_t chunk_btree;
    uint32_t chunk_dims[5];
    int chunk_rank;
} objinfo;

static int parse_messages(const rfile *r, const unsigned char *p, size_t len, int *left, objinfo *o, int depth);

static int parse_object(const rfile *r, uint64_t addr, objinfo *o)
{
    memset(o, 0, sizeof(*o));
    const unsigned char *h = rptr(r, addr, 16);
    if (!h) return fail(MCRAT_IO_ERR_FORMAT, "object header outside the file");
    if (h[0] != 1) return fail(MCRAT_IO_ERR_FORMAT, "object header version %d is not supported (only version 1)", h[0]);
    int nmsg = (int)get_le(h + 2, 2);
    size_t size = (size_t)get_le(h + 8, 4);
    const unsigned char *p = rptr(r, addr + 16, size);
    if (!p) return fail(MCRAT_IO_ERR_FORMAT, "object header messages outside the file");
    return parse_messages(r, p, size, &nmsg, o, 0);
}

static int parse_messages(const rfile *r, const unsigned char *p, size_t len, int *left, objinfo *o, int depth)
{
    size_t at = 0;
    while (at + 8 <= len && *left > 0) {
        const int type = (int)get_le(p + at, 2);
        const size_t sz = (size_t)get_le(p + at + 2, 2);
        const unsigned char *d = p + at + 8;
        if (at + 8 + sz > len) return fail(MCRAT_IO_ERR_FORMAT, "object header message overruns its block");
        (*left)--;
        if (type == 0x0011 && sz >= 16) {
            o->have_symtab = 1;
            o->btree = get_le(d, 8);
            o->heap = get_le(d + 8, 8);
        } else if (type == 0x0001 && sz >= 8) {
            const int ver = d[0];
            o->rank = d[1];
            if (o->rank > 4) return fail(MCRAT_IO_ERR_FORMAT, "dataspace rank %d not supported", o->rank);
            const size_t hdr = (ver == 1) ? 8 : 4;
            if (sz < hdr + 8 * (size_t)o->rank) return fail(MCRAT_IO_ERR_FORMAT, "short dataspace message");
            for (int k = 0; k < o->rank; ++k) o->dims[k] = get_le(d + hdr + 8 * (size_t)k, 8);
            o->have_space = 1;
        } else if (type == 0x0003 && sz >= 8) {
            o->type_class = d[0] & 0x0f;
            o->type_size = (int)get_le(d + 4, 4);
            if ((d[1] & 1) != 0) return fail(MCRAT_IO_ERR_FORMAT, "big-endian datatypes are not supported");
            o->have_type = 1;
        } else if (type == 0x0008 && sz >= 2) {
            const int ver = d[0];
            if (ver == 3) {
                o->layout_class = d[1];
                if (o->layout_class == 1 && sz >= 18) {
                    o->data_addr = get_le(d + 2, 8);
                    o->data_size = get_le(d + 10, 8);
                } else if (o->layout_class == 2 && sz >= 3 + 8) {
                    o->chunk_rank = d[2];
                    if (o->chunk_rank > 5) return fail(MCRAT_IO_ERR_FORMAT, "chunk rank %d not supported", o->chunk_rank);
                    o->chunk_btree = get_le(d + 3, 8);
                    for (int k = 0; k < o->chunk_rank; ++k) o->chunk_dims[k] = (uint32_t)get_le(d + 11 + 4 * (size_t)k, 4);
                } else {
                    return fail(MCRAT_IO_ERR_FORMAT, "data layout class %d is not supported", o->layout_class);
                }
            } else if (ver == 1 || ver == 2) {
                /* version 1/2: rank (1), class (1), reserved (5), [address (8) unless compact], dims (4 each) */
                const int rank = d[1];
                o->layout_class = d[2];
                if (o->layout_class == 1) {
                    o->data_addr = get_le(d + 8, 8);
                    o->data_size = 0; /* from dataspace x datatype */
                } else if (o->layout_class == 2) {
                    o->chunk_btree = get_le(d + 8, 8);
                    o->chunk_rank = rank;
                    if (rank > 5) return fail(MCRAT_IO_ERR_FORMAT, "chunk rank %d not supported", rank);
                    for (int k = 0; k < rank; ++k) o->chunk_dims[k] = (uint32_t)get_le(d + 16 + 4 * (size_t)k, 4);
                } else {
                    return fail(MCRAT_IO_ERR_FORMAT, "data layout class %d is not supported", o->layout_class);
                }
            } else {
                return fail(MCRAT_IO_ERR_FORMAT, "data layout message version %d is not supported", ver);
            }
            o->have_layout = 1;
        } else if (type == 0x000b) {
            return fail(MCRAT_IO_ERR_FORMAT, "filtered (compressed) datasets are not supported");
        } else if (type == 0x0010 && sz >= 16 && depth < 8) {
            /* object header continuation */
            const uint64_t caddr = get_le(d, 8);
            const size_t clen = (size_t)get_le(d + 8, 8);
            const unsigned char *c = rptr(r, caddr, clen);
            if (!c) return fail(MCRAT_IO_ERR_FORMAT, "object header continuation outside the file");
            int rc = parse_messages(r, c, clen, left, o, depth + 1);
            if (rc) return rc;
        }
        at += 8 + sz;
    }
    return MCRAT_IO_OK;
}

typedef int (*link_cb)(void *ctx, const char *name, uint64_t ohdr);

static int walk_group_btree(const rfile *r, uint64_t node, const unsigned char *heap_data, size_t heap_len, link_cb cb, void *ctx, int depth)
{
    if (depth > 16) return fail(MCRAT_IO_ERR_FORMAT, "group B-tree too deep");
    const unsigned char *t = rptr(r, node, 24);
    if (!t || memcmp(t, "TREE", 4) != 0) return fail(MCRAT_IO_ERR_FORMAT, "bad group B-tree node");
    if (t[4] != 0) return fail(MCRAT_IO_ERR_FORMAT, "B-tree node type %d where a group node was expected", t[4]);
    const int level = t[5], used = (int)get_le(t + 6, 2);
    const unsigned char *kc = rptr(r, node + 24, (size_t)(2 * used + 1) * 8);
    if (!kc) return fail(MCRAT_IO_ERR_FORMAT, "group B-tree node outside the file");
    for (int i = 0; i < used; ++i) {
        const uint64_t child = get_le(kc + 8 + 16 * (size_t)i, 8);
        if (level > 0) {
            int rc = walk_group_btree(r, child, heap_data, heap_len, cb, ctx, depth + 1);
            if (rc) return rc;
            continue;
        }
        const unsigned char *s = rptr(r, child, 8);
        if (!s || memcmp(s, "SNOD", 4) != 0) return fail(MCRAT_IO_ERR_FORMAT, "bad symbol table node");
        const int nsym = (int)get_le(s + 6, 2);
        const unsigned char *e = rptr(r, child + 8, (size_t)nsym * 40);
        if (!e) return fail(MCRAT_IO_ERR_FORMAT, "symbol table node outside the file");
        for (int k = 0; k < nsym; ++k) {
            const uint64_t noff = get_le(e + 40 * (size_t)k, 8);
            if (noff >= heap_len) return fail(MCRAT_IO_ERR_FORMAT, "link name outside the local heap");
            const char *nm = (const char *)heap_data + noff;
            if (!memchr(nm, 0, heap_len - noff)) return fail(MCRAT_IO_ERR_FORMAT, "unterminated link name");
            int rc = cb(ctx, nm, get_le(e + 40 * (size_t)k + 8, 8));
            if (rc) return rc;
        }
    }
    return MCRAT_IO_OK;
}

static int walk_group(const rfile *r, const objinfo *g, link_cb cb, void *ctx)
{
    const unsigned char *h = rptr(r, g->heap, 32);
    if (!h || memcmp(h, "HEAP", 4) != 0) return fail(MCRAT_IO_ERR_FORMAT, "bad local heap");
    const size_t dlen = (size_t)get_le(h + 8, 8);
    const unsigned char *hd = rptr(r, get_le(h + 24, 8), dlen);
    if (!hd) return fail(MCRAT_IO_ERR_FORMAT, "local heap data outside the file");
    return walk_group_btree(r, g->btree, hd, dlen, cb, ctx, 0);
}

static int rfile_open(const char *path, rfile *r)
{
    memset(r, 0, sizeof(*r));
    FILE *fp = fopen(path, "rb");
    if (!fp) return fail(MCRAT_IO_ERR_OPEN, "cannot open %s: %s", path, strerror(errno));
    fseek(fp, 0, SEEK_END);
    long n = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    if (n < 96) {
        fclose(fp);
        return fail(MCRAT_IO_ERR_FORMAT, "%s is too short to be an HDF5 file", path);
    }
    r->b = (unsigned char *)malloc((size_t)n);
    if (!r->b) {
        fclose(fp);
        return MCRAT_IO_ERR_NOMEM;
    }
    r->n = (size_t)n;
    if (fread(r->b, 1, r->n, fp) != r->n) {
        fclose(fp);
        free(r->b);
        return fail(MCRAT_IO_ERR_OPEN, "short read on %s", path);
    }
    fclose(fp);
    static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    size_t at = 0;
    int found = 0;
    while (at + 96 <= r->n) { /* the superblock sits at 0, 512, 1024, 2048, ... */
        if (memcmp(r->b + at, sig, 8) == 0) {
            found = 1;
            break;
        }
        at = at ? at * 2 : 512;
    }
    if (!found) {
        free(r->b);
        return fail(MCRAT_IO_ERR_FORMAT, "%s: no HDF5 signature", path);
    }
    const unsigned char *sb = r->b + at;
    if (sb[8] != 0 && sb[8] != 1) {
        const int ver = sb[8];
        free(r->b);
        return fail(MCRAT_IO_ERR_FORMAT, "%s: superblock version %d is not supported (only 0 and 1)", path, ver);
    }
    if (sb[13] != 8 || sb[14] != 8) {
        free(r->b);
        return fail(MCRAT_IO_ERR_FORMAT, "%s: only 8-byte offsets and lengths are supported", path);
    }
    r->leafK = (int)get_le(sb + 16, 2);
    r->intK = (int)get_le(sb + 18, 2);
    const size_t v1 = (sb[8] == 1) ? 4 : 0; /* version 1 adds indexed-storage K + reserved */
    r->base = get_le(sb + 24 + v1, 8); /* addresses in the file are relative to the base address */
    return (int)at; /* >= 0: offset of the superblock */
}

static uint64_t rfile_root(const rfile *r, size_t sb_at)
{
    const unsigned char *sb = r->b + sb_at;
    const size_t v1 = (sb[8] == 1) ? 4 : 0;
    return get_le(sb + 56 + v1 + 8, 8);
}

typedef struct {
    const char *want;
    uint64_t ohdr;
    int found;
} find_ctx;

static int find_cb(void *ctx, const char *name, uint64_t ohdr)
{
    find_ctx *f = (find_ctx *)ctx;
    if (strcmp(name, f->want) == 0) {
        f->ohdr = ohdr;
        f->found = 1;
    }
    return 0;
}

/* resolves "a/b" from the root; fills the object info of the target */
static int resolve(const rfile *r, size_t sb_at, const char *name, objinfo *o)
{
    uint64_t addr = rfile_root(r, sb_at);
    int rc = parse_object(r, addr, o);
    if (rc) return rc;
    char tmp[256];
    snprintf(tmp, sizeof(tmp), "%s", name ? name : "");
    char *save = NULL;
    for (char *tok = strtok_r(tmp, "/", &save); tok; tok = strtok_r(NULL, "/", &save)) {
        if (!o->have_symtab) return fail(MCRAT_IO_ERR_NOTFOUND, "'%s': not a group on the way to '%s'", tok, name);
        find_ctx fc = {tok, 0, 0};
        rc = walk_group(r, o, find_cb, &fc);
        if (rc) return rc;
        if (!fc.found) return fail(MCRAT_IO_ERR_NOTFOUND, "no object '%s' (looking for '%s')", tok, name);
        rc = parse_object(r, fc.ohdr, o);
        if (rc) return rc;
    }
    return MCRAT_IO_OK;
}

static int dataset_elems(const objinfo *o, size_t *n)
{
    if (!o->have_space || !o->have_type || !o->have_layout) return fail(MCRAT_IO_ERR_FORMAT, "not a dataset (missing dataspace / datatype / layout)");
    size_t tot = 1;
    for (int k = 0; k < o->rank; ++k) tot *= (size_t)o->dims[k];
    if (o->rank == 0) tot = 1;
    *n = tot;
    return MCRAT_IO_OK;
}

/* un-filtered chunked 1-D storage (what the reference's per-rank files use): B-tree v1, node type 1 */
static int read_chunks(const rfile *r, const objinfo *o, uint64_t node, unsigned char *out, size_t total_bytes, int es, int depth)
{
    if (depth > 16) return fail(MCRAT_IO_ERR_FORMAT, "chunk B-tree too deep");
    const unsigned char *t = rptr(r, node, 24);
    if (!t || memcmp(t, "TREE", 4) != 0 || t[4] != 1) return fail(MCRAT_IO_ERR_FORMAT, "bad chunk B-tree node");
    const int level = t[5], used = (int)get_le(t + 6, 2);
    const size_t keysz = 8 + 8 * (size_t)o->chunk_rank; /* chunk size, filter mask, offsets (rank incl. the element dim) */
    const unsigned char *kc = rptr(r, node + 24, (size_t)used * (keysz + 8) + keysz);
    if (!kc) return fail(MCRAT_IO_ERR_FORMAT, "chunk B-tree node outside the file");
    for (int i = 0; i < used; ++i) {
        const unsigned char *key = kc + (size_t)i * (keysz + 8);
        const uint64_t child = get_le(key + keysz, 8);
        if (level > 0) {
            int rc = read_chunks(r, o, child, out, total_bytes, es, depth + 1);
            if (rc) return rc;
            continue;
        }
        const size_t csize = (size_t)get_le(key, 4);
        if (get_le(key + 4, 4) != 0) return fail(MCRAT_IO_ERR_FORMAT, "filtered chunks are not supported");
        const uint64_t off0 = get_le(key + 8, 8); /* element offset along dimension 0 */
        const unsigned char *src = rptr(r, child, csize);
        if (!src) return fail(MCRAT_IO_ERR_FORMAT, "chunk outside the file");
        size_t dst = (size_t)off0 * (size_t)es;
        if (dst >= total_bytes) continue;
        size_t nbytes = csize;
        if (nbytes > total_bytes - dst) nbytes = total_bytes - dst; /* the last chunk may extend past the dataset */
        memcpy(out + dst, src, nbytes);
    }
    return MCRAT_IO_OK;
}

static int read_dataset_raw(const rfile *r, const objinfo *o, void *out, size_t n, int want_i8)
{
    size_t tot = 0;
    int rc = dataset_elems(o, &tot);
    if (rc) return rc;
    if (tot != n) return fail(MCRAT_IO_ERR_ARG, "dataset has %zu elements, caller asked for %zu", tot, n);
    const int is_i8 = (o->type_class == 0 && o->type_size == 1);
    const int is_f64 = (o->type_class == 1 && o->type_size == 8);
    if (!is_i8 && !is_f64) return fail(MCRAT_IO_ERR_FORMAT, "datatype class %d size %d is not supported", o->type_class, o->type_size);
    if (is_i8 != want_i8) return fail(MCRAT_IO_ERR_ARG, "dataset element type does not match the output buffer");
    const int es = is_i8 ? 1 : 8;
    if (n == 0) return MCRAT_IO_OK;
    if (o->layout_class == 1) {
        const unsigned char *src = rptr(r, o->data_addr, n * (size_t)es);
        if (!src) return fail(MCRAT_IO_ERR_FORMAT, "dataset raw data outside the file");
        memcpy(out, src, n * (size_t)es);
        return MCRAT_IO_OK;
    }
    if (o->chunk_rank != 2) return fail(MCRAT_IO_ERR_FORMAT, "only 1-D chunked datasets are supported");
    memset(out, 0, n * (size_t)es);
    if (o->chunk_btree == H5_UNDEF) return MCRAT_IO_OK;
    return read_chunks(r, o, o->chunk_btree, (unsigned char *)out, n * (size_t)es, es, 0);
}

API long long mcrat_b200_h5_dataset_length(const char *path, const char *name)
{
    rfile r;
    int at = rfile_open(path, &r);
    if (at < 0) return at;
    objinfo o;
    int rc = resolve(&r, (size_t)at, name, &o);
    size_t n = 0;
    if (rc == MCRAT_IO_OK) rc = dataset_elems(&o, &n);
    free(r.b);
    return rc ? rc : (long long)n;
}

API int mcrat_b200_h5_read_dataset(const char *path, const char *name, double *out_f64, signed char *out_i8, size_t n)
{
    if (!out_f64 && !out_i8) return fail(MCRAT_IO_ERR_ARG, "h5_read_dataset: no output buffer");
    rfile r;
    int at = rfile_open(path, &r);
    if (at < 0) return at;
    objinfo o;
    int rc = resolve(&r, (size_t)at, name, &o);
    if (rc == MCRAT_IO_OK) {
        const int is_i8 = (o.type_class == 0 && o.type_size == 1);
        if (is_i8 && !out_i8) rc = fail(MCRAT_IO_ERR_ARG, "dataset %s holds 8-bit integers", name);
        else if (!is_i8 && !out_f64) rc = fail(MCRAT_IO_ERR_ARG, "dataset %s holds doubles", name);
        else rc = read_dataset_raw(&r, &o, is_i8 ? (void *)out_i8 : (void *)out_f64, n, is_i8);
    }
    free(r.b);
    return rc;
}

typedef struct {
    char *buf;
    size_t len, cap;
    int count;
} list_ctx;

static int list_cb(void *ctx, const char *name, uint64_t ohdr)
{
    (void)ohdr;
    list_ctx *l = (list_ctx *)ctx;
    size_t n = strlen(name);
    if (l->buf && l->len + n + 2 <= l->cap) {
        memcpy(l->buf + l->len, name, n);
        l->buf[l->len + n] = '\n';
        l->buf[l->len + n + 1] = 0;
        l->len += n + 1;
    }
    l->count++;
    return 0;
}

API int mcrat_b200_h5_list(const char *path, const char *group, char *buf, size_t buflen)
{
    rfile r;
    int at = rfile_open(path, &r);
    if (at < 0) return at;
    objinfo o;
    int rc = resolve(&r, (size_t)at, group, &o);
    list_ctx l = {buf, 0, buflen, 0};
    if (buf && buflen) buf[0] = 0;
    if (rc == MCRAT_IO_OK) {
        if (!o.have_symtab) rc = fail(MCRAT_IO_ERR_FORMAT, "'%s' is not a group", group ? group : "/");
        else rc = walk_group(&r, &o, list_cb, &l);
    }
    free(r.b);
    return rc ? rc : l.count;
}

/* loads a whole file (our subset) into the in-memory model: used to append to mc_proc files */
typedef struct {
    const rfile *r;
    h5_file *f;
    h5_group *g;
    int rc;
} load_ctx;

static int load_dset_cb(void *ctx, const char *name, uint64_t ohdr);

static int load_root_cb(void *ctx, const char *name, uint64_t ohdr)
{
    load_ctx *L = (load_ctx *)ctx;
    objinfo o;
    int rc = parse_object(L->r, ohdr, &o);
    if (rc) return rc;
    if (o.have_symtab) {
        if (!h5_group_get(L->f, name, 1)) return MCRAT_IO_ERR_NOMEM;
        load_ctx dctx = *L;
        dctx.g = h5_group_get(L->f, name, 0); /* stays valid: no group is added while its datasets load */
        return walk_group(L->r, &o, load_dset_cb, &dctx);
    }
    load_ctx dctx = *L;
    dctx.g = h5_group_get(L->f, "", 0);
    return load_dset_cb(&dctx, name, ohdr);
}

static int load_dset_cb(void *ctx, const char *name, uint64_t ohdr)
{
    load_ctx *L = (load_ctx *)ctx;
    objinfo o;
    int rc = parse_object(L->r, ohdr, &o);
    if (rc) return rc;
    if (o.have_symtab) return MCRAT_IO_OK; /* nested groups below the first level are not part of the layout */
    size_t n = 0;
    rc = dataset_elems(&o, &n);
    if (rc) return rc;
    const int is_i8 = (o.type_class == 0 && o.type_size == 1);
    void *tmp = malloc(n * (is_i8 ? 1 : 8) + 8);
    if (!tmp) return MCRAT_IO_ERR_NOMEM;
    rc = read_dataset_raw(L->r, &o, tmp, n, is_i8);
    if (rc == MCRAT_IO_OK) rc = h5_dset_append(L->g, name, is_i8, tmp, n);
    if (rc == MCRAT_IO_OK && o.layout_class == 2) { /* an extendible dataset stays one, with the chunk size it was created with */
        h5_dset *d = h5_dset_get(L->g, name, 0);
        if (d) d->chunk = o.chunk_dims[0] ? o.chunk_dims[0] : 1;
    }
    free(tmp);
    return rc;
}

static int h5_load(const char *path, h5_file *f)
{
    rfile r;
    int at = rfile_open(path, &r);
    if (at < 0) return at;
    memset(f, 0, sizeof(*f));
    if (!h5_group_get(f, "", 1)) {
        free(r.b);
        return MCRAT_IO_ERR_NOMEM;
    }
    objinfo root;
    int rc = parse_object(&r, rfile_root(&r, (size_t)at), &root);
    if (rc == MCRAT_IO_OK && !root.have_symtab) rc = fail(MCRAT_IO_ERR_FORMAT, "%s: root object is not a symbol-table group", path);
    if (rc == MCRAT_IO_OK) {
        load_ctx L = {&r, f, NULL, 0};
        rc = walk_group(&r, &root, load_root_cb, &L);
    }
    free(r.b);
    if (rc) h5_free(f);
    return rc;
}

/* ============================================================================================
 * photon output
 * ============================================================================================ */
static const char *F64_NAMES_ALL[] = {"P0", "P1", "P2", "P3", "COMV_P0", "COMV_P1", "COMV_P2", "COMV_P3",
                                      "R0", "R1", "R2", "S0", "S1", "S2", "S3", "NS", "PW"};

static double photon_field(const mcrat_photon *p, int k)
{
    switch (k) {
    case 0: return p->p0;
    case 1: return p->p1;
    case 2: return p->p2;
    case 3: return p->p3;
    case 4: return p->comv_p0;
    case 5: return p->comv_p1;
    case 6: return p->comv_p2;
    case 7: return p->comv_p3;
    case 8: return p->r0;
    case 9: return p->r1;
    case 10: return p->r2;
    case 11: return p->s0;
    case 12: return p->s1;
    case 13: return p->s2;
    case 14: return p->s3;
    case 15: return p->num_scatt;
    default: return p->weight;
    }
}

static int field_enabled(int k, const mcrat_b200_io_switches *sw)
{
    if (k >= 4 && k <= 7) return sw->comv_switch != 0;
    if (k >= 11 && k <= 14) return sw->stokes_switch != 0;
    return 1;
}

static int file_exists(const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) return 0;
    fclose(f);
    return 1;
}

API int mcrat_b200_print_photons(const char *dir, int angle_rank, int frame, const mcrat_photon *photons, int list_capacity,
                                 const mcrat_b200_io_switches *sw)
{
    if (!dir || !sw || list_capacity < 0 || (list_capacity > 0 && !photons)) return fail(MCRAT_IO_ERR_ARG, "print_photons: bad argument");
    char path[1024], gname[64];
    snprintf(path, sizeof(path), "%s%smc_proc_%d.h5", dir, (dir[0] && dir[strlen(dir) - 1] != '/') ? "/" : "", angle_rank);
    snprintf(gname, sizeof(gname), "%d", frame);
    h5_file f;
    memset(&f, 0, sizeof(f));
    int rc = MCRAT_IO_OK;
    if (file_exists(path)) {
        rc = h5_load(path, &f);
        if (rc) return rc;
    } else if (!h5_group_get(&f, "", 1)) {
        return MCRAT_IO_ERR_NOMEM;
    }
    h5_group *g = h5_group_get(&f, gname, 1);
    if (!g) {
        h5_free(&f);
        return MCRAT_IO_ERR_NOMEM;
    }
    /* photons with weight != 0, in slot order (Src/mcrat_io.c:150-193) */
    size_t n = 0;
    for (int i = 0; i < list_capacity; ++i)
        if (photons[i].weight != 0) n++;
    double *col = (double *)malloc((n ? n : 1) * sizeof(double));
    signed char *types = (signed char *)malloc(n ? n : 1);
    if (!col || !types) rc = MCRAT_IO_ERR_NOMEM;
    for (int k = 0; k < 17 && rc == MCRAT_IO_OK; ++k) {
        if (!field_enabled(k, sw)) continue;
        size_t c = 0;
        for (int i = 0; i < list_capacity; ++i)
            if (photons[i].weight != 0) col[c++] = photon_field(&photons[i], k);
        const int is_new = h5_dset_get(g, F64_NAMES_ALL[k], 0) == NULL;
        rc = h5_dset_append(g, F64_NAMES_ALL[k], 0, col, n);
        /* H5Pset_chunk(prop, rank, dims) with dims = the number of photons of the first write, Src/mcrat_io.c:254 */
        if (rc == MCRAT_IO_OK && is_new) h5_dset_get(g, F64_NAMES_ALL[k], 0)->chunk = n ? n : 1;
    }
    if (rc == MCRAT_IO_OK && sw->save_type) {
        size_t c = 0;
        for (int i = 0; i < list_capacity; ++i)
            if (photons[i].weight != 0) types[c++] = (signed char)photons[i].type;
        const int is_new = h5_dset_get(g, "PT", 0) == NULL;
        rc = h5_dset_append(g, "PT", 1, types, n);
        if (rc == MCRAT_IO_OK && is_new) h5_dset_get(g, "PT", 0)->chunk = n ? n : 1;
    }
    free(col);
    free(types);
    if (rc == MCRAT_IO_OK) rc = h5_write(path, &f);
    h5_free(&f);
    return rc;
}

API int mcrat_b200_merge_frame(const char *dir, int frame, const int *ranks, int nranks, const mcrat_b200_io_switches *sw)
{
    if (!dir || !sw || !ranks || nranks < 1) return fail(MCRAT_IO_ERR_ARG, "merge_frame: bad argument");
    const char *sep = (dir[0] && dir[strlen(dir) - 1] != '/') ? "/" : "";
    h5_file out;
    memset(&out, 0, sizeof(out));
    h5_group *root = h5_group_get(&out, "", 1);
    if (!root) return MCRAT_IO_ERR_NOMEM;
    int rc = MCRAT_IO_OK, merged = 0;
    /* create every dataset up front so that an all-empty frame still has the full set of names */
    for (int k = 0; k < 17 && rc == MCRAT_IO_OK; ++k)
        if (field_enabled(k, sw)) rc = h5_dset_append(root, F64_NAMES_ALL[k], 0, NULL, 0);
    if (rc == MCRAT_IO_OK && sw->save_type) rc = h5_dset_append(root, "PT", 1, NULL, 0);
    for (int q = 0; q < nranks && rc == MCRAT_IO_OK; ++q) {
        char path[1024], name[128];
        snprintf(path, sizeof(path), "%s%smc_proc_%d.h5", dir, sep, ranks[q]);
        if (!file_exists(path)) continue;
        snprintf(name, sizeof(name), "%d/P0", frame);
        long long n = mcrat_b200_h5_dataset_length(path, name);
        if (n == MCRAT_IO_ERR_NOTFOUND) continue; /* this rank did not write the frame */
        if (n < 0) {
            rc = (int)n;
            break;
        }
        double *col = (double *)malloc(((size_t)n ? (size_t)n : 1) * sizeof(double));
        signed char *types = (signed char *)malloc((size_t)n ? (size_t)n : 1);
        if (!col || !types) rc = MCRAT_IO_ERR_NOMEM;
        for (int k = 0; k < 17 && rc == MCRAT_IO_OK; ++k) {
            if (!field_enabled(k, sw)) continue;
            snprintf(name, sizeof(name), "%d/%s", frame, F64_NAMES_ALL[k]);
            rc = mcrat_b200_h5_read_dataset(path, name, col, NULL, (size_t)n);
            if (rc == MCRAT_IO_OK) rc = h5_dset_append(root, F64_NAMES_ALL[k], 0, col, (size_t)n);
            root = h5_group_get(&out, "", 0);
        }
        if (rc == MCRAT_IO_OK && sw->save_type) {
            snprintf(name, sizeof(name), "%d/PT", frame);
            rc = mcrat_b200_h5_read_dataset(path, name, NULL, types, (size_t)n);
            if (rc == MCRAT_IO_OK) rc = h5_dset_append(root, "PT", 1, types, (size_t)n);
        }
        free(col);
        free(types);
        merged++;
    }
    if (rc == MCRAT_IO_OK) {
        char path[1024];
        snprintf(path, sizeof(path), "%s%smcdata_%d.h5", dir, sep, frame);
        rc = h5_write(path, &out);
    }
    h5_free(&out);
    (void)merged;
    return rc;
}
