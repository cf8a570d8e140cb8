// cs_emit.cuh -- K6: photonEmitCyclosynch in its all-cells mode (inject_single_switch == 0), Src/mc_cyclosynch.c:1176-1464.
// Part of the single translation unit mcrat_b200.cu (included there, after aux_kernels.cuh); not a stand-alone header.
//
// The reference, once per hydro frame (Src/mcrat.c:747):
//   1. selects the cells of the emission shell: rmin <= r(outer corner), r(inner corner) < rmax, theta(outer) >= theta_min,
//      theta(inner) < theta_max (:1211-1229);
//   2. searches the photon weight: per pass and selected cell, the number of seed photons is a Poisson draw with mean
//      (integral of the black-body photon spectrum from 10 Hz to the cell's cyclotron frequency, gsl_integration_qags at
//      epsrel 1e-2) x cell volume / weight; the weight is multiplied by 10 (too many photons) or 0.5 (none) until the
//      total lies in [1, CYCLOSYNCHROTRON_REBIN_E_PERC x max_photons] (:1248-1318);
//   3. creates the photons cell by cell at the cell centre (random azimuth in 2-D), isotropic in the fluid frame at the
//      cyclotron frequency, boosted to the lab frame, type 'p' (:1353-1460);
//   4. puts photon k into the k-th null slot of the list (addToPhotonList, Src/photons.c:167-205).
// On the device: (1) and the integral / volume of every selected cell are one thread per cell, compacted in cell order;
// (2) is one block, a thread per selected cell and pass, the Poisson draws from a Philox stream keyed by (cell rank,
// emission call, pass); (3) is one thread per photon, its draws keyed by (photon index, emission call); (4) reuses the
// null-slot ranking of the rebin kernels.  The integral does not depend on the weight, so it is evaluated once per cell
// instead of once per pass (same value).  With the replay harness (a recorded gsl_rng stream that must be consumed in the
// reference's order) steps (2) and (3) run on one thread, sequentially, like the reference.
#pragma once

// hydroCoordinateToSpherical, Src/geometry.c:66-106
__device__ inline void hydro_coord_to_spherical(int dims, int g, double &r, double &theta, double r0, double r1, double r2)
{
    double sph_r = 0, sph_theta = 0;
    if (dims != D_THREE) {
        if (g == G_CARTESIAN || g == G_CYLINDRICAL) {
            sph_r = sqrt(r0 * r0 + r1 * r1);
            sph_theta = atan2(r0, r1);
        }
        if (g == G_SPHERICAL) {
            sph_r = r0;
            sph_theta = r1;
        }
    } else {
        if (g == G_CARTESIAN) {
            sph_r = sqrt(r0 * r0 + r1 * r1 + r2 * r2);
            sph_theta = acos(r2 / sph_r);
        }
        if (g == G_SPHERICAL) {
            sph_r = r0;
            sph_theta = r1;
        }
        if (g == G_POLAR) {
            sph_r = sqrt(r0 * r0 + r2 * r2);
            sph_theta = acos(r2 / sph_r);
        }
    }
    r = sph_r;
    theta = sph_theta;
}

// hydroElementVolume, Src/geometry.c:255-296
__device__ inline double hydro_element_volume(int dims, int g, double c0, double c1, double c2, double s0, double s1, double s2)
{
    double V = 0;
    const double r0_max = c0 + 0.5 * s0, r0_min = c0 - 0.5 * s0, r1_max = c1 + 0.5 * s1, r1_min = c1 - 0.5 * s1;
    if (dims != D_THREE) {
        if (g == G_CARTESIAN || g == G_CYLINDRICAL) V = PI * (r0_max * r0_max - r0_min * r0_min) * s1;
        if (g == G_SPHERICAL) V = (2.0 * PI / 3.0) * (r0_max * r0_max * r0_max - r0_min * r0_min * r0_min) * (cos(r1_min) - cos(r1_max));
    } else {
        const double r2_max = c2 + 0.5 * s2, r2_min = c2 - 0.5 * s2;
        if (g == G_CARTESIAN) V = s0 * s1 * s2;
        if (g == G_SPHERICAL)
            V = (1.0 / 3.0) * (r0_max * r0_max * r0_max - r0_min * r0_min * r0_min) * (cos(r1_min) - cos(r1_max)) * (r2_max - r2_min);
        if (g == G_POLAR) V = 0.5 * (r0_max * r0_max - r0_min * r0_min) * s1 * s2;
    }
    return V;
}

// blackbody_ph_spect, Src/mc_cyclosynch.c:185-195
__device__ inline double blackbody_ph_spect(double nu, double temp)
{
    return (8 * PI * nu * nu) / (exp(PL_CONST * nu / (K_B * temp)) - 1) / (C_LIGHT * C_LIGHT * C_LIGHT);
}

// (7, 15) Gauss-Kronrod rule and its error estimate |K15 - G7| (the rule GSL's qags applies to every interval)
__device__ inline void gk15_blackbody(double temp, double a, double b, double &res, double &err)
{
    const double xgk[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                           0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                           0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                           0.207784955007898467600689403773245, 0.000000000000000000000000000000000};
    const double wgk[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                           0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                           0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                           0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
    const double wg[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                          0.381830050505118944950369775488975, 0.417959183673469387755102040816327};
    const double c = 0.5 * (a + b), h = 0.5 * (b - a);
    const double fc = blackbody_ph_spect(c, temp);
    double rk = fc * wgk[7], rg = fc * wg[3];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const double dx = h * xgk[j];
        const double f1 = blackbody_ph_spect(c - dx, temp), f2 = blackbody_ph_spect(c + dx, temp);
        rk += wgk[j] * (f1 + f2);
        if (j & 1) rg += wg[j / 2] * (f1 + f2);
    }
    res = rk * h;
    err = fabs((rk - rg) * h);
}

// Adaptive quadrature with global bisection of the worst interval, to epsrel (the call site of the reference:
// gsl_integration_qags(&F, 10, nu_c, 0, 1e-2, 10000, ...), Src/mc_cyclosynch.c:1285).  The integrand is the
// Rayleigh-Jeans tail of a black body -- almost a parabola -- so the first rule already meets 1e-2 by many orders
// of magnitude; CS_QUAD_SEGS intervals are kept for the rest.  Returns false if they do not suffice.
constexpr int CS_QUAD_SEGS = 48;
__device__ inline bool integrate_blackbody_tail(double temp, double a, double b, double epsrel, double &result)
{
    double sa[CS_QUAD_SEGS], sb[CS_QUAD_SEGS], sr[CS_QUAD_SEGS], se[CS_QUAD_SEGS];
    int n = 1;
    sa[0] = a;
    sb[0] = b;
    gk15_blackbody(temp, a, b, sr[0], se[0]);
    for (;;) {
        double tot = 0, err = 0;
        int worst = 0;
        for (int i = 0; i < n; ++i) {
            tot += sr[i];
            err += se[i];
            if (se[i] > se[worst]) worst = i;
        }
        const double tol = fmax(0.0, epsrel * fabs(tot));
        if (err <= tol || n >= CS_QUAD_SEGS) {
            result = tot;
            return err <= tol;
        }
        const double wa = sa[worst], wb = sb[worst], mid = 0.5 * (wa + wb);
        sb[worst] = mid;
        gk15_blackbody(temp, wa, mid, sr[worst], se[worst]);
        sa[n] = mid;
        sb[n] = wb;
        gk15_blackbody(temp, mid, wb, sr[n], se[n]);
        n++;
    }
}

// uniform source keyed by (stream, slot, call / pass): counter (draw / 2, iter_lo, slot, stream + 8 * iter_hi)
struct KeyedRng {
    uint32_t k0, k1, slot, c1, c3;
    uint64_t draw;
    __device__ KeyedRng(uint32_t k0_, uint32_t k1_, uint32_t stream, uint32_t slot_, uint64_t iter)
        : k0(k0_), k1(k1_), slot(slot_), c1((uint32_t)iter), c3(stream + ((uint32_t)(iter >> 32) << 3)), draw(0) {}
    __device__ double uniform()
    {
        double a, b;
        philox_doubles((uint32_t)(draw >> 1), c1, slot, c3, k0, k1, a, b);
        const double u = (draw & 1ull) ? b : a;
        draw++;
        return u;
    }
    __device__ double uniform_pos() { return uniform(); } // strictly inside (0, 1) already
};

// gsl_ran_poisson as the oracle / the reference's GSL stand-in draw it: Knuth's product of uniforms below a mean of 10,
// Hoermann's transformed rejection (PTRS) above
template <class Rng>
__device__ inline unsigned int ran_poisson(Rng &r, double mu)
{
    if (!(mu > 0)) return 0;
    if (mu < 10.0) {
        const double emu = exp(-mu);
        double prod = 1.0;
        unsigned int k = 0;
        do {
            prod *= r.uniform();
            k++;
        } while (prod > emu);
        return k - 1;
    }
    const double smu = sqrt(mu);
    const double b = 0.931 + 2.53 * smu;
    const double a = -0.059 + 0.02483 * b;
    const double inv_alpha = 1.1239 + 1.1328 / (b - 3.4);
    const double v_r = 0.9277 - 3.6224 / (b - 2.0);
    for (int guard = 0; guard < (1 << 20); ++guard) {
        const double U = r.uniform() - 0.5;
        const double V = r.uniform_pos();
        const double us = 0.5 - fabs(U);
        const double kf = floor((2.0 * a / us + b) * U + mu + 0.43);
        if (us >= 0.07 && V <= v_r) return (unsigned int)kf;
        if (kf < 0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(inv_alpha) - log(a / (us * us) + b) <= -mu + kf * log(mu) - lgamma(kf + 1.0)) return (unsigned int)kf;
    }
    return (unsigned int)mu;
}

// The reference stores gsl_ran_poisson's unsigned result in an int and sums ints (Src/mc_cyclosynch.c:1290-1295): with a
// suggested weight so small that a cell would emit more than 2^31 photons that overflows (undefined behaviour).  Here such
// a count saturates and the sum is 64 bits wide, so the search simply sees "too many photons" and raises the weight.
__device__ inline int poisson_count(unsigned int k) { return k > (unsigned int)INT_MAX ? INT_MAX : (int)k; }

struct CsEmitWork {
    int n_cells;
    int *flag;         // [n_cells] 1 = cell lies in the emission shell
    int *block_base;   // [ceil(n_cells / 256)] exclusive prefix of the selected cells per 256-cell block
    int *sel_cell;     // [n_sel] selected cells in array order
    double *integ, *vol, *nu_c; // [n_sel]
    int *count;        // [n_sel] Poisson count of the accepted pass
    long long *offset; // [n_sel + 1] exclusive prefix of count
    int *meta;         // [0] n_sel, [1] ph_tot, [2] weight-search passes, [3] status (0 ok, 1 quadrature, 2 pass limit)
    double *weight;    // [0] ph_weight_adjusted
};

__device__ inline void cs_cell_corners(const DevCtx &d, int i, double &ri, double &ti, double &ro, double &to)
{
    const CellCols &c = d.cells;
    if (d.dims == D_THREE) {
        hydro_coord_to_spherical(d.dims, d.geom, ri, ti, fabs(c.r0[i]) - 0.5 * c.s0[i], fabs(c.r1[i]) - 0.5 * c.s1[i],
                                 fabs(c.r2[i]) - 0.5 * c.s2[i]);
        hydro_coord_to_spherical(d.dims, d.geom, ro, to, fabs(c.r0[i]) + 0.5 * c.s0[i], fabs(c.r1[i]) + 0.5 * c.s1[i],
                                 fabs(c.r2[i]) + 0.5 * c.s2[i]);
    } else {
        hydro_coord_to_spherical(d.dims, d.geom, ri, ti, c.r0[i] - 0.5 * c.s0[i], c.r1[i] - 0.5 * c.s1[i], 0);
        hydro_coord_to_spherical(d.dims, d.geom, ro, to, c.r0[i] + 0.5 * c.s0[i], c.r1[i] + 0.5 * c.s1[i], 0);
    }
}

// step 1a: shell test per cell, selected cells counted per 256-cell block
__global__ void __launch_bounds__(256) cs_select_kernel(DevCtx d, CsEmitWork w, double rmin, double rmax, double theta_min,
                                                        double theta_max)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    int sel = 0;
    if (i < w.n_cells) {
        double ri, ti, ro, to;
        cs_cell_corners(d, i, ri, ti, ro, to);
        sel = ((rmin <= ro) && (ri < rmax) && (to >= theta_min) && (ti < theta_max)) ? 1 : 0;
        w.flag[i] = sel;
    }
    const int c = __syncthreads_count(sel);
    if (threadIdx.x == 0) w.block_base[blockIdx.x] = c;
}

// step 1b: exclusive scan of the block counts (one block; a million cells are 4096 counts)
__global__ void __launch_bounds__(1024) cs_block_scan_kernel(CsEmitWork w, int nblocks)
{
    __shared__ int part[1024];
    const int per = (nblocks + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(nblocks, lo + per);
    int s = 0;
    for (int b = lo; b < hi; ++b) s += w.block_base[b];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int t = 0; t < 1024; ++t) {
            const int c = part[t];
            part[t] = run;
            run += c;
        }
        w.meta[0] = run;
    }
    __syncthreads();
    int run = part[threadIdx.x];
    for (int b = lo; b < hi; ++b) {
        const int c = w.block_base[b];
        w.block_base[b] = run;
        run += c;
    }
}

// step 1c: compaction in cell order + the per-cell quantities of the weight search
__global__ void __launch_bounds__(256) cs_compact_kernel(DevCtx d, CsEmitWork w)
{
    __shared__ int warp_off[8];
    const int i = blockIdx.x * 256 + threadIdx.x;
    const bool sel = (i < w.n_cells) && w.flag[i];
    const unsigned ball = __ballot_sync(0xffffffffu, sel);
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    if (lane == 0) warp_off[wp] = __popc(ball);
    __syncthreads();
    if (!sel) return;
    int j = w.block_base[blockIdx.x];
    for (int q = 0; q < wp; ++q) j += warp_off[q];
    j += __popc(ball & ((1u << lane) - 1u));
    const double b_field = cell_b_field(d, i);
    const double nu_c = calc_cyclotron_freq(b_field);
    double integ = 0;
    if (!integrate_blackbody_tail(d.cells.temp[i], 10, nu_c, 1e-2, integ)) w.meta[3] = 1;
    w.sel_cell[j] = i;
    w.integ[j] = integ;
    w.nu_c[j] = nu_c;
    w.vol[j] = hydro_element_volume(d.dims, d.geom, d.cells.r0[i], d.cells.r1[i], d.cells.r2[i], d.cells.s0[i], d.cells.s1[i],
                                    d.cells.s2[i]);
}

// step 2: the weight search (Src/mc_cyclosynch.c:1248-1318) and the exclusive prefix of the accepted counts.  One block.
constexpr int CS_MAX_PASSES = 400;
__global__ void __launch_bounds__(1024) cs_weight_kernel(DevCtx d, CsEmitWork w, double ph_weight, double max_photons, uint32_t epoch)
{
    __shared__ long long red[32];
    __shared__ long long sh_tot;
    __shared__ double sh_weight;
    __shared__ int sh_done;
    const int n_sel = w.meta[0];
    const int min_photons = (n_sel == 0) ? 0 : 1;
    if (threadIdx.x == 0) {
        sh_weight = ph_weight;
        sh_done = 0;
    }
    __syncthreads();
    int pass = 0;
    for (; pass < CS_MAX_PASSES; ++pass) {
        const double weight = sh_weight;
        long long mine = 0;
        if (d.replay) {
            // a recorded gsl_rng stream: consumed by one thread in the reference's order (cells ascending)
            if (threadIdx.x == 0) {
                EventRng rng;
                rng.replay = 1;
                rng.buf = d.replay_buf;
                rng.pos = d.gs->replay_cursor;
                rng.n = d.gs->replay_n;
                rng.exhausted = 0;
                rng.pre = nullptr;
                rng.npre = 0;
                rng.draw = 0;
                for (int j = 0; j < n_sel; ++j) {
                    double mean = w.integ[j];
                    mean *= w.vol[j] / (weight);
                    const int c = poisson_count(ran_poisson(rng, mean));
                    w.count[j] = c;
                    mine += c;
                }
                d.gs->replay_cursor = rng.pos;
                if (rng.exhausted) raise_error(d.gs, MCRAT_B200_ERR_REPLAY, -1, ERR_SITE_CS_EMIT);
            }
        } else {
            for (int j = threadIdx.x; j < n_sel; j += 1024) {
                KeyedRng rng(d.k0, d.k1 ^ d.shard_base, 4u, (uint32_t)j, ((uint64_t)epoch << 32) | (uint64_t)pass);
                double mean = w.integ[j];
                mean *= w.vol[j] / (weight);
                const int c = poisson_count(ran_poisson(rng, mean));
                w.count[j] = c;
                mine += c;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, off);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mine;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long tot = 0;
            for (int k = 0; k < 32; ++k) tot += red[k];
            sh_tot = tot;
            if ((double)tot > max_photons)
                sh_weight = weight * 10;
            else if (tot < min_photons)
                sh_weight = weight * 0.5;
            else
                sh_done = 1;
        }
        __syncthreads();
        if (sh_done) break;
    }
    if (threadIdx.x == 0) {
        w.meta[1] = (int)sh_tot;
        w.meta[2] = pass + 1;
        if (!sh_done) w.meta[3] = 2;
        w.weight[0] = sh_weight;
    }
    // exclusive prefix of the counts (thread t owns a contiguous run of cells)
    __shared__ long long part[1024];
    const int per = (n_sel + 1023) / 1024;
    const int lo = min(n_sel, threadIdx.x * per), hi = min(n_sel, lo + per);
    long long s = 0;
    for (int j = lo; j < hi; ++j) s += w.count[j];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int t = 0; t < 1024; ++t) {
            const long long c = part[t];
            part[t] = run;
            run += c;
        }
        w.offset[n_sel] = run;
    }
    __syncthreads();
    long long run = part[threadIdx.x];
    for (int j = lo; j < hi; ++j) {
        w.offset[j] = run;
        run += w.count[j];
    }
}

// one emitted photon, Src/mc_cyclosynch.c:1371-1448 (the same statements cs_emit_single runs for a replacement photon)
template <class Rng>
__device__ inline void cs_fill_photon(const DevCtx &d, Rng &rng, int cell, double nu_c, double weight, mcrat_photon &e)
{
    const int ndim3 = (d.dims == D_THREE);
    const double fr_dum = nu_c;
    double position_phi = 0;
    if (!ndim3) position_phi = rng.uniform() * 2 * PI;
    const double com_v_phi = rng.uniform() * 2 * PI;
    const double com_v_theta = rng.uniform() * PI;
    double p_comv[4], boost[3], l_boost[4], pos[3];
    p_comv[0] = PL_CONST * fr_dum / C_LIGHT;
    p_comv[1] = (PL_CONST * fr_dum / C_LIGHT) * sin(com_v_theta) * cos(com_v_phi);
    p_comv[2] = (PL_CONST * fr_dum / C_LIGHT) * sin(com_v_theta) * sin(com_v_phi);
    p_comv[3] = (PL_CONST * fr_dum / C_LIGHT) * cos(com_v_theta);
    const double cr0 = d.cells.r0[cell], cr1 = d.cells.r1[cell], cr2 = d.cells.r2[cell];
    if (ndim3)
        hydro_vector_to_cartesian(d.dims, d.geom, boost, d.cells.v0[cell], d.cells.v1[cell], d.cells.v2[cell], cr0, cr1, cr2);
    else if (d.dims == D_TWO_POINT_FIVE)
        hydro_vector_to_cartesian(d.dims, d.geom, boost, d.cells.v0[cell], d.cells.v1[cell], d.cells.v2[cell], cr0, cr1, position_phi);
    else
        hydro_vector_to_cartesian(d.dims, d.geom, boost, d.cells.v0[cell], d.cells.v1[cell], 0, cr0, cr1, position_phi);
    boost[0] *= -1;
    boost[1] *= -1;
    boost[2] *= -1;
    lorentz_boost(boost, p_comv, l_boost, true);
    if (ndim3)
        hydro_coord_to_mcrat(d.dims, d.geom, pos, cr0, cr1, cr2);
    else
        hydro_coord_to_mcrat(d.dims, d.geom, pos, cr0, cr1, position_phi);
    memset(&e, 0, sizeof(e));
    e.type = 'p';
    e.p0 = l_boost[0]; e.p1 = l_boost[1]; e.p2 = l_boost[2]; e.p3 = l_boost[3];
    e.comv_p0 = p_comv[0]; e.comv_p1 = p_comv[1]; e.comv_p2 = p_comv[2]; e.comv_p3 = p_comv[3];
    e.r0 = pos[0]; e.r1 = pos[1]; e.r2 = pos[2];
    e.s0 = 1; e.s1 = 0; e.s2 = 0; e.s3 = 0;
    e.num_scatt = 0;
    e.weight = weight;
    e.nearest_block_index = 0; // Src/mc_cyclosynch.c:1443
    e.recalc_properties = 1;
}

// step 3: one thread per photon (Philox) or one thread for all of them in list order (replay)
__global__ void __launch_bounds__(256) cs_fill_kernel(DevCtx d, CsEmitWork w, mcrat_photon *emitted, uint32_t epoch)
{
    const int n_sel = w.meta[0], ph_tot = w.meta[1];
    const double weight = w.weight[0];
    if (d.replay) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            EventRng rng;
            rng.replay = 1;
            rng.buf = d.replay_buf;
            rng.pos = d.gs->replay_cursor;
            rng.n = d.gs->replay_n;
            rng.exhausted = 0;
            rng.pre = nullptr;
            rng.npre = 0;
            rng.draw = 0;
            int k = 0;
            for (int j = 0; j < n_sel && k < ph_tot; ++j)
                for (int q = 0; q < w.count[j]; ++q) cs_fill_photon(d, rng, w.sel_cell[j], w.nu_c[j], weight, emitted[k++]);
            d.gs->replay_cursor = rng.pos;
            if (rng.exhausted) raise_error(d.gs, MCRAT_B200_ERR_REPLAY, -1, ERR_SITE_CS_EMIT);
        }
        return;
    }
    for (int k = blockIdx.x * 256 + threadIdx.x; k < ph_tot; k += gridDim.x * 256) {
        // the cell this photon belongs to: last j with offset[j] <= k
        int lo = 0, hi = n_sel;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (w.offset[mid] <= k) lo = mid; else hi = mid;
        }
        KeyedRng rng(d.k0, d.k1 ^ d.shard_base, 5u, (uint32_t)k, (uint64_t)epoch << 32);
        cs_fill_photon(d, rng, w.sel_cell[lo], w.nu_c[lo], weight, emitted[k]);
    }
}

// step 4: null slots per 256-slot block (then rebin_scan_kernel), and placement of photon k into the k-th null slot
__global__ void __launch_bounds__(256) count_null_kernel(DevCtx d)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int is_null = (i < d.cap) && (d.ph.type[i] == 'N');
    const int c = __syncthreads_count(is_null);
    if (threadIdx.x == 0) d.prefix_block[blockIdx.x] = c;
}

__global__ void __launch_bounds__(256) place_photons_kernel(DevCtx d, const mcrat_photon *src, int n_src)
{
    __shared__ int warp_off[8];
    const int i = blockIdx.x * 256 + threadIdx.x;
    const bool is_null = (i < d.cap) && (d.ph.type[i] == 'N');
    const unsigned ball = __ballot_sync(0xffffffffu, is_null);
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    if (lane == 0) warp_off[wp] = __popc(ball);
    __syncthreads();
    int k = d.prefix_block[blockIdx.x];
    for (int q = 0; q < wp; ++q) k += warp_off[q];
    k += __popc(ball & ((1u << lane) - 1u));
    if (!is_null || k >= n_src) return;
    const mcrat_photon p = src[k];
    if (p.type == 'N') return;
    d.ph.type[i] = p.type;
    store_momentum(d.ph, i, p.p0, p.p1, p.p2, p.p3);
    d.ph.c0[i] = p.comv_p0; d.ph.c1[i] = p.comv_p1; d.ph.c2[i] = p.comv_p2; d.ph.c3[i] = p.comv_p3;
    d.ph.r0[i] = p.r0; d.ph.r1[i] = p.r1; d.ph.r2[i] = p.r2;
    d.ph.safe[i] = 0;
    d.ph.s0[i] = p.s0; d.ph.s1[i] = p.s1; d.ph.s2[i] = p.s2; d.ph.s3[i] = p.s3;
    d.ph.nscatt[i] = p.num_scatt;
    d.ph.weight[i] = p.weight;
    d.ph.idx[i] = p.nearest_block_index;
    d.ph.tts[i] = p.time_to_scatter;
    store_tau(d.ph, i, p.total_optical_depth);
    unsigned char f = 0;
    if ((p.type != 'p') && (p.weight != 0)) f |= F_MOVABLE; // Src/mclib.c:1070
    if (p.recalc_properties == 1) f |= F_RECALC;
    d.ph.flags[i] = f;
}

// new slots of a grown list are null photons (reallocatePhotonListMemory, Src/photons.c:57-80)
__global__ void __launch_bounds__(256) init_null_kernel(DevCtx d, int first, int n)
{
    for (int j = blockIdx.x * 256 + threadIdx.x; j < n; j += gridDim.x * 256) {
        const int i = first + j;
        set_null_photon(d, i);
        d.ph.tts[i] = 0;
    }
}

// photonEmitCyclosynch with inject_single_switch == 1 behind the step-by-step surface (Src/mcrat.c:792-803 after a
// photonEvent call): the scattered pool photon becomes a comptonised one, a fresh pool photon goes into the first null
// slot, the scattered one is re-positioned inside its cell -- cs_emit_single, the statements the frame loop runs in its
// event kernel, continuing the event's own uniform stream where the event stopped (ShardState.last_event_draw; replay:
// the recorded stream's cursor).  out[0] = slot of the new photon, or -1: no null slot, the host grows the list first.
__global__ void __launch_bounds__(256) cs_emit_single_kernel(DevCtx d, int scatt, int *out)
{
    ShardState &st = d.sh[0];
    int first_null = INT_MAX;
    for (int j = threadIdx.x; j < st.count; j += 256)
        if (d.ph.type[st.first + j] == 'N') {
            first_null = st.first + j;
            break;
        }
    double dummy = 0;
    block_argmin<256>(dummy, first_null);
    if (threadIdx.x != 0) return;
    if (first_null == INT_MAX) {
        out[0] = -1;
        return;
    }
    GlobalState &gs = *d.gs;
    EventRng rng;
    rng.replay = d.replay;
    rng.k0 = d.k0;
    rng.k1 = d.k1 ^ d.shard_base;
    rng.iter = st.iter - 1; // the event that scattered this photon
    rng.draw = st.last_event_draw;
    rng.buf = d.replay_buf;
    rng.pos = d.replay ? gs.replay_cursor : 0;
    rng.n = d.replay ? gs.replay_n : 0;
    rng.exhausted = 0;
    rng.pre = nullptr;
    rng.npre = 0;
    gs.cs_comptonized_w += d.ph.weight[scatt]; // Src/mcrat.c:794-795
    d.ph.type[scatt] = 'k';
    if (d.ph.weight[scatt] != 0) d.ph.flags[scatt] |= F_MOVABLE;
    cs_emit_single(d, rng, scatt, first_null);
    gs.cs_emitted += 1;
    gs.cs_scatt_num += 1;
    st.last_event_draw = rng.draw;
    if (d.replay) {
        gs.replay_cursor = rng.pos;
        if (rng.exhausted) raise_error(&gs, MCRAT_B200_ERR_REPLAY, scatt, ERR_SITE_CS_EMIT);
    }
    out[0] = first_null;
}
