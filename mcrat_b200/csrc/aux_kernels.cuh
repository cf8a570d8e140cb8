// aux_kernels.cuh -- K5 absorption, K7 hot cross-section table, statistics, K8 rebin, peak probes.
// Part of the single translation unit mcrat_b200.cu (included there, in this order); not a stand-alone header.
#pragma once

// ------------------------------------------------------------------------------------------
// K5: phAbsCyclosynch, Src/mc_cyclosynch.c:1571-1644
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cs_absorb_kernel(DevCtx d)
{
    int abs_cnt = 0, scatt_cnt = 0;
    double abs_w = 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < d.cap; i += gridDim.x * 256) {
        const double w = d.ph.weight[i];
        const int idx = d.ph.idx[i];
        if ((w != 0) && (idx != -1)) {
            double b;
            if (d.b_calc == B_TOTAL_E || d.b_calc == B_INTERNAL_E) {
                double el_dens = d.cells.dens[idx] / M_P;
                b = calc_b(d.b_calc, d.epsilon_b, el_dens, d.cells.temp[idx]);
            } else if (d.dims == D_TWO) {
                double b0 = d.cells.B0[idx], b1 = d.cells.B1[idx];
                b = sqrt(b0 * b0 + b1 * b1);
            } else {
                double b0 = d.cells.B0[idx], b1 = d.cells.B1[idx], b2 = d.cells.B2[idx];
                b = sqrt(b0 * b0 + b1 * b1 + b2 * b2);
            }
            const double nu_c = calc_cyclotron_freq(b);
            const char type = d.ph.type[i];
            if ((d.ph.c0[i] * C_LIGHT / PL_CONST <= nu_c) || (type == 'p')) {
                abs_cnt++;
                if (!((type != 'i') && (type != 'c'))) abs_w += w;
                // setNullPhoton, Src/photons.c:208-251
                d.ph.type[i] = 'N';
                d.ph.weight[i] = 0;
                d.ph.idx[i] = -1;
                d.ph.safe[i] = 0;
                d.ph.flags[i] = 0;
                store_momentum(d.ph, i, 0, 0, 0, 0);
                d.ph.c0[i] = 0; d.ph.c1[i] = 0; d.ph.c2[i] = 0; d.ph.c3[i] = 0;
                d.ph.r0[i] = 0; d.ph.r1[i] = 0; d.ph.r2[i] = 0;
                d.ph.s0[i] = 0; d.ph.s1[i] = 0; d.ph.s2[i] = 0; d.ph.s3[i] = 0;
                d.ph.nscatt[i] = 0;
                store_tau(d.ph, i, 0);
            } else if ((type == 'k') || (type == 'c')) {
                scatt_cnt++;
            }
        }
    }
    // block reduction, then one atomic per block
    __shared__ int s_abs[8], s_sc[8];
    __shared__ double s_w[8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        abs_cnt += __shfl_xor_sync(0xffffffffu, abs_cnt, off);
        scatt_cnt += __shfl_xor_sync(0xffffffffu, scatt_cnt, off);
        abs_w += __shfl_xor_sync(0xffffffffu, abs_w, off);
    }
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    if (lane == 0) {
        s_abs[wp] = abs_cnt;
        s_sc[wp] = scatt_cnt;
        s_w[wp] = abs_w;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int a = 0, s = 0;
        double w = 0;
        for (int k = 0; k < 8; ++k) {
            a += s_abs[k];
            s += s_sc[k];
            w += s_w[k];
        }
        if (a) atomicAdd(&d.gs->abs_count, a);
        if (s) atomicAdd(&d.gs->cs_scatt_count, s);
        if (w != 0) atomicAdd(&d.gs->abs_weight, w);
    }
}


// ------------------------------------------------------------------------------------------
// K7: thermal (hot) Klein-Nishina cross-section table, Src/hot_x_section.c:82-206.
// One block per table point (221 x 81); each point is the reference's plain Monte Carlo estimate
// (Src/hot_x_section.c:324-357: `calls` uniform samples of (gamma, mu) over
// [1, 1 + 12 theta] x [-1, 1], integrand = Maxwell-Juttner pdf x boosted KN cross section,
// result = 0.5 * volume * mean), drawn from a Philox stream keyed by the point.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hot_table_kernel(double *table, long long calls, uint32_t k0, uint32_t k1, int point0)
{
    const int point = point0 + blockIdx.x; // i * (N_T + 1) + j, the reference's loop order (:90-105); table[] starts at point0
    const int i = point / (N_T + 1), j = point - i * (N_T + 1);
    const double dt = (LOG_T_MAX - LOG_T_MIN) / N_T, dph_e = (LOG_PH_E_MAX - LOG_PH_E_MIN) / N_PH_E;
    const double comv_ph_e = pow(10., LOG_PH_E_MIN + i * dph_e);
    const double theta = pow(10., LOG_T_MIN + j * dt);
    double result;
    if (theta < pow(10., LOG_T_MIN) && comv_ph_e < pow(10., LOG_PH_E_MIN)) {
        result = 1; // :336-337
    } else if (theta < pow(10., LOG_T_MIN)) {
        result = kn_cross_section(comv_ph_e); // :338-339
    } else {
        const double normalization = maxwell_juttner_norm(theta);
        const double xl0 = 1, xu0 = 1. + 12 * theta;
        double sum = 0;
        for (long long n = threadIdx.x; n < calls; n += 256) {
            double u1, u2;
            philox_doubles((uint32_t)n, (uint32_t)(n >> 32), (uint32_t)point, 2u, k0, k1, u1, u2);
            double gamma = xl0 + u1 * (xu0 - xl0);
            double mu = -1 + u2 * (1 - (-1));
            sum += maxwell_juttner_pdf(gamma, theta, normalization) * boosted_cross_section(comv_ph_e, mu, gamma);
        }
        __shared__ double red[8];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
        __syncthreads();
        double tot = 0;
        for (int k = 0; k < 8; ++k) tot += red[k];
        const double vol = (xu0 - xl0) * (1 - (-1));
        result = 0.5 * (vol * (tot / (double)calls));
    }
    if (threadIdx.x == 0) table[blockIdx.x] = log10(result);
}

// ------------------------------------------------------------------------------------------
// photon statistics (Src/mclib.c:1358-1515): per-block partials, finished on the host
// ------------------------------------------------------------------------------------------
struct StatPartial {
    double e_sum, w_sum, ns_sum, r_sum, r_min, r_max, th_min, th_max;
    long long count;
    int ns_max, ns_min;
};

__global__ void __launch_bounds__(256) stats_kernel(DevCtx d, StatPartial *out)
{
    StatPartial a;
    a.e_sum = a.w_sum = a.ns_sum = a.r_sum = 0;
    a.r_min = DBL_MAX; a.r_max = 0; a.th_min = DBL_MAX; a.th_max = 0;
    a.count = 0; a.ns_max = 0; a.ns_min = INT_MAX;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < d.cap; i += gridDim.x * 256) {
        const double w = d.ph.weight[i];
        const bool live = (w != 0);
        if (!d.cs || live) { // Src/mclib.c:1373-1379, 1401-1404
            a.e_sum += d.ph.p0[i] * w;
            a.w_sum += w;
            double ns = d.ph.nscatt[i];
            double r = sqrt(d.ph.r0[i] * d.ph.r0[i] + d.ph.r1[i] * d.ph.r1[i] + d.ph.r2[i] * d.ph.r2[i]);
            a.ns_sum += ns;
            a.r_sum += r;
            if (ns > a.ns_max) a.ns_max = (int)ns;
            if (ns < a.ns_min) a.ns_min = (int)ns;
            a.count++;
        }
        if (live) { // Src/mclib.c:1479-1508
            double r = sqrt(d.ph.r0[i] * d.ph.r0[i] + d.ph.r1[i] * d.ph.r1[i] + d.ph.r2[i] * d.ph.r2[i]);
            double th = acos(d.ph.r2[i] / r);
            if (r > a.r_max) a.r_max = r;
            if (r < a.r_min) a.r_min = r;
            if (th > a.th_max) a.th_max = th;
            if (th < a.th_min) a.th_min = th;
        }
    }
    __shared__ StatPartial sh[256];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            StatPartial &x = sh[threadIdx.x];
            const StatPartial &y = sh[threadIdx.x + s];
            x.e_sum += y.e_sum; x.w_sum += y.w_sum; x.ns_sum += y.ns_sum; x.r_sum += y.r_sum;
            x.r_min = fmin(x.r_min, y.r_min); x.r_max = fmax(x.r_max, y.r_max);
            x.th_min = fmin(x.th_min, y.th_min); x.th_max = fmax(x.th_max, y.th_max);
            x.count += y.count;
            x.ns_max = max(x.ns_max, y.ns_max); x.ns_min = min(x.ns_min, y.ns_min);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

// ------------------------------------------------------------------------------------------
// K8: rebinCyclosynchCompPhotons, Src/mc_cyclosynch.c:244-710, on the device.  The (log E, theta[, phi])
// binning of the reference, including gsl_histogram2d's uniform ranges and its find() (linear guess,
// then bisection); every bin's weighted sums are accumulated in slot order, as the reference's one
// sequential loop does, so the rebinned photons agree with the CPU's to rounding of acos / atan2 / sincos.
// ------------------------------------------------------------------------------------------
constexpr double RAD_TO_DEG = 180.0 / PI, DEG_TO_RAD = PI / 180.0; // Src/mcrat.h:80-81

struct RebinRange { // struct PhotonRangeInfo, Src/mc_cyclosynch.h
    double p0_min, p0_max, theta_min, theta_max, phi_min, phi_max;
    int valid_photon_count, synch_photon_count;
};

struct RebinParams { // struct BinningParams + the histogram ranges
    int num_bins, num_bins_theta, num_bins_phi, total_bins;
    const double *range_e, *range_theta, *range_phi; // num_bins+1, num_bins_theta+1, num_bins_phi+1 edges
};

struct RebinBin { // struct BinStats
    double weighted_r, weighted_theta, weighted_phi_offset, weighted_stokes[4], weighted_scatt_count, total_weight;
    double weighted_phi_dir, weighted_theta_dir, weighted_energy, weighted_phi_pos;
};

__device__ __forceinline__ bool rebin_eligible(char type) { return (type != 'N') && (type != 'p') && (type != 'i'); }

// calculate_photon_position, Src/mc_cyclosynch.c:246-270
__device__ __forceinline__ void rebin_position(int ndim3, double x, double y, double z, double &r, double &theta, double &phi)
{
    r = sqrt(x * x + y * y + z * z);
    if (r < DBL_MIN) {
        theta = 0.0;
        phi = 0.0;
    } else {
        theta = acos(z / r);
        if (ndim3) {
            double phi_rad = atan2(y, x);
            phi = fmod(phi_rad * RAD_TO_DEG + 360.0, 360.0);
        } else {
            phi = 0;
        }
    }
}

// collect_photon_statistics, Src/mc_cyclosynch.c:273-322 (per-block partials; min / max are exact in any order)
__global__ void __launch_bounds__(256) rebin_range_kernel(DevCtx d, RebinRange *out)
{
    const int ndim3 = (d.dims == D_THREE);
    RebinRange a;
    a.p0_min = DBL_MAX; a.p0_max = 0.0; a.theta_min = DBL_MAX; a.theta_max = 0.0; a.phi_min = DBL_MAX; a.phi_max = 0.0;
    a.valid_photon_count = 0; a.synch_photon_count = 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < d.cap; i += gridDim.x * 256) {
        const char type = d.ph.type[i];
        if (rebin_eligible(type)) {
            const double p0 = d.ph.p0[i];
            if (p0 > 0) {
                a.p0_min = fmin(a.p0_min, p0);
                a.p0_max = fmax(a.p0_max, p0);
                a.valid_photon_count++;
            }
            double r, theta, phi;
            rebin_position(ndim3, d.ph.r0[i], d.ph.r1[i], d.ph.r2[i], r, theta, phi);
            a.theta_min = fmin(a.theta_min, theta);
            a.theta_max = fmax(a.theta_max, theta);
            if (ndim3) {
                a.phi_min = fmin(a.phi_min, phi);
                a.phi_max = fmax(a.phi_max, phi);
            }
        }
        if (type == 'p') a.synch_photon_count++;
    }
    __shared__ RebinRange sh[256];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            RebinRange &x = sh[threadIdx.x];
            const RebinRange &y = sh[threadIdx.x + s];
            x.p0_min = fmin(x.p0_min, y.p0_min); x.p0_max = fmax(x.p0_max, y.p0_max);
            x.theta_min = fmin(x.theta_min, y.theta_min); x.theta_max = fmax(x.theta_max, y.theta_max);
            x.phi_min = fmin(x.phi_min, y.phi_min); x.phi_max = fmax(x.phi_max, y.phi_max);
            x.valid_photon_count += y.valid_photon_count; x.synch_photon_count += y.synch_photon_count;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

// gsl_histogram find(): 0 on success (histogram/find.c: linear guess, then bisection)
__device__ __forceinline__ int hist_find(int n, const double *range, double x, int &i)
{
    if (x < range[0] || x >= range[n]) return 1;
    {
        double u = (x - range[0]) / (range[n] - range[0]);
        size_t g = (size_t)(u * n);
        if (g < (size_t)n && x >= range[g] && x < range[g + 1]) {
            i = (int)g;
            return 0;
        }
    }
    int lower = 0, upper = n;
    while (upper - lower > 1) {
        int mid = (upper + lower) / 2;
        if (x >= range[mid])
            lower = mid;
        else
            upper = mid;
    }
    i = lower;
    return 0;
}

// bin index of every slot (or -1), Src/mc_cyclosynch.c:453-472
__global__ void __launch_bounds__(256) rebin_index_kernel(DevCtx d, RebinParams p, int *bin_of)
{
    const int ndim3 = (d.dims == D_THREE);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < d.cap; i += gridDim.x * 256) {
        int b = -1;
        if (rebin_eligible(d.ph.type[i])) {
            double r, theta, phi;
            rebin_position(ndim3, d.ph.r0[i], d.ph.r1[i], d.ph.r2[i], r, theta, phi);
            const double le = log10(d.ph.p0[i]);
            int ix = 0, iy = 0, iz = 0;
            // gsl_histogram2d_find(h_energy_theta, ...): the second index is only written if the first was found
            if (hist_find(p.num_bins, p.range_e, le, ix) == 0) hist_find(p.num_bins_theta, p.range_theta, theta, iy);
            if (ndim3) {
                if (hist_find(p.num_bins, p.range_e, le, ix) == 0) hist_find(p.num_bins_phi, p.range_phi, phi, iz);
                if (hist_find(p.num_bins_theta, p.range_theta, theta, iy) == 0) hist_find(p.num_bins_phi, p.range_phi, phi, iz);
            }
            // calculate_bin_index, :432-446
            if (ix < 0 || ix >= p.num_bins || iy < 0 || iy >= p.num_bins_theta)
                b = -2;
            else if (ndim3)
                b = (iz < 0 || iz >= p.num_bins_phi) ? -2 : iz * p.num_bins * p.num_bins_theta + ix * p.num_bins_theta + iy;
            else
                b = ix * p.num_bins_theta + iy;
            if (b == -2 || b >= p.total_bins) {
                raise_error(d.gs, MCRAT_B200_ERR_STATE, -1, ERR_SITE_REBIN); // the reference exits here (:469-472)
                b = -1;
            }
        }
        bin_of[i] = b;
    }
}

// accumulate_bin_statistics + create_rebinned_photons (:448-585): one thread per bin walks the list in slot
// order (warp-uniform reads of bin_of[]), which keeps the reference's order of additions inside every bin
__global__ void __launch_bounds__(128) rebin_accumulate_kernel(DevCtx d, RebinParams p, const int *bin_of, mcrat_photon *out)
{
    const int ndim3 = (d.dims == D_THREE);
    const int b = blockIdx.x * 128 + threadIdx.x;
    RebinBin s;
    memset(&s, 0, sizeof(s));
    for (int i = 0; i < d.cap; ++i) {
        if (bin_of[i] != b || b >= p.total_bins) continue;
        const double w = d.ph.weight[i];
        const double x = d.ph.r0[i], y = d.ph.r1[i], z = d.ph.r2[i];
        const double p0 = d.ph.p0[i], p1 = d.ph.p1[i], p2 = d.ph.p2[i], p3 = d.ph.p3[i];
        double r, theta, phi;
        rebin_position(ndim3, x, y, z, r, theta, phi);
        s.weighted_r += r * w;
        s.weighted_theta += theta * w;
        s.weighted_phi_offset += (atan2(p2, p1) - atan2(y, x)) * RAD_TO_DEG * w;
        s.weighted_stokes[0] += d.ph.s0[i] * w;
        s.weighted_stokes[1] += d.ph.s1[i] * w;
        s.weighted_stokes[2] += d.ph.s2[i] * w;
        s.weighted_stokes[3] += d.ph.s3[i] * w;
        s.weighted_scatt_count += d.ph.nscatt[i] * w;
        s.total_weight += w;
        double phi_dir = fmod(atan2(p2, p1) * RAD_TO_DEG + 360.0, 360.0);
        double theta_dir = acos(p3 / p0) * RAD_TO_DEG;
        s.weighted_phi_dir += phi_dir * w;
        s.weighted_theta_dir += theta_dir * w;
        s.weighted_energy += p0 * w;
        if (ndim3) s.weighted_phi_pos += phi * w;
    }
    if (b >= p.total_bins) return;
    mcrat_photon q;
    memset(&q, 0, sizeof(q)); // calloc'ed in the reference (:505)
    if (s.total_weight <= 0) {
        q.type = 'N';
        q.weight = 0;
        q.nearest_block_index = -1;
        q.recalc_properties = 0;
    } else {
        q.type = 'k';
        q.weight = s.total_weight;
        double avg_energy = s.weighted_energy / s.total_weight;
        double avg_phi_dir = s.weighted_phi_dir / s.total_weight;
        double avg_theta_dir = s.weighted_theta_dir / s.total_weight;
        double avg_r = s.weighted_r / s.total_weight;
        double avg_theta_pos = s.weighted_theta / s.total_weight;
        q.p0 = avg_energy;
        q.p1 = avg_energy * sin(avg_theta_dir * DEG_TO_RAD) * cos(avg_phi_dir * DEG_TO_RAD);
        q.p2 = avg_energy * sin(avg_theta_dir * DEG_TO_RAD) * sin(avg_phi_dir * DEG_TO_RAD);
        q.p3 = avg_energy * cos(avg_theta_dir * DEG_TO_RAD);
        double pos_phi;
        if (ndim3) {
            double avg_phi_pos = s.weighted_phi_pos / s.total_weight;
            pos_phi = avg_phi_pos * DEG_TO_RAD;
        } else {
            double avg_phi_offset = s.weighted_phi_offset / s.total_weight;
            pos_phi = (avg_phi_dir - avg_phi_offset) * DEG_TO_RAD;
        }
        q.r0 = avg_r * sin(avg_theta_pos) * cos(pos_phi);
        q.r1 = avg_r * sin(avg_theta_pos) * sin(pos_phi);
        q.r2 = avg_r * cos(avg_theta_pos);
        q.s0 = s.weighted_stokes[0] / s.total_weight;
        q.s1 = s.weighted_stokes[1] / s.total_weight;
        q.s2 = s.weighted_stokes[2] / s.total_weight;
        q.s3 = s.weighted_stokes[3] / s.total_weight;
        q.num_scatt = (int)(s.weighted_scatt_count / s.total_weight + 0.5);
        q.nearest_block_index = 0;
        q.recalc_properties = 1;
    }
    out[b] = q;
}

__device__ __forceinline__ void set_null_photon(DevCtx &d, int i) // setNullPhoton, Src/photons.c:208-251
{
    d.ph.type[i] = 'N';
    d.ph.weight[i] = 0;
    d.ph.idx[i] = -1;
    d.ph.safe[i] = 0;
    d.ph.flags[i] = 0;
    store_momentum(d.ph, i, 0, 0, 0, 0);
    d.ph.c0[i] = 0; d.ph.c1[i] = 0; d.ph.c2[i] = 0; d.ph.c3[i] = 0;
    d.ph.r0[i] = 0; d.ph.r1[i] = 0; d.ph.r2[i] = 0;
    d.ph.s0[i] = 0; d.ph.s1[i] = 0; d.ph.s2[i] = 0; d.ph.s3[i] = 0;
    d.ph.nscatt[i] = 0;
    store_tau(d.ph, i, 0);
}

// :588-596 null every 'k' / 'c' photon, then count the null slots of each 256-slot block
__global__ void __launch_bounds__(256) rebin_null_kernel(DevCtx d)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    int is_null = 0;
    if (i < d.cap) {
        const char t = d.ph.type[i];
        if (t == 'c' || t == 'k') set_null_photon(d, i);
        is_null = (t == 'c' || t == 'k' || t == 'N');
    }
    const int c = __syncthreads_count(is_null);
    if (threadIdx.x == 0) d.prefix_block[blockIdx.x] = c;
}

__global__ void rebin_scan_kernel(DevCtx d, int nblocks, int *total_null)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int run = 0;
        for (int b = 0; b < nblocks; ++b) {
            int c = d.prefix_block[b];
            d.prefix_block[b] = run;
            run += c;
        }
        *total_null = run;
    }
}

// addToPhotonList (Src/photons.c:167-205): rebinned photon k goes to the k-th null slot of the list
__global__ void __launch_bounds__(256) rebin_place_kernel(DevCtx d, const mcrat_photon *rebinned, int total_bins)
{
    __shared__ int warp_off[8];
    const int i = blockIdx.x * 256 + threadIdx.x;
    const bool is_null = (i < d.cap) && (d.ph.type[i] == 'N');
    const unsigned ball = __ballot_sync(0xffffffffu, is_null);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_off[w] = __popc(ball);
    __syncthreads();
    int k = d.prefix_block[blockIdx.x];
    for (int q = 0; q < w; ++q) k += warp_off[q];
    k += __popc(ball & ((1u << lane) - 1u));
    if (!is_null || k >= total_bins) return;
    const mcrat_photon p = rebinned[k];
    if (p.type == 'N') return; // only the non-null rebinned photons are copied (:190-199)
    d.ph.type[i] = p.type;
    store_momentum(d.ph, i, p.p0, p.p1, p.p2, p.p3);
    d.ph.c0[i] = 0; d.ph.c1[i] = 0; d.ph.c2[i] = 0; d.ph.c3[i] = 0;
    d.ph.r0[i] = p.r0; d.ph.r1[i] = p.r1; d.ph.r2[i] = p.r2;
    d.ph.safe[i] = 0;
    d.ph.s0[i] = p.s0; d.ph.s1[i] = p.s1; d.ph.s2[i] = p.s2; d.ph.s3[i] = p.s3;
    d.ph.nscatt[i] = p.num_scatt;
    d.ph.weight[i] = p.weight;
    d.ph.idx[i] = p.nearest_block_index;
    d.ph.tts[i] = 0;
    store_tau(d.ph, i, 0);
    d.ph.flags[i] = (unsigned char)(((p.weight != 0) ? F_MOVABLE : 0) | F_RECALC);
}

__global__ void rebin_count_kernel(const mcrat_photon *rebinned, int total_bins, int *null_bins)
{
    int c = 0;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < total_bins; b += gridDim.x * blockDim.x)
        if (rebinned[b].type == 'N') c++;
    if (c) atomicAdd(null_bins, c);
}

// ------------------------------------------------------------------------------------------
// peak probes (roofline denominators measured on the same GPU, same run)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double seed)
{
    // FP64-pipe issue rate: 16 independent DFMA chains per thread, 64 warps per SM
    double x[16];
    const double a = 1.0 + seed * 1e-9, b = seed * 1e-12;
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = seed + k + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += x[k];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) copy_kernel(const double4 *__restrict__ a, double4 *__restrict__ b, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
