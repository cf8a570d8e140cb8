// pass_kernels.cuh -- device helpers, AoS <-> SoA, K4+K2 fused pass.
// Part of the single translation unit mcrat_b200.cu (included there, in this order); not a stand-alone header.
#pragma once

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
// every kernel of the frame loop evaluates the same stop condition, so one iteration of a shard
// is either executed completely or not at all
__device__ __forceinline__ bool loop_stopped(const GlobalState &gs, const ShardState &sh)
{
    return (gs.error != 0) | sh.done | sh.pause_cs | (gs.max_iters >= 0 && sh.iters_done >= gs.max_iters);
}

__device__ __forceinline__ int shard_of(const DevCtx &d, int slot) { return slot / d.shard_size; }

__device__ __forceinline__ bool lex_less(double ta, int ia, double tb, int ib) { return (ta < tb) || (ta == tb && ia < ib); }

// ---- skipping the containment re-check while a photon provably cannot have left its cell ------------------------
// The reference re-checks every photon's cached cell in every iteration (Src/mclib.c:469-597), and almost always
// finds it unchanged.  A photon that sits at distance >= dist from the boundary of (its cell intersected with the domain)
// stays inside while the total length of its pushes is < dist, whatever its direction; all movable photons of a
// shard are pushed by the same times, so one counter per shard (ShardState.path, integer units, every push rounded
// up, plus a bound on the rounding of the position update) and one threshold per photon (PhotonCols.safe) decide it.
// The re-check that is skipped has no side effect when it succeeds, so the photons are bit-identical; the margins
// (safe_distance) are far above the rounding of the reference's coordinate evaluation.
constexpr double PATH_SCALE = 256.0;
constexpr unsigned long long PATH_SAT = 1ull << 62;

__device__ __forceinline__ unsigned long long path_units(double dt, double pad)
{
    const double x = (C_LIGHT * fabs(dt) * (1.0 + 1e-9) + pad) * PATH_SCALE;
    if (!(x < 4e18)) return PATH_SAT;
    return __double2ull_ru(x);
}

__device__ __forceinline__ unsigned long long path_add(unsigned long long a, unsigned long long b)
{
    const unsigned long long c = a + b;
    return (a >= PATH_SAT || b >= PATH_SAT || c >= PATH_SAT) ? PATH_SAT : c;
}

// the pending pushes become part of the path; called wherever a shard's push list is replaced or cleared
__device__ __forceinline__ void fold_path(ShardState &sh, double pad)
{
    unsigned long long p = sh.path;
    for (int k = 0; k < sh.n_dt; ++k) p = path_add(p, path_units(sh.dt_list[k], pad));
    sh.path = p;
}

__device__ __forceinline__ unsigned long long path_after_pending(const ShardState &sh, double pad)
{
    unsigned long long p = sh.path;
    for (int k = 0; k < sh.n_dt; ++k) p = path_add(p, path_units(sh.dt_list[k], pad));
    return p;
}

// Lower bound (cm) on the Euclidean distance from the photon at hydro coordinates h (inside cell blk and inside the
// domain) to the nearest point outside either.  Per coordinate the margin m = min(half size - |h - c|, h - dom_lo,
// dom_hi - h); a length coordinate (x, y, z, cylindrical / spherical radius) is 1-Lipschitz in the position, an
// angle seen from the origin (axis) changes by at most asin(length / r) (asin(length / rho)), and sin(m) >= 0.8 m
// on [0, 1].  Margins below 1e-7 of the coordinate's scale (1e-6 rad) give 0: never skipped.  The factor 0.5
// leaves half of every margin for the rounding of the coordinates themselves.
__device__ __forceinline__ double margin_length(double h, double c, double hs, double lo, double hi)
{
    const double m = fmin(hs - fabs(h - c), fmin(h - lo, hi - h));
    return (m > 1e-7 * fmax(fabs(h), fabs(c))) ? m : 0.0;
}

__device__ __forceinline__ double margin_angle(double h, double c, double hs, double lo, double hi, double full, double lever)
{
    double m = fmin(hs - fabs(h - c), fmin(h - lo, hi - h));
    m = fmin(m, fmin(h, full - h)); // the pole / the wrap of the azimuth
    return (m > 1e-6) ? 0.8 * lever * fmin(m, 1.0) : 0.0;
}

// not inlined, arguments by value: the pass kernel runs at 64 registers and takes this path for a fraction of a
// percent of its photons.  dom: the domain bounds in global memory (DevCtx.dom_dev).
__device__ __noinline__ unsigned long long safe_path(const double4 *geoA, const double2 *geoB, const double *dom, int geom,
                                                      int ndim3, int blk, double h0, double h1, double h2, double v0,
                                                      double v1, double v2, int movable, unsigned long long s_now)
{
    if (movable) { // |v| <= c up to rounding is what path_units assumes
        const double b2 = (v0 * v0 + v1 * v1 + v2 * v2) / (C_LIGHT * C_LIGHT);
        if (!(b2 <= 1.0 + 1e-9)) return 0ull;
    }
    const double4 a = geoA[blk];
    double dist;
    if (!ndim3) { // (c0, c1, h0, h1)
        dist = margin_length(h0, a.x, a.z, dom[0], dom[1]);
        if (geom == G_SPHERICAL)
            dist = fmin(dist, margin_angle(h1, a.y, a.w, dom[2], dom[3], PI, h0));
        else
            dist = fmin(dist, margin_length(h1, a.y, a.w, dom[2], dom[3]));
    } else { // (c0, c1, c2, h0) + (h1, h2)
        const double2 b = geoB[blk];
        if (geom == G_SPHERICAL) {
            dist = margin_length(h0, a.x, a.w, dom[0], dom[1]);
            dist = fmin(dist, margin_angle(h1, a.y, b.x, dom[2], dom[3], PI, h0));
            dist = fmin(dist, margin_angle(h2, a.z, b.y, dom[4], dom[5], 2.0 * PI, h0 * sin(h1)));
        } else if (geom == G_POLAR) {
            dist = margin_length(h0, a.x, a.w, dom[0], dom[1]);
            dist = fmin(dist, margin_angle(h1, a.y, b.x, dom[2], dom[3], 2.0 * PI, h0));
            dist = fmin(dist, margin_length(h2, a.z, b.y, dom[4], dom[5]));
        } else {
            dist = margin_length(h0, a.x, a.w, dom[0], dom[1]);
            dist = fmin(dist, margin_length(h1, a.y, b.x, dom[2], dom[3]));
            dist = fmin(dist, margin_length(h2, a.z, b.y, dom[4], dom[5]));
        }
    }
    if (!(dist > 0 && dist < 1e300)) return 0ull; // also NaN
    const double x = 0.5 * dist * PATH_SCALE;
    if (!(x >= 1.0)) return 0ull;
    const unsigned long long u = (x < 4e18) ? __double2ull_rd(x) : PATH_SAT;
    const unsigned long long t = path_add(s_now, u);
    return t >= PATH_SAT ? PATH_SAT - 1 : t;
}

// warp-shuffle + shared-memory arg-min over (time, slot); ties broken by lowest slot
template <int THREADS>
__device__ __forceinline__ void block_argmin(double &t, int &i)
{
    __shared__ double sh_t[THREADS / 32];
    __shared__ int sh_i[THREADS / 32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ot = __shfl_xor_sync(0xffffffffu, t, off);
        int oi = __shfl_xor_sync(0xffffffffu, i, off);
        if (lex_less(ot, oi, t, i)) {
            t = ot;
            i = oi;
        }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) {
        sh_t[w] = t;
        sh_i[w] = i;
    }
    __syncthreads();
    if (w == 0) {
        t = (lane < THREADS / 32) ? sh_t[lane] : DBL_MAX;
        i = (lane < THREADS / 32) ? sh_i[lane] : INT_MAX;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double ot = __shfl_xor_sync(0xffffffffu, t, off);
            int oi = __shfl_xor_sync(0xffffffffu, i, off);
            if (lex_less(ot, oi, t, i)) {
                t = ot;
                i = oi;
            }
        }
    }
}

__device__ __forceinline__ bool in_cell(int ndim3, const CellCols &c, int blk, double h0, double h1, double h2)
{
    // Src/geometry.c:394-417 checkInBlock: 2|x-c| - size <= 0  <=>  |x-c| <= size/2 (both exact scalings)
    double4 a = c.geoA[blk];
    if (!ndim3) return (fabs(h0 - a.x) <= a.z) & (fabs(h1 - a.y) <= a.w);
    double2 b = c.geoB[blk];
    return (fabs(h0 - a.x) <= a.w) & (fabs(h1 - a.y) <= b.x) & (fabs(h2 - a.z) <= b.y);
}

__device__ __forceinline__ CellState load_cell_state(const CellCols &c, int i)
{
    CellState s;
    s.v0 = c.v0[i];
    s.v1 = c.v1[i];
    s.v2 = c.v2[i];
    s.r0 = c.r0[i];
    s.r1 = c.r1[i];
    s.r2 = c.r2[i];
    s.gamma = c.gamma[i];
    s.dens_lab = c.dens_lab[i];
    s.temp = c.temp[i];
    return s;
}

__device__ __forceinline__ void fluid_beta_of(const DevCtx &d, const CellState &c, double ph_r0, double ph_r1, double *fb)
{
    if (d.dims == D_THREE) {
        hydro_vector_to_cartesian(d.dims, d.geom, fb, c.v0, c.v1, c.v2, c.r0, c.r1, c.r2);
    } else if (d.dims == D_TWO_POINT_FIVE) {
        double ph_phi = atan2(ph_r1, ph_r0);
        hydro_vector_to_cartesian(d.dims, d.geom, fb, c.v0, c.v1, c.v2, c.r0, c.r1, ph_phi);
    } else {
        double ph_phi = atan2(ph_r1, ph_r0);
        hydro_vector_to_cartesian(d.dims, d.geom, fb, c.v0, c.v1, 0, c.r0, c.r1, ph_phi);
    }
}

// time_to_scatter of an in-domain photon, Src/mclib.c:675-687
__device__ __forceinline__ double free_path_time(double tau, double xi)
{
    double mfp = (-1.0 / tau) * log(xi);
    return mfp / C_LIGHT;
}

// x / C_LIGHT, correctly rounded, without the general division: q = RN(x * rc) with rc = RN(1 / C_LIGHT) is within
// an ulp of the quotient, r = x - q * C_LIGHT is exact in one FMA, and RN(q + r * rc) is the correctly rounded
// quotient (Markstein's theorem; C_LIGHT's significand is not all ones).  Outside the range where q, r stay normal
// and finite the true division runs.  tests/test_gpu_parity.py::test_division_by_c_is_exact compares the two bit
// for bit; tools/div_by_c_check.c does so on the host over 4e9 significands.
__device__ __forceinline__ double div_by_c(double x)
{
    const double rc = 1.0 / C_LIGHT; // folded at compile time, correctly rounded
    const double ax = fabs(x);
    if (!(ax > 1e-280 && ax < 1e300)) return x / C_LIGHT;
    const double q = x * rc;
    const double r = fma(-q, C_LIGHT, x);
    return fma(r, rc, q);
}

// free_path_time with ntau = -1/tau already formed (PhotonCols.ntau): bit-identical to it
__device__ __forceinline__ double free_path_time_n(double ntau, double xi)
{
    double mfp = ntau * log(xi);
    return div_by_c(mfp);
}

__device__ __forceinline__ void store_momentum(PhotonCols &ph, int i, double p0, double p1, double p2, double p3)
{
    ph.p0[i] = p0; ph.p1[i] = p1; ph.p2[i] = p2; ph.p3[i] = p3;
    const double div = 1.0 / p0; // Src/mclib.c:1074
    ph.v0[i] = p1 * div * C_LIGHT;
    ph.v1[i] = p2 * div * C_LIGHT;
    ph.v2[i] = p3 * div * C_LIGHT;
}

__device__ __forceinline__ void store_tau(PhotonCols &ph, int i, double tau)
{
    ph.tau[i] = tau;
    ph.ntau[i] = -1.0 / tau;
}

// the pushes with v_k = (p_k / p0) * C_LIGHT already formed (PhotonCols.v*): bit-identical to apply_pushes
__device__ __forceinline__ void apply_pushes_v(const ShardState &sh, int n_dt, double v0, double v1, double v2, double &r0,
                                               double &r1, double &r2)
{
    for (int k = 0; k < n_dt; ++k) {
        double t = sh.dt_list[k];
        r0 += v0 * t;
        r1 += v1 * t;
        r2 += v2 * t;
    }
}

// pending pushes of a shard's last event, applied one by one: the reference pushes once per
// candidate it tries (Src/mclib.c:1138, 1332) and FP addition is not associative
__device__ __forceinline__ void apply_pushes(const ShardState &sh, int n_dt, double p0, double p1, double p2, double p3,
                                             double &r0, double &r1, double &r2)
{
    double div = 1.0 / p0; // Src/mclib.c:1074-1080
    for (int k = 0; k < n_dt; ++k) {
        double t = sh.dt_list[k];
        r0 += p1 * div * C_LIGHT * t;
        r1 += p2 * div * C_LIGHT * t;
        r2 += p3 * div * C_LIGHT * t;
    }
}

// ------------------------------------------------------------------------------------------
// AoS <-> SoA (the boundary: `struct photon` records <-> device columns)
// ------------------------------------------------------------------------------------------
__global__ void unpack_kernel(DevCtx d, const mcrat_photon *aos, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        mcrat_photon p = aos[i];
        d.ph.type[i] = p.type;
        store_momentum(d.ph, i, p.p0, p.p1, p.p2, p.p3);
        d.ph.c0[i] = p.comv_p0; d.ph.c1[i] = p.comv_p1; d.ph.c2[i] = p.comv_p2; d.ph.c3[i] = p.comv_p3;
        d.ph.r0[i] = p.r0; d.ph.r1[i] = p.r1; d.ph.r2[i] = p.r2;
        d.ph.s0[i] = p.s0; d.ph.s1[i] = p.s1; d.ph.s2[i] = p.s2; d.ph.s3[i] = p.s3;
        d.ph.nscatt[i] = p.num_scatt;
        d.ph.weight[i] = p.weight;
        d.ph.idx[i] = p.nearest_block_index;
        d.ph.tts[i] = p.time_to_scatter;
        store_tau(d.ph, i, p.total_optical_depth);
        d.ph.safe[i] = 0;
        unsigned char f = 0;
        if ((p.type != 'p') && (p.weight != 0)) f |= F_MOVABLE; // Src/mclib.c:1070
        if (p.recalc_properties == 1) f |= F_RECALC;
        d.ph.flags[i] = f;
    }
}

__global__ void pack_kernel(DevCtx d, mcrat_photon *aos, int first, int n)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        int i = first + j;
        mcrat_photon p;
        memset(&p, 0, sizeof(p));
        p.type = d.ph.type[i];
        p.p0 = d.ph.p0[i]; p.p1 = d.ph.p1[i]; p.p2 = d.ph.p2[i]; p.p3 = d.ph.p3[i];
        p.comv_p0 = d.ph.c0[i]; p.comv_p1 = d.ph.c1[i]; p.comv_p2 = d.ph.c2[i]; p.comv_p3 = d.ph.c3[i];
        p.r0 = d.ph.r0[i]; p.r1 = d.ph.r1[i]; p.r2 = d.ph.r2[i];
        p.s0 = d.ph.s0[i]; p.s1 = d.ph.s1[i]; p.s2 = d.ph.s2[i]; p.s3 = d.ph.s3[i];
        p.num_scatt = d.ph.nscatt[i];
        p.recalc_properties = (d.ph.flags[i] & F_RECALC) ? 1 : 0;
        p.weight = d.ph.weight[i];
        p.nearest_block_index = d.ph.idx[i];
        p.time_to_scatter = d.ph.tts[i];
        p.total_optical_depth = d.ph.tau[i];
        aos[j] = p;
    }
}

// cell SoA -> scan layout (centres + half sizes), NaN-padded so that padding never matches
__global__ void build_geo_kernel(int ndim3, int n, int n_padded, const double *c0, const double *c1, const double *c2,
                                 const double *s0, const double *s1, const double *s2, double4 *geoA, double2 *geoB)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_padded; i += gridDim.x * blockDim.x) {
        const double qnan = __longlong_as_double(0x7ff8000000000000ll);
        if (i < n) {
            if (!ndim3) {
                geoA[i] = make_double4(c0[i], c1[i], 0.5 * s0[i], 0.5 * s1[i]);
            } else {
                geoA[i] = make_double4(c0[i], c1[i], c2[i], 0.5 * s0[i]);
                geoB[i] = make_double2(0.5 * s1[i], 0.5 * s2[i]);
            }
        } else {
            geoA[i] = make_double4(qnan, qnan, qnan, qnan);
            if (ndim3) geoB[i] = make_double2(qnan, qnan);
        }
    }
}

// ------------------------------------------------------------------------------------------
// K4+K2: fused push + locate re-check + free-path draw + block arg-min.
// Grid = nshards x blocks_per_shard: a block never straddles two sub-shards.
// ------------------------------------------------------------------------------------------
#ifndef MCRAT_PASS_THREADS
#define MCRAT_PASS_THREADS 256
#endif
// 5 blocks of 256 threads per SM (48 registers; the few spilled words are L1 hits) and a grid of two full waves:
// 10^7 photons, 2-D: 174.6 us per pass against 187.1 us at 4 blocks / 64 registers and 183.3 us at 6 / 40 -- once
// the re-check is skipped the kernel is a latency-bound stream and the extra loads in flight pay
#ifndef MCRAT_PASS_MINB
#define MCRAT_PASS_MINB 5
#endif
#ifndef MCRAT_PASS_CTAS_PER_SM
#define MCRAT_PASS_CTAS_PER_SM 10
#endif
constexpr int PASS_THREADS = MCRAT_PASS_THREADS;

// One block's share of a shard: photons j = b*THREADS + tid, stride nblk*THREADS.
// LOCAL_RELOC = false: relocating photons go to the global list (gs.reloc_count[parity]) that
// K1/K1b/K1c + finish_kernel work off; true: to the shard's own region [first, first+count) of the
// list (persistent loop: shards advance independently of each other).
template <bool FUSE_MFP, bool LOCAL_RELOC, int THREADS>
__device__ __forceinline__ void pass_body(DevCtx &d, const ShardState &sh, const int s, const int b, const int nblk,
                                          const int sw, const int parity, double &best_t, int &best_i, int *reloc_flag = nullptr)
{
    const int n_dt = sh.n_dt;
    const int pushed = sh.pushed_slot;
    const unsigned long long iter = sh.iter;
    const uint32_t k1 = d.k1 ^ (d.shard_base + (uint32_t)s);
    const int ndim3 = (d.dims == D_THREE);
    const double default_t = 1e12 / C_LIGHT; // Src/mclib.c:620, 684-687
    const int first = sh.first, count = sh.count;
    // path counter once the pending pushes are applied (what this pass does), and whether re-checks may be skipped:
    // never on a new hydro frame (everything re-locates) nor in cyclo-synchrotron runs (Src/mclib.c:510-515 re-locates
    // by the photon's state, not its position)
    const unsigned long long s_now = path_after_pending(sh, d.path_pad);
    const bool may_skip = d.recheck_skip && sw == 0 && !d.cs && s_now < PATH_SAT;
    const bool verify = d.recheck_skip == 2;

    const int mini = LOCAL_RELOC ? sh.mini_slot : -1;
    for (int j = b * THREADS + threadIdx.x; j < count; j += nblk * THREADS) {
        const int i = first + j;
        if (i == mini) continue; // the event block runs this photon's pass itself (persistent loop)
        // every column this photon can need is requested up front (one round trip to HBM instead
        // of three dependent ones); the momentum is used by the pushes, tau by the free-path draw
        unsigned char flags;
        int idx = 0;
        unsigned long long safe = 0;
        double r0, r1, r2, v0, v1, v2, ntau;
        if (d.stream_hints) {
            flags = __ldcs(d.ph.flags + i);
            if (may_skip) safe = __ldcs(d.ph.safe + i); else idx = __ldcs(d.ph.idx + i);
            r0 = __ldcs(d.ph.r0 + i); r1 = __ldcs(d.ph.r1 + i); r2 = __ldcs(d.ph.r2 + i);
            v0 = __ldcs(d.ph.v0 + i); v1 = __ldcs(d.ph.v1 + i); v2 = __ldcs(d.ph.v2 + i);
            ntau = FUSE_MFP ? __ldcs(d.ph.ntau + i) : 0.0;
        } else {
            flags = d.ph.flags[i];
            if (may_skip) safe = d.ph.safe[i]; else idx = d.ph.idx[i];
            r0 = d.ph.r0[i]; r1 = d.ph.r1[i]; r2 = d.ph.r2[i];
            v0 = d.ph.v0[i]; v1 = d.ph.v1[i]; v2 = d.ph.v2[i];
            ntau = FUSE_MFP ? d.ph.ntau[i] : 0.0;
        }
        if (n_dt > 0 && (flags & F_MOVABLE) && i != pushed) {
            apply_pushes_v(sh, n_dt, v0, v1, v2, r0, r1, r2);
            if (d.stream_hints) {
                __stcs(d.ph.r0 + i, r0);
                __stcs(d.ph.r1 + i, r1);
                __stcs(d.ph.r2 + i, r2);
            } else {
                d.ph.r0[i] = r0;
                d.ph.r1[i] = r1;
                d.ph.r2[i] = r2;
            }
        }
        // findContainingHydroCell, Src/mclib.c:469-597
        const bool skip = may_skip && (s_now < safe); // provably still inside its cell and the domain
        double t = default_t;
        bool have_t = true;
        bool inside = skip;
        if (!skip || verify) {
            if (may_skip) idx = d.ph.idx[i];
            double h0, h1, h2;
            coord_to_hydro(d.dims, d.geom, r0, r1, r2, h0, h1, h2);
            bool in_domain;
            if (!ndim3)
                in_domain = ((h1 < d.cells.dom[3]) && (h1 > d.cells.dom[2]) && (h0 < d.cells.dom[1]) && (h0 > d.cells.dom[0])) &&
                            (idx != -1);
            else
                in_domain = ((h2 < d.cells.dom[5]) && (h2 > d.cells.dom[4]) && (h1 < d.cells.dom[3]) && (h1 > d.cells.dom[2]) &&
                             (h0 < d.cells.dom[1]) && (h0 > d.cells.dom[0])) &&
                            (idx != -1);
            inside = false;
            if (in_domain) {
                int blk = (sw == 0) ? idx : 0;
#if defined(MCRAT_EXP_NOGATHER)
                bool inb = true; // ablation build: no cell-geometry gather (see MCRAT_EXP_NOCOMPUTE)
#else
                bool inb = in_cell(ndim3, d.cells, blk, h0, h1, h2);
#endif
                if (d.cs && blk == 0) { // Src/mclib.c:510-515
                    if ((d.ph.c0[i]) + (d.ph.c1[i]) + (d.ph.c2[i]) + (d.ph.c3[i]) == 0) inb = false;
                }
                if (sw == 1 || !inb) {
                    int pos = LOCAL_RELOC ? first + atomicAdd(&d.sh[s].reloc_n, 1) : atomicAdd(&d.gs->reloc_count[parity], 1);
                    if (LOCAL_RELOC && reloc_flag) *reloc_flag = 1; // cluster team: tells the event block to look at the list
                    d.reloc_slot[pos] = i;
                    d.reloc_h0[pos] = h0;
                    d.reloc_h1[pos] = h1;
                    d.reloc_h2[pos] = h2;
                    d.reloc_best[pos] = INT_MAX;
                    have_t = false; // finish completes this photon
                } else {
                    inside = true;
                    if (may_skip && !skip)
                        d.ph.safe[i] = safe_path(d.cells.geoA, d.cells.geoB, d.dom_dev, d.geom, ndim3, blk, h0, h1, h2, v0, v1, v2,
                                                 flags & F_MOVABLE, s_now);
                }
            } else {
                if (idx != -1) d.ph.idx[i] = -1; // Src/mclib.c:589-595
            }
            if (skip && !inside) raise_error(d.gs, MCRAT_B200_ERR_STATE, i, ERR_SITE_PASS_VERIFY); // verify mode: the bound was wrong
        }
        if (inside && FUSE_MFP) {
            // calcMeanFreePath, Src/mclib.c:657-687
            if (flags & F_RECALC) {
                if (skip) idx = d.ph.idx[i];
                CellState c = load_cell_state(d.cells, idx);
                int terr = 0;
                const FallbackRng fr = {d.k0, k1, (uint32_t)j, iter, d.replay};
                const double tau = optical_depth(d.dims, d.geom, d.tau_calc, d.table, c, r0, r1, d.ph.p1[i], d.ph.p2[i],
                                                 d.ph.p3[i], d.ph.c0[i], &terr, &fr);
                if (terr) raise_error(d.gs, MCRAT_B200_ERR_TABLE, i, ERR_SITE_PASS_TABLE);
                store_tau(d.ph, i, tau);
                ntau = -1.0 / tau;
                d.ph.flags[i] = flags & ~F_RECALC;
            }
#if defined(MCRAT_EXP_NOCOMPUTE)
            // ablation build (profiles/ncu_r01_summary.md, "pass kernel: where the time goes"): the memory
            // pattern alone, without Philox / log / divisions.  Never defined in the product build.
            t = ntau * (double)j;
#else
            double xi = philox_mfp_uniform(d.k0, k1, iter, (uint32_t)j);
            t = free_path_time_n(ntau, xi);
#endif
        }
        if (FUSE_MFP && have_t) {
            if (d.stream_hints)
                __stcs(d.ph.tts + i, t);
            else
                d.ph.tts[i] = t;
            if (lex_less(t, i, best_t, best_i)) {
                best_t = t;
                best_i = i;
            }
        }
    }
}

template <bool FUSE_MFP>
__global__ void __launch_bounds__(PASS_THREADS, MCRAT_PASS_MINB) pass_kernel(DevCtx d, int sw, int parity)
{
    const int s = blockIdx.x / d.blocks_per_shard;
    const int b = blockIdx.x - s * d.blocks_per_shard;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        d.gs->reloc_count[parity ^ 1] = 0;
        d.gs->scan_work = 0;
    }
    double best_t = DBL_MAX;
    int best_i = INT_MAX;
    if (!loop_stopped(*d.gs, d.sh[s]))
        pass_body<FUSE_MFP, false, PASS_THREADS>(d, d.sh[s], s, b, d.blocks_per_shard, sw, parity, best_t, best_i);
    if (FUSE_MFP) {
        block_argmin<PASS_THREADS>(best_t, best_i);
        if (threadIdx.x == 0) {
            d.bm_t[blockIdx.x] = best_t;
            d.bm_i[blockIdx.x] = best_i;
        }
    }
}

// push only: updatePhotonPosition called directly by the driver (Src/mcrat.c:841), and the
// materialisation of pending event pushes before a download
__global__ void __launch_bounds__(PASS_THREADS) flush_push_kernel(DevCtx d)
{
    for (int i = blockIdx.x * PASS_THREADS + threadIdx.x; i < d.cap; i += gridDim.x * PASS_THREADS) {
        const ShardState &sh = d.sh[shard_of(d, i)];
        const int n_dt = sh.n_dt;
        if (n_dt == 0) continue;
        unsigned char flags = d.ph.flags[i];
        if ((flags & F_MOVABLE) && i != sh.pushed_slot) {
            double r0 = d.ph.r0[i], r1 = d.ph.r1[i], r2 = d.ph.r2[i];
            apply_pushes_v(sh, n_dt, d.ph.v0[i], d.ph.v1[i], d.ph.v2[i], r0, r1, r2);
            d.ph.r0[i] = r0;
            d.ph.r1[i] = r1;
            d.ph.r2[i] = r2;
        }
    }
}

__global__ void clear_push_kernel(DevCtx d)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < d.nshards; s += gridDim.x * blockDim.x) {
        fold_path(d.sh[s], d.path_pad);
        d.sh[s].n_dt = 0;
        d.sh[s].pushed_slot = -1;
    }
}

__global__ void set_push_kernel(DevCtx d, double t)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < d.nshards; s += gridDim.x * blockDim.x) {
        fold_path(d.sh[s], d.path_pad);
        d.sh[s].dt_list[0] = t;
        d.sh[s].n_dt = 1;
        d.sh[s].pushed_slot = -1;
    }
}
