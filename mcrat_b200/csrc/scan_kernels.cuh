// scan_kernels.cuh -- K1 / K1b / K1c scans, finish, un-fused mean free path.
// Part of the single translation unit mcrat_b200.cu (included there, in this order); not a stand-alone header.
#pragma once

// ------------------------------------------------------------------------------------------
// K1: photon x cell containment scan.  Photons in registers (SCAN_P2 / SCAN_P3 per thread), cells streamed
// through a double-buffered shared-memory tile filled by TMA bulk copies.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// One photon-cell containment test, checkInBlock (Src/geometry.c:394-417) as |x - c| <= size / 2 per dimension.
// Written in PTX so that it compiles to the minimum the FP64 pipe can do bit-equivalently -- one DADD (d = x - c) and
// one DSETP (|d| <= h, |.| as a source modifier) per dimension, the DSETPs chained through their predicate operand --
// plus ONE integer instruction, the predicated minimum that keeps the lowest containing index (first match wins,
// Src/geometry.c:361-368).  From C++ the compiler emits three independent DSETP.GTU and combines them with an
// unconditional VIMNMX and two SELs per test: 9 issue slots per 6 FP64-pipe instructions instead of 7.
// NaN (padding cells, padding photons) compares false.
__device__ __forceinline__ void scan_test2(double x0, double x1, const double4 &a, int cell, int &best)
{
#if MCRAT_SCAN_PTX_TEST
    const double d0 = x0 - a.x, d1 = x1 - a.y;
    asm("{\n\t"
        ".reg .pred p;\n\t"
        ".reg .f64 e0, e1;\n\t"
        "abs.f64 e0, %1;\n\t"
        "abs.f64 e1, %2;\n\t"
        "setp.le.f64 p, e0, %3;\n\t"
        "setp.le.and.f64 p, e1, %4, p;\n\t"
        "@p min.s32 %0, %0, %5;\n\t"
        "}"
        : "+r"(best)
        : "d"(d0), "d"(d1), "d"(a.z), "d"(a.w), "r"(cell));
#else
    bool hit = (fabs(x0 - a.x) <= a.z) & (fabs(x1 - a.y) <= a.w);
    if (hit) best = min(best, cell);
#endif
}

__device__ __forceinline__ void scan_test3(double x0, double x1, double x2, const double4 &a, const double2 &b, int cell, int &best)
{
#if MCRAT_SCAN_PTX_TEST
    const double d0 = x0 - a.x, d1 = x1 - a.y, d2 = x2 - a.z;
    asm("{\n\t"
        ".reg .pred p;\n\t"
        ".reg .f64 e0, e1, e2;\n\t"
        "abs.f64 e0, %1;\n\t"
        "abs.f64 e1, %2;\n\t"
        "abs.f64 e2, %3;\n\t"
        "setp.le.f64 p, e0, %4;\n\t"
        "setp.le.and.f64 p, e1, %5, p;\n\t"
        "setp.le.and.f64 p, e2, %6, p;\n\t"
        "@p min.s32 %0, %0, %7;\n\t"
        "}"
        : "+r"(best)
        : "d"(d0), "d"(d1), "d"(d2), "d"(a.w), "d"(b.x), "d"(b.y), "r"(cell));
#else
    bool hit = (fabs(x0 - a.x) <= a.w) & (fabs(x1 - a.y) <= b.x) & (fabs(x2 - a.z) <= b.y);
    if (hit) best = min(best, cell);
#endif
}

// Work distribution: the launch is PERSISTENT -- as many CTAs as the device holds at once, each pulling work items
// (photon chunk pc, cell chunk cc) from a counter in global memory until none is left.  An item is SCAN_THREADS x
// SCAN_P photons against tiles_per_item tiles of SCAN_TILE cells.  A static grid of the same items left the last
// wave of CTAs 40-60 % empty (4785 CTAs over 888 resident slots = 5.39 waves at 10^5 photons: a tenth of the
// kernel's time with half of the SMs idle); with the counter the SMs run dry only during the last item of each CTA,
// and items are sized so that this is <= 2-3 % of the kernel (scan_grid).  Items are numbered cell-chunk-major, so
// the CTAs running at any moment stream the same few tiles out of L2.
template <int NDIM3>
__global__ void __launch_bounds__(SCAN_THREADS, NDIM3 ? MCRAT_SCAN_MINB3 : MCRAT_SCAN_MINB) scan_kernel(DevCtx d, int parity, int tiles_per_item)
{
    GlobalState &gs = *d.gs;
    if (gs.error != 0) return;
    constexpr int SCAN_P = NDIM3 ? SCAN_P3 : SCAN_P2;
    constexpr int PCHUNK = SCAN_THREADS * SCAN_P;
    const int count = gs.reloc_count[parity];
    if (count <= 0) return;
    const int pchunks = (count + PCHUNK - 1) / PCHUNK;
    const int ntiles_total = d.cells.n_padded / SCAN_TILE;
    const int cchunks = (ntiles_total + tiles_per_item - 1) / tiles_per_item;
    const unsigned total_items = (unsigned)pchunks * (unsigned)cchunks;

    extern __shared__ __align__(128) unsigned char scan_smem[];
    constexpr uint32_t BYTES_A = SCAN_TILE * sizeof(double4);
    constexpr uint32_t BYTES_B = NDIM3 ? SCAN_TILE * sizeof(double2) : 0;
    double4(*sA)[SCAN_TILE] = reinterpret_cast<double4(*)[SCAN_TILE]>(scan_smem);
    double2(*sB)[SCAN_TILE] = reinterpret_cast<double2(*)[SCAN_TILE]>(scan_smem + 2 * BYTES_A);
    uint64_t *bar = reinterpret_cast<uint64_t *>(scan_smem + 2 * BYTES_A + 2 * BYTES_B);
    __shared__ unsigned sh_item;

    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    unsigned g = 0; // tiles this CTA has consumed so far: stage = g & 1, mbarrier phase = (g >> 1) & 1

    for (;;) {
        __syncthreads(); // sh_item of the previous item has been read by everyone; both stages are free
        if (threadIdx.x == 0) sh_item = atomicAdd(&gs.scan_work, 1u);
        __syncthreads();
        const unsigned item = sh_item;
        if (item >= total_items) break;
        const int cc = (int)(item / (unsigned)pchunks), pc = (int)(item - (unsigned)cc * (unsigned)pchunks);
        const int pbase = pc * PCHUNK;
        const int tile0 = cc * tiles_per_item;
        const int ntiles = min(tiles_per_item, ntiles_total - tile0);

        if (threadIdx.x == 0) {
            const int s0 = g & 1;
            mbar_expect_tx(&bar[s0], BYTES_A + BYTES_B);
            tma_bulk_g2s(&sA[s0][0], d.cells.geoA + (size_t)tile0 * SCAN_TILE, BYTES_A, &bar[s0]);
            if (NDIM3) tma_bulk_g2s(&sB[s0][0], d.cells.geoB + (size_t)tile0 * SCAN_TILE, BYTES_B, &bar[s0]);
        }
        double x0[SCAN_P], x1[SCAN_P], x2[SCAN_P];
        int best[SCAN_P];
#pragma unroll
        for (int p = 0; p < SCAN_P; ++p) {
            int j = pbase + p * SCAN_THREADS + threadIdx.x;
            bool ok = j < count;
            x0[p] = ok ? d.reloc_h0[j] : qnan;
            x1[p] = ok ? d.reloc_h1[j] : qnan;
            x2[p] = (ok && NDIM3) ? d.reloc_h2[j] : qnan;
            best[p] = INT_MAX;
        }

        for (int t = 0; t < ntiles; ++t, ++g) {
            const int s = g & 1;
            if (threadIdx.x == 0 && t + 1 < ntiles) {
                mbar_expect_tx(&bar[s ^ 1], BYTES_A + BYTES_B);
                tma_bulk_g2s(&sA[s ^ 1][0], d.cells.geoA + (size_t)(tile0 + t + 1) * SCAN_TILE, BYTES_A, &bar[s ^ 1]);
                if (NDIM3)
                    tma_bulk_g2s(&sB[s ^ 1][0], d.cells.geoB + (size_t)(tile0 + t + 1) * SCAN_TILE, BYTES_B, &bar[s ^ 1]);
            }
            mbar_wait(&bar[s], (g >> 1) & 1u);
            const int cbase = (tile0 + t) * SCAN_TILE;
_Pragma(MCRAT_PRAGMA_STR(unroll MCRAT_SCAN_UNROLL))
            for (int c = 0; c < SCAN_TILE; ++c) {
                const double4 a = sA[s][c];
                if (!NDIM3) {
#pragma unroll
                    for (int p = 0; p < SCAN_P; ++p) scan_test2(x0[p], x1[p], a, cbase + c, best[p]);
                } else {
                    const double2 b = sB[s][c];
#pragma unroll
                    for (int p = 0; p < SCAN_P; ++p) scan_test3(x0[p], x1[p], x2[p], a, b, cbase + c, best[p]);
                }
            }
            __syncthreads(); // everyone is done with stage s before it is refilled two tiles later
        }
#pragma unroll
        for (int p = 0; p < SCAN_P; ++p) {
            int j = pbase + p * SCAN_THREADS + threadIdx.x;
            if (best[p] != INT_MAX && j < count) atomicMin(&d.reloc_best[j], best[p]);
        }
        if (threadIdx.x == 0 && cc == 0) {
            int nph = min(count - pbase, PCHUNK);
            atomicAdd((unsigned long long *)&d.gs->cell_evals, (unsigned long long)nph * (unsigned long long)d.cells.n);
        }
    }
}

// K1b: the same containment test, cell-parallel, for a short relocation list
template <int NDIM3>
__global__ void __launch_bounds__(256) scan_few_kernel(DevCtx d, int parity)
{
    const GlobalState &gs = *d.gs;
    if (gs.error != 0) return;
    const int count = gs.reloc_count[parity];
    if (count == 0) return;
    __shared__ double sx0[FEW_RMAX], sx1[FEW_RMAX], sx2[FEW_RMAX];
    __shared__ int sbest[FEW_RMAX];
    for (int base = 0; base < count; base += FEW_RMAX) {
        const int r = min(FEW_RMAX, count - base);
        __syncthreads();
        for (int j = threadIdx.x; j < r; j += blockDim.x) {
            sx0[j] = d.reloc_h0[base + j];
            sx1[j] = d.reloc_h1[base + j];
            sx2[j] = NDIM3 ? d.reloc_h2[base + j] : 0.0;
            sbest[j] = INT_MAX;
        }
        __syncthreads();
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < d.cells.n; c += gridDim.x * blockDim.x) {
            const double4 a = d.cells.geoA[c];
            double2 b = make_double2(0, 0);
            if (NDIM3) b = d.cells.geoB[c];
            for (int j = 0; j < r; ++j) {
                bool hit;
                if (!NDIM3)
                    hit = (fabs(sx0[j] - a.x) <= a.z) & (fabs(sx1[j] - a.y) <= a.w);
                else
                    hit = (fabs(sx0[j] - a.x) <= a.w) & (fabs(sx1[j] - a.y) <= b.x) & (fabs(sx2[j] - a.z) <= b.y);
                if (hit) atomicMin(&sbest[j], c);
            }
        }
        __syncthreads();
        for (int j = threadIdx.x; j < r; j += blockDim.x)
            if (sbest[j] != INT_MAX) atomicMin(&d.reloc_best[base + j], sbest[j]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd((unsigned long long *)&d.gs->cell_evals, (unsigned long long)count * (unsigned long long)d.cells.n);
}

// ------------------------------------------------------------------------------------------
// K1c: the same first-match search through a two-level bounding-box index (opt-in,
// mcrat_b200_config.scan_index).  The reference carries a disabled uniform-bucket accelerator
// (Src/geometry.c:423-676, switched off at Src/mcrat_io.c:1985); this index is built over the
// cells *in array order*, so walking boxes and cells in ascending index and stopping at the
// first hit returns exactly the cell findContainingBlock returns (lowest containing index).
// A box is padded outward by a few ulps so that every cell test that can succeed is reached.
// ------------------------------------------------------------------------------------------
__global__ void build_box1_kernel(int ndim3, int n, const double4 *geoA, const double2 *geoB, double *box1, int nbox1)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nbox1; b += gridDim.x * blockDim.x) {
        double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (int c = b * BOX_T; c < min(n, b * BOX_T + BOX_T); ++c) {
            double4 a = geoA[c];
            double cc[3], hh[3];
            if (!ndim3) {
                cc[0] = a.x; cc[1] = a.y; cc[2] = 0; hh[0] = a.z; hh[1] = a.w; hh[2] = 0;
            } else {
                double2 q = geoB[c];
                cc[0] = a.x; cc[1] = a.y; cc[2] = a.z; hh[0] = a.w; hh[1] = q.x; hh[2] = q.y;
            }
            for (int k = 0; k < 3; ++k) {
                double pad = 8.0 * 2.220446049250313e-16 * (fabs(cc[k]) + fabs(hh[k]));
                lo[k] = fmin(lo[k], cc[k] - hh[k] - pad);
                hi[k] = fmax(hi[k], cc[k] + hh[k] + pad);
            }
        }
        for (int k = 0; k < 3; ++k) {
            box1[6 * b + 2 * k] = lo[k];
            box1[6 * b + 2 * k + 1] = hi[k];
        }
    }
}

__global__ void build_box2_kernel(const double *box1, int nbox1, double *box2, int nbox2)
{
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nbox2; b += gridDim.x * blockDim.x) {
        double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (int c = b * BOX_T; c < min(nbox1, b * BOX_T + BOX_T); ++c)
            for (int k = 0; k < 3; ++k) {
                lo[k] = fmin(lo[k], box1[6 * c + 2 * k]);
                hi[k] = fmax(hi[k], box1[6 * c + 2 * k + 1]);
            }
        for (int k = 0; k < 3; ++k) {
            box2[6 * b + 2 * k] = lo[k];
            box2[6 * b + 2 * k + 1] = hi[k];
        }
    }
}

__device__ __forceinline__ bool in_box(int ndim3, const double *bx, double x0, double x1, double x2)
{
    bool in = (x0 >= bx[0]) & (x0 <= bx[1]) & (x1 >= bx[2]) & (x1 <= bx[3]);
    if (ndim3) in = in & (x2 >= bx[4]) & (x2 <= bx[5]);
    return in;
}

// One warp locates one photon.  Level-2 boxes are all tested first (independent loads, 32 per
// round, hits kept as one bit per round and lane), then the hits are descended in ascending order:
// 32 level-1 boxes per level-2 box and 32 cells per level-1 box, one per lane; the lowest lane of
// the first ballot with a containing cell is the lowest containing index = findContainingBlock's answer.
__device__ __forceinline__ int warp_locate_indexed(const DevCtx &d, const double x0, const double x1, const double x2,
                                                   long long &cells_tested, long long &boxes_tested)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int ndim3 = (d.dims == D_THREE);
    const CellCols &c = d.cells;
    int best = INT_MAX;
    for (int chunk = 0; chunk < c.nbox2 && best == INT_MAX; chunk += 2048) {
        const int rounds = min(64, (c.nbox2 - chunk + 31) / 32);
        unsigned long long mine = 0;
#pragma unroll 4
        for (int r = 0; r < rounds; ++r) {
            const int b2 = chunk + r * 32 + lane;
            if (b2 < c.nbox2 && in_box(ndim3, c.box2 + 6 * b2, x0, x1, x2)) mine |= 1ull << r;
        }
        boxes_tested += min(c.nbox2 - chunk, 2048);
        for (int r = 0; r < rounds && best == INT_MAX; ++r) {
            unsigned m2 = __ballot_sync(full, (mine >> r) & 1ull);
            while (m2 && best == INT_MAX) {
                const int B2 = chunk + r * 32 + (__ffs(m2) - 1);
                m2 &= m2 - 1;
                const int b1 = B2 * BOX_T + lane;
                unsigned m1 = __ballot_sync(full, b1 < c.nbox1 && in_box(ndim3, c.box1 + 6 * b1, x0, x1, x2));
                boxes_tested += min(BOX_T, c.nbox1 - B2 * BOX_T);
                while (m1 && best == INT_MAX) {
                    const int B1 = B2 * BOX_T + (__ffs(m1) - 1);
                    m1 &= m1 - 1;
                    const int cell = B1 * BOX_T + lane;
                    const unsigned mc = __ballot_sync(full, cell < c.n && in_cell(ndim3, c, cell, x0, x1, x2));
                    cells_tested += min(BOX_T, c.n - B1 * BOX_T);
                    if (mc) best = B1 * BOX_T + (__ffs(mc) - 1);
                }
            }
        }
    }
    return best;
}

__global__ void __launch_bounds__(128) scan_index_kernel(DevCtx d, int parity)
{
    const GlobalState &gs = *d.gs;
    if (gs.error != 0) return;
    const int count = gs.reloc_count[parity];
    long long cells_tested = 0, boxes_tested = 0;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int j = warp; j < count; j += nwarps) {
        const int best = warp_locate_indexed(d, d.reloc_h0[j], d.reloc_h1[j], d.reloc_h2[j], cells_tested, boxes_tested);
        if ((threadIdx.x & 31) == 0) d.reloc_best[j] = best;
    }
    if ((threadIdx.x & 31) == 0 && (cells_tested | boxes_tested)) {
        atomicAdd((unsigned long long *)&d.gs->cell_evals, (unsigned long long)cells_tested);
        atomicAdd((unsigned long long *)&d.gs->box_evals, (unsigned long long)boxes_tested);
    }
}

// ------------------------------------------------------------------------------------------
// finish: relocated photons get their new cell, comoving 4-momentum and optical depth
// (Src/mclib.c:536-584); in the fused loop also their free-path draw
// ------------------------------------------------------------------------------------------
constexpr int FIN_THREADS = 128;

// one relocated photon: new cell (or -1), comoving 4-momentum, optical depth, free-path draw
template <bool FUSE_MFP>
__device__ __forceinline__ bool finish_one(DevCtx &d, ShardState &sh, const int s, const int i, const int b, const int sw)
{
    double t = 1e12 / C_LIGHT;
    bool missing = false;
    if (b == INT_MAX) {
        d.ph.idx[i] = -1; // Src/mclib.c:536, 581-584
        d.ph.safe[i] = 0;
        missing = true;
        const int k = atomicAdd(&d.gs->nf_logged, 1);
        if (k < NF_LOG_CAP) { // for the rank's log line, Src/geometry.c:373-388
            double h0, h1, h2;
            coord_to_hydro(d.dims, d.geom, d.ph.r0[i], d.ph.r1[i], d.ph.r2[i], h0, h1, h2);
            d.gs->nf_slot[k] = i;
            d.gs->nf_h[3 * k] = h0;
            d.gs->nf_h[3 * k + 1] = h1;
            d.gs->nf_h[3 * k + 2] = h2;
        }
    } else {
        d.ph.idx[i] = b;
        d.ph.safe[i] = 0; // the next pass re-checks the new cell and sets the threshold
        double p[4] = {d.ph.p0[i], d.ph.p1[i], d.ph.p2[i], d.ph.p3[i]};
        double r0 = d.ph.r0[i], r1 = d.ph.r1[i];
        CellState c = load_cell_state(d.cells, b);
        double fb[3], pc[4];
        fluid_beta_of(d, c, r0, r1, fb);
        lorentz_boost(fb, p, pc, true);
        d.ph.c0[i] = pc[0];
        d.ph.c1[i] = pc[1];
        d.ph.c2[i] = pc[2];
        d.ph.c3[i] = pc[3];
        int terr = 0;
        const FallbackRng fr = {d.k0, d.k1 ^ (d.shard_base + (uint32_t)s), (uint32_t)(i - sh.first), sh.iter, d.replay};
        double tau = optical_depth(d.dims, d.geom, d.tau_calc, d.table, c, r0, r1, p[1], p[2], p[3], pc[0], &terr, &fr);
        if (terr) raise_error(d.gs, MCRAT_B200_ERR_TABLE, i, ERR_SITE_FINISH_TABLE);
        store_tau(d.ph, i, tau);
        d.ph.flags[i] = d.ph.flags[i] & ~F_RECALC;
        if (sw == 0) atomicAdd((unsigned long long *)&sh.reloc_total, 1ull); // Src/mclib.c:579, 608-611
        if (FUSE_MFP) {
            const uint32_t k1 = d.k1 ^ (d.shard_base + (uint32_t)s);
            double xi = philox_mfp_uniform(d.k0, k1, sh.iter, (uint32_t)(i - sh.first));
            t = free_path_time(tau, xi);
        }
    }
    if (FUSE_MFP) d.ph.tts[i] = t;
    return missing;
}

template <bool FUSE_MFP>
__global__ void __launch_bounds__(FIN_THREADS) finish_kernel(DevCtx d, int sw, int parity)
{
    const GlobalState &gs = *d.gs;
    if (gs.error != 0) return;
    const int count = gs.reloc_count[parity];
    int missing = 0;
    long long ref_evals = 0; // the reference's loop stops at the first hit (Src/geometry.c:361-368)
    for (int j = blockIdx.x * FIN_THREADS + threadIdx.x; j < count; j += gridDim.x * FIN_THREADS)
    {
        const int i = d.reloc_slot[j], s = shard_of(d, i);
        const int best = d.reloc_best[j];
        if (sw == 1) ref_evals += (best == INT_MAX) ? d.cells.n : best + 1;
        if (finish_one<FUSE_MFP>(d, d.sh[s], s, i, best, sw)) missing++;
    }
    if (missing) atomicAdd(&d.gs->not_found, missing);
    if (sw == 1) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ref_evals += __shfl_xor_sync(0xffffffffu, ref_evals, off);
        if ((threadIdx.x & 31) == 0 && ref_evals) atomicAdd((unsigned long long *)&d.gs->ref_equiv_evals, (unsigned long long)ref_evals);
    }
}

// ------------------------------------------------------------------------------------------
// unfused calcMeanFreePath (step API and replay harness; single shard), Src/mclib.c:617-714
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mfp_count_kernel(DevCtx d)
{
    int i = blockIdx.x * 256 + threadIdx.x;
    int in = (i < d.cap) && (d.ph.idx[i] != -1);
    int c = __syncthreads_count(in);
    if (threadIdx.x == 0) d.prefix_block[blockIdx.x] = c;
}

__global__ void mfp_scan_kernel(DevCtx d, int nblocks)
{
    // single thread: exclusive scan of per-block counts (replay harness only; small lists)
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long run = 0;
        for (int b = 0; b < nblocks; ++b) {
            int c = d.prefix_block[b];
            d.prefix_block[b] = (int)run;
            run += (unsigned long long)c;
        }
        d.gs->replay_base = d.gs->replay_cursor;
        d.gs->replay_cursor += run;
        if (d.gs->replay_cursor > d.gs->replay_n) raise_error(d.gs, MCRAT_B200_ERR_REPLAY, -1, ERR_SITE_MFP_REPLAY);
    }
}

__global__ void __launch_bounds__(256) mfp_kernel(DevCtx d, int write_blockmin)
{
    const GlobalState &gs = *d.gs;
    const ShardState &sh = d.sh[0];
    if (gs.error != 0) return;
    __shared__ int warp_off[8];
    const int i = blockIdx.x * 256 + threadIdx.x;
    const bool valid = i < d.cap;
    const int idx = valid ? d.ph.idx[i] : -1;
    const bool in = valid && idx != -1;
    double t = 1e12 / C_LIGHT;
    // rank of this photon among the in-domain photons of the block (stream order = slot order)
    unsigned ball = __ballot_sync(0xffffffffu, in);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_off[w] = __popc(ball);
    __syncthreads();
    int off = 0;
    for (int k = 0; k < w; ++k) off += warp_off[k];
    off += __popc(ball & ((1u << lane) - 1u));
    if (in) {
        unsigned char flags = d.ph.flags[i];
        double tau;
        if (flags & F_RECALC) {
            CellState c = load_cell_state(d.cells, idx);
            int terr = 0;
            const FallbackRng fr = {d.k0, d.k1 ^ d.shard_base, (uint32_t)i, sh.iter, d.replay};
            tau = optical_depth(d.dims, d.geom, d.tau_calc, d.table, c, d.ph.r0[i], d.ph.r1[i], d.ph.p1[i], d.ph.p2[i],
                                d.ph.p3[i], d.ph.c0[i], &terr, &fr);
            if (terr) raise_error(d.gs, MCRAT_B200_ERR_TABLE, i, ERR_SITE_MFP_TABLE);
            store_tau(d.ph, i, tau);
            d.ph.flags[i] = flags & ~F_RECALC;
        } else {
            tau = d.ph.tau[i];
        }
        double xi;
        if (d.replay)
            xi = d.replay_buf[gs.replay_base + (unsigned long long)d.prefix_block[blockIdx.x] + (unsigned long long)off];
        else
            xi = philox_mfp_uniform(d.k0, d.k1 ^ d.shard_base, sh.iter, (uint32_t)i);
        t = free_path_time(tau, xi);
    }
    int bi = valid ? i : INT_MAX;
    double bt = valid ? t : DBL_MAX;
    if (valid) d.ph.tts[i] = t;
    block_argmin<256>(bt, bi);
    if (threadIdx.x == 0 && write_blockmin) {
        d.bm_t[blockIdx.x] = bt;
        d.bm_i[blockIdx.x] = bi;
    }
}

// head of the time-ordered list for lists with more than BLOCKMIN_CAP*256 slots in the unfused path
__global__ void __launch_bounds__(256) argmin_all_kernel(DevCtx d)
{
    double bt = DBL_MAX;
    int bi = INT_MAX;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < d.cap; i += gridDim.x * 256) {
        double t = d.ph.tts[i];
        if (lex_less(t, i, bt, bi)) {
            bt = t;
            bi = i;
        }
    }
    block_argmin<256>(bt, bi);
    if (threadIdx.x == 0) {
        d.bm_t[blockIdx.x] = bt;
        d.bm_i[blockIdx.x] = bi;
    }
}
