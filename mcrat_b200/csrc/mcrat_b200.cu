// mcrat_b200.cu -- kernels and C ABI of the B200-native MCRaT hot path (sm_100a, FP64).
//
// Kernel map (DESIGN.md has the data layout and the roofline of each):
//   K1  scan_kernel        photon x cell containment scan, cells staged in shared memory by
//                          TMA bulk copies (cp.async.bulk + mbarrier), photons register-tiled.
//                          Replaces findContainingBlock's O(N_cells) loop,
//                          Src/geometry.c:350-391 -> :394-417, for a whole relocation list.
//   K1b scan_few_kernel    the same test, cell-parallel, for the handful of photons that
//                          leave their cell in a steady-state iteration.
//   K4+K2 pass_kernel      fused push (Src/mclib.c:1054-1100) + domain / in-cell re-check
//                          (Src/mclib.c:469-597) + free-path draw (Src/mclib.c:646-692) +
//                          warp-shuffle / block arg-min over time_to_scatter (replaces the
//                          qsort of Src/mclib.c:702-710: only the head of the order is used).
//       finish_kernel      re-boost + optical depth of relocated photons (Src/mclib.c:538-580).
//   K3  event_kernel       global arg-min + photonEvent (Src/mclib.c:1107-1356): fluid-frame
//                          boost, electron sampling, polarised Klein-Nishina scatter, Stokes
//                          rotations, boost back; plus the driver's bookkeeping of
//                          Src/mcrat.c:777-846.
//   K5  cs_absorb_kernel   phAbsCyclosynch, Src/mc_cyclosynch.c:1571-1644.
//   K7  hot_table_kernel   thermal Klein-Nishina cross-section table, Src/hot_x_section.c:82-206.
//   K8  rebin_*_kernel     rebinCyclosynchCompPhotons, Src/mc_cyclosynch.c:244-710.
//       frame_loop_kernel  the whole while-loop of Src/mcrat.c:761-851 in one cooperative launch.
// Files (one translation unit, included below in this order): state.cuh (loop state, columns, DevCtx),
// pass_kernels.cuh (helpers, AoS <-> SoA, K4+K2), scan_kernels.cuh (K1 / K1b / K1c, finish, un-fused free path),
// event.cuh (K3), frame_loop.cuh (persistent loop), aux_kernels.cuh (K5, K7, statistics, K8, peak probes);
// device_math.cuh holds the physics in the reference's operation order.  This file: the host side and the C ABI.
// No CPU fallback exists: every entry point fails with MCRAT_B200_ERR_CUDA without a device.
#include "../../include/mcrat_b200.h"
#include "device_math.cuh"

#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>
#include <string>
#include <vector>

using namespace mcrat;

static_assert(sizeof(mcrat_photon) == 176, "struct photon layout (Src/mcrat.h:142-171) must be 176 bytes");

#define API extern "C" __attribute__((visibility("default")))

// the kernels, in dependency order (one translation unit: everything is inlined into the __global__ functions)
#include "state.cuh"
#include "pass_kernels.cuh"
#include "scan_kernels.cuh"
#include "event.cuh"
#include "frame_loop.cuh"
#include "aux_kernels.cuh"
#include "cs_emit.cuh"

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static thread_local std::string g_create_error;

struct mcrat_b200_ctx {
    mcrat_b200_config cfg;
    DevCtx d;
    cudaStream_t stream;
    bool own_stream;
    cudaStream_t stream2;          // second stream of the interleaved streamed loop (the other half of the sub-shards)
    cudaEvent_t ev_pass[2], ev_join; // pass of half h has finished; stream2 has drained
    int num_sms;
    std::string err;
    // device allocations
    std::vector<void *> ph_allocs, cell_allocs, misc_allocs;
    mcrat_photon *aos_dev;
    size_t aos_cap;
    int ph_cap_alloc;
    bool have_hydro, have_photons;
    int pass_parity;
    int last_nb_mfp;
    int want_shards;    // sub-shards requested for the next set_photons
    int loop_mode;      // MCRAT_B200_LOOP_AUTO / _STREAMED / _PERSISTENT
    double cs_rebin_e_perc, cs_rebin_ang, cs_rebin_ang_phi; // CYCLOSYNCHROTRON_REBIN_E_PERC / _ANG / _ANG_PHI, Src/mcrat.h:308-322
    int occ_scan[2];              // resident CTAs per SM of scan_kernel<0> / <1>
    int occ_loop256, occ_loop128; // resident blocks per SM of frame_loop_kernel<256>, frame_loop_solo_kernel<256> / <128>
    int cluster_state;            // 0 untried, 1 the cluster team kernel launches on this device, -1 it does not (cooperative team instead)
    long long launches; // kernels launched through this context
    double hydro_fps;   // of the frame last uploaded (calcCyclosynchRLimits, Src/mc_cyclosynch.c:1206-1207)
    int hydro_scatt_frame, hydro_inj_frame;
    unsigned int emit_epoch; // photonEmitCyclosynch (all cells) calls so far: part of the key of the emission streams
    GlobalState *gs_host;          // pinned
    std::vector<ShardState> sh_host;
    double *replay_dev;
    size_t replay_cap;
    double *table_dev;
    StatPartial *stat_dev;
    // profiling
    cudaEvent_t ev0, ev1;
    mcrat_b200_kernel_times times;
};

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                          \
            return MCRAT_B200_ERR_CUDA;                                                              \
        }                                                                                            \
    } while (0)

static int fail(mcrat_b200_ctx *ctx, int code, const char *msg)
{
    ctx->err = msg;
    return code;
}

template <typename T>
static cudaError_t dev_alloc(std::vector<void *> &pool, T **p, size_t n)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, (n ? n : 1) * sizeof(T));
    if (e == cudaSuccess) {
        pool.push_back(q);
        *p = (T *)q;
    }
    return e;
}

static void free_pool(std::vector<void *> &pool)
{
    for (void *p : pool) cudaFree(p);
    pool.clear();
}

enum { KC_SCAN = 0, KC_PASS = 1, KC_EVENT = 2, KC_OTHER = 3 };

struct Timed {
    mcrat_b200_ctx *ctx;
    int cls;
    bool on;
    Timed(mcrat_b200_ctx *c, int k) : ctx(c), cls(k), on(c->cfg.profile != 0)
    {
        if (on) cudaEventRecord(ctx->ev0, ctx->stream);
    }
    ~Timed()
    {
        if (!on) return;
        cudaEventRecord(ctx->ev1, ctx->stream);
        cudaEventSynchronize(ctx->ev1);
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        mcrat_b200_kernel_times &t = ctx->times;
        if (cls == KC_SCAN) { t.scan_ms += ms; t.scan_launches++; }
        else if (cls == KC_PASS) { t.pass_ms += ms; t.pass_launches++; }
        else if (cls == KC_EVENT) { t.event_ms += ms; t.event_launches++; }
        else { t.other_ms += ms; t.other_launches++; }
    }
};

API int mcrat_b200_abi_version(void) { return MCRAT_B200_ABI_VERSION; }

API int mcrat_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

API const char *mcrat_b200_last_error(const mcrat_b200_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

static size_t scan_smem_bytes(int ndim3)
{
    return 2 * SCAN_TILE * sizeof(double4) + (ndim3 ? 2 * SCAN_TILE * sizeof(double2) : 0) + 2 * sizeof(uint64_t);
}

API int mcrat_b200_create(const mcrat_b200_config *cfg, mcrat_b200_ctx **out)
{
    if (!cfg || !out) {
        g_create_error = "null argument";
        return MCRAT_B200_ERR_ARG;
    }
    if (cfg->abi_version != MCRAT_B200_ABI_VERSION) {
        g_create_error = "ABI version mismatch";
        return MCRAT_B200_ERR_ARG;
    }
    if (cfg->dimensions < 0 || cfg->dimensions > 2 || cfg->geometry < 0 || cfg->geometry > 3) {
        g_create_error = "bad DIMENSIONS / GEOMETRY";
        return MCRAT_B200_ERR_ARG;
    }
    // the combinations the reference supports, Src/geometry.c:20-58
    if (cfg->dimensions != MCRAT_THREE && cfg->geometry == MCRAT_POLAR) {
        g_create_error = "POLAR geometry exists only in 3-D";
        return MCRAT_B200_ERR_ARG;
    }
    if (cfg->dimensions == MCRAT_THREE && cfg->geometry == MCRAT_CYLINDRICAL) {
        g_create_error = "CYLINDRICAL geometry exists only in 2-D / 2.5-D";
        return MCRAT_B200_ERR_ARG;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)";
        return MCRAT_B200_ERR_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) {
        g_create_error = "device ordinal out of range";
        return MCRAT_B200_ERR_ARG;
    }
    mcrat_b200_ctx *ctx = new mcrat_b200_ctx();
    ctx->cfg = *cfg;
    memset(&ctx->d, 0, sizeof(ctx->d));
    memset(&ctx->times, 0, sizeof(ctx->times));
    ctx->aos_dev = nullptr;
    ctx->aos_cap = 0;
    ctx->ph_cap_alloc = 0;
    ctx->have_hydro = ctx->have_photons = false;
    ctx->pass_parity = 0;
    ctx->last_nb_mfp = 0;
    ctx->want_shards = 1;
    ctx->loop_mode = MCRAT_B200_LOOP_AUTO;
    ctx->cs_rebin_e_perc = 0.1;
    ctx->cs_rebin_ang = 0.5;
    ctx->cs_rebin_ang_phi = 10;
    ctx->occ_loop256 = ctx->occ_loop128 = 0;
    ctx->occ_scan[0] = ctx->occ_scan[1] = 0;
    ctx->launches = 0;
    ctx->hydro_fps = 0;
    ctx->hydro_scatt_frame = ctx->hydro_inj_frame = 0;
    ctx->emit_epoch = 0;
    ctx->replay_dev = nullptr;
    ctx->replay_cap = 0;
    ctx->table_dev = nullptr;
    // the cluster team kernel is correct (bit-identical photons) but no faster than the cooperative one (profiles/ncu_r02_summary.md):
    // it runs only on request
    ctx->cluster_state = getenv("MCRAT_B200_CLUSTER_TEAM") ? 0 : -1;
    ctx->stat_dev = nullptr;
    ctx->gs_host = nullptr;
    auto bail = [&](cudaError_t err, const char *what) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
        delete ctx;
        return MCRAT_B200_ERR_CUDA;
    };
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    ctx->num_sms = prop.multiProcessorCount;
    if (cfg->stream) {
        ctx->stream = (cudaStream_t)cfg->stream;
        ctx->own_stream = false;
    } else {
        if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
        ctx->own_stream = true;
    }
    if ((e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    for (int k = 0; k < 2; ++k)
        if ((e = cudaEventCreateWithFlags(&ctx->ev_pass[k], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaMallocHost((void **)&ctx->gs_host, sizeof(GlobalState))) != cudaSuccess) return bail(e, "cudaMallocHost");
    DevCtx &d = ctx->d;
    d.dims = cfg->dimensions;
    d.geom = cfg->geometry;
    d.stokes = cfg->stokes_switch ? 1 : 0;
    d.tau_calc = cfg->tau_calculation == MCRAT_TABLE ? TAU_TABLE : TAU_DIRECT;
    d.cs = cfg->cyclosynch_switch ? 1 : 0;
    d.b_calc = cfg->b_field_calc;
    d.epsilon_b = cfg->epsilon_b;
    d.k0 = (uint32_t)cfg->seed ^ 0x4D435261u;
    d.k1 = (uint32_t)(cfg->seed >> 32);
    d.shard_base = cfg->shard;
    d.replay = cfg->rng_mode == MCRAT_RNG_REPLAY ? 1 : 0;
    d.recheck_skip = getenv("MCRAT_B200_NO_RECHECK_SKIP") ? 0 : 1;
    d.path_pad = 0;
    d.mj_rounds = getenv("MCRAT_B200_MJ_ROUNDS") ? atoi(getenv("MCRAT_B200_MJ_ROUNDS")) : (1 << 13);
    if (d.mj_rounds < 0) d.mj_rounds = 0;
    d.nshards = 1;
    d.shard_size = 1;
    d.blocks_per_shard = 1;
    if ((e = dev_alloc(ctx->misc_allocs, &d.gs, 1)) != cudaSuccess) return bail(e, "cudaMalloc");
    {
        double *dom = nullptr;
        if ((e = dev_alloc(ctx->misc_allocs, &dom, 6)) != cudaSuccess) return bail(e, "cudaMalloc");
        d.dom_dev = dom;
    }
    memset(ctx->gs_host, 0, sizeof(GlobalState));
    ctx->gs_host->cs_max_photons = INT_MAX;
    if ((e = cudaMemcpyAsync(d.gs, ctx->gs_host, sizeof(GlobalState), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess)
        return bail(e, "cudaMemcpy");
    if ((e = dev_alloc(ctx->misc_allocs, &d.sh, MAX_SHARDS)) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMemsetAsync(d.sh, 0, sizeof(ShardState) * MAX_SHARDS, ctx->stream)) != cudaSuccess) return bail(e, "cudaMemset");
    ctx->sh_host.assign(MAX_SHARDS, ShardState());
    memset(ctx->sh_host.data(), 0, sizeof(ShardState) * MAX_SHARDS);
    if ((e = dev_alloc(ctx->misc_allocs, &d.bm_t, BLOCKMIN_CAP)) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = dev_alloc(ctx->misc_allocs, &d.bm_i, BLOCKMIN_CAP)) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = dev_alloc(ctx->misc_allocs, &d.bm_idx, BLOCKMIN_CAP)) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = dev_alloc(ctx->misc_allocs, &d.bm_temp, BLOCKMIN_CAP)) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = dev_alloc(ctx->misc_allocs, &ctx->stat_dev, 1024)) != cudaSuccess) return bail(e, "cudaMalloc");
    // interpolation grids, Src/hot_x_section.c:470-480
    {
        std::vector<double> grids(N_PH_E + 1 + N_T + 1);
        double dt = (LOG_T_MAX - LOG_T_MIN) / N_T, dph_e = (LOG_PH_E_MAX - LOG_PH_E_MIN) / N_PH_E;
        for (int i = 0; i <= N_PH_E; i++) grids[i] = LOG_PH_E_MIN + i * dph_e;
        for (int i = 0; i <= N_T; i++) grids[N_PH_E + 1 + i] = LOG_T_MIN + i * dt;
        double *g = nullptr;
        if ((e = dev_alloc(ctx->misc_allocs, &g, grids.size() + (size_t)(N_PH_E + 1) * (N_T + 1))) != cudaSuccess)
            return bail(e, "cudaMalloc");
        if ((e = cudaMemcpyAsync(g, grids.data(), grids.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess)
            return bail(e, "cudaMemcpy");
        if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bail(e, "cudaStreamSynchronize");
        d.table.xa = g;
        d.table.ya = g + N_PH_E + 1;
        d.table.za = g + N_PH_E + 1 + N_T + 1;
        ctx->table_dev = g + N_PH_E + 1 + N_T + 1;
    }
    if ((e = cudaFuncSetAttribute(frame_loop_cluster_kernel<256>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)) != cudaSuccess) {
        (void)cudaGetLastError();
        ctx->cluster_state = -1; // no 16-block clusters here: the cooperative team kernel does the job
    }
    // The two grids of the persistent stream share SMs.  An SM runs with ONE split of its 256 KB between L1 and shared memory;
    // blocks of a kernel that prefers another split wait until the SM has drained (measured: with different preferences the
    // event blocks took 128 SMs for themselves and the pass ran on the remaining 20).  Both kernels therefore ask for the same
    // small carve-out: 32 KB of shared memory covers 5 pass blocks + 1 event block, and the pass keeps its spills in L1.
    {
        int carve = 14; // per cent of 228 KB, rounded up by the driver to the 32 KB configuration
        if (const char *c = getenv("MCRAT_B200_STREAM_CARVEOUT")) carve = atoi(c);
        if (carve != -2) { // -2: leave the kernels' attributes alone (experiments)
            if ((e = cudaFuncSetAttribute(frame_stream_pass_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve)) != cudaSuccess)
                return bail(e, "cudaFuncSetAttribute");
            if ((e = cudaFuncSetAttribute(frame_stream_event_kernel<EVT_THREADS_MANY>, cudaFuncAttributePreferredSharedMemoryCarveout, carve)) != cudaSuccess)
                return bail(e, "cudaFuncSetAttribute");
        }
    }
    if ((e = cudaFuncSetAttribute(scan_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem_bytes(0))) != cudaSuccess)
        return bail(e, "cudaFuncSetAttribute");
    if ((e = cudaFuncSetAttribute(scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem_bytes(1))) != cudaSuccess)
        return bail(e, "cudaFuncSetAttribute");
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_scan[0], scan_kernel<0>, SCAN_THREADS, scan_smem_bytes(0))) != cudaSuccess)
        return bail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_scan[1], scan_kernel<1>, SCAN_THREADS, scan_smem_bytes(1))) != cudaSuccess)
        return bail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_loop256, frame_loop_kernel<256>, 256, 0)) != cudaSuccess)
        return bail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_loop128, frame_loop_solo_kernel<128>, 128, 0)) != cudaSuccess)
        return bail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    {
        int occ_solo256 = 0;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_solo256, frame_loop_solo_kernel<256>, 256, 0)) != cudaSuccess)
            return bail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
        if (occ_solo256 < ctx->occ_loop256) ctx->occ_loop256 = occ_solo256;
    }
    *out = ctx;
    return MCRAT_B200_OK;
}

API void mcrat_b200_destroy(mcrat_b200_ctx *ctx)
{
    if (!ctx) return;
    cudaStreamSynchronize(ctx->stream);
    free_pool(ctx->ph_allocs);
    free_pool(ctx->cell_allocs);
    free_pool(ctx->misc_allocs);
    if (ctx->aos_dev) cudaFree(ctx->aos_dev);
    if (ctx->replay_dev) cudaFree(ctx->replay_dev);
    if (ctx->gs_host) cudaFreeHost(ctx->gs_host);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamSynchronize(ctx->stream2);
    cudaEventDestroy(ctx->ev_pass[0]);
    cudaEventDestroy(ctx->ev_pass[1]);
    cudaEventDestroy(ctx->ev_join);
    cudaStreamDestroy(ctx->stream2);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

API int mcrat_b200_synchronize(mcrat_b200_ctx *ctx)
{
    CK(cudaStreamSynchronize(ctx->stream));
    return MCRAT_B200_OK;
}

static int fetch_global(mcrat_b200_ctx *ctx)
{
    CK(cudaMemcpyAsync(ctx->gs_host, ctx->d.gs, sizeof(GlobalState), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MCRAT_B200_OK;
}

static int fetch_state(mcrat_b200_ctx *ctx)
{
    CK(cudaMemcpyAsync(ctx->gs_host, ctx->d.gs, sizeof(GlobalState), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->sh_host.data(), ctx->d.sh, sizeof(ShardState) * ctx->d.nshards, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MCRAT_B200_OK;
}

static int check_launch(mcrat_b200_ctx *ctx, const char *what, int n = 1)
{
    ctx->launches += n;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
        return MCRAT_B200_ERR_CUDA;
    }
    return MCRAT_B200_OK;
}

static int grid_for(mcrat_b200_ctx *ctx, int n, int threads, int per_sm)
{
    int g = (n + threads - 1) / threads;
    int cap = ctx->num_sms * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return g;
}

API int mcrat_b200_set_hydro(mcrat_b200_ctx *ctx, int n, const double *const *fields, const double *domains, double fps,
                             int scatt_frame_number, int inj_frame_number)
{
    if (!ctx || !fields || !domains || n < 0) return ctx ? fail(ctx, MCRAT_B200_ERR_ARG, "set_hydro: bad argument") : MCRAT_B200_ERR_ARG;
    ctx->hydro_fps = fps;
    ctx->hydro_scatt_frame = scatt_frame_number;
    ctx->hydro_inj_frame = inj_frame_number;
    CK(cudaSetDevice(ctx->cfg.device));
    CellCols &c = ctx->d.cells;
    const int ndim3 = ctx->d.dims == D_THREE;
    const int n_padded = n > 0 ? ((n + SCAN_TILE - 1) / SCAN_TILE) * SCAN_TILE : SCAN_TILE;
    // a frame with the same number of cells reuses the device arrays (one frame per step in the driver)
    if (!ctx->have_hydro || c.n != n) {
        CK(cudaStreamSynchronize(ctx->stream)); // kernels of the previous frame may still read the arrays about to be freed
        free_pool(ctx->cell_allocs);
        double *cols[19];
        for (int f = 0; f < 19; ++f) CK(dev_alloc(ctx->cell_allocs, &cols[f], (size_t)n));
        double4 *geoA = nullptr;
        double2 *geoB = nullptr;
        CK(dev_alloc(ctx->cell_allocs, &geoA, (size_t)n_padded));
        CK(dev_alloc(ctx->cell_allocs, &geoB, (size_t)(ndim3 ? n_padded : 1)));
        c.geoA = geoA;
        c.geoB = geoB;
        c.nbox1 = (n + BOX_T - 1) / BOX_T;
        c.nbox2 = (c.nbox1 + BOX_T - 1) / BOX_T;
        double *b1 = nullptr, *b2 = nullptr;
        CK(dev_alloc(ctx->cell_allocs, &b1, (size_t)6 * (c.nbox1 ? c.nbox1 : 1)));
        CK(dev_alloc(ctx->cell_allocs, &b2, (size_t)6 * (c.nbox2 ? c.nbox2 : 1)));
        c.box1 = b1;
        c.box2 = b2;
        CK(dev_alloc(ctx->cell_allocs, &c.k2, (size_t)n));
    }
    CK(cudaMemsetAsync(c.k2, 0, (size_t)(n ? n : 1) * sizeof(double), ctx->stream)); // new temperatures
    c.n = n;
    c.n_padded = n_padded;
    double *cols[19];
    for (int f = 0; f < 19; ++f) {
        cols[f] = (double *)ctx->cell_allocs[f];
        if (fields[f])
            CK(cudaMemcpyAsync(cols[f], fields[f], (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        else
            CK(cudaMemsetAsync(cols[f], 0, (size_t)(n ? n : 1) * sizeof(double), ctx->stream));
    }
    c.r0 = cols[0]; c.r1 = cols[1]; c.r2 = cols[2];
    c.s0 = cols[3]; c.s1 = cols[4]; c.s2 = cols[5];
    c.v0 = cols[8]; c.v1 = cols[9]; c.v2 = cols[10];
    c.dens = cols[11]; c.dens_lab = cols[12]; c.temp = cols[14]; c.gamma = cols[15];
    c.B0 = cols[16]; c.B1 = cols[17]; c.B2 = cols[18];
    build_geo_kernel<<<grid_for(ctx, c.n_padded, 256, 8), 256, 0, ctx->stream>>>(
        ndim3, n, c.n_padded, cols[0], cols[1], cols[2], cols[3], cols[4], cols[5], (double4 *)c.geoA, (double2 *)c.geoB);
    if (int rc = check_launch(ctx, "build_geo_kernel")) return rc;
    {   // the index is always built (two tiny kernels): the persistent loop re-locates through it;
        // cfg.scan_index only decides whether the *streamed* path and the full rescan use it
        build_box1_kernel<<<grid_for(ctx, c.nbox1, 128, 8), 128, 0, ctx->stream>>>(ndim3, n, c.geoA, c.geoB, (double *)c.box1, c.nbox1);
        build_box2_kernel<<<grid_for(ctx, c.nbox2, 128, 8), 128, 0, ctx->stream>>>(c.box1, c.nbox1, (double *)c.box2, c.nbox2);
        if (int rc = check_launch(ctx, "build_box_kernels", 2)) return rc;
    }
    for (int k = 0; k < 6; ++k) c.dom[k] = domains[k];
    {   // safe_path(): a photon inside the domain has |position| <= R; one push rounds each component by <= ulp/2
        double r2sum = 0;
        for (int k = 0; k < (ndim3 ? 3 : 2); ++k) {
            const double m = fmax(fabs(domains[2 * k]), fabs(domains[2 * k + 1]));
            r2sum += m * m;
        }
        ctx->d.path_pad = 4e-16 * sqrt(r2sum);
        CK(cudaMemcpyAsync((void *)ctx->d.dom_dev, domains, 6 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        // thresholds refer to the cells of the previous frame
        if (ctx->have_photons && ctx->d.cap > 0)
            CK(cudaMemsetAsync(ctx->d.ph.safe, 0, sizeof(unsigned long long) * (size_t)ctx->d.cap, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream)); // the caller may reuse its (possibly pinned) arrays when this returns
    ctx->have_hydro = true;
    return MCRAT_B200_OK;
}

API int mcrat_b200_set_thermal_table(mcrat_b200_ctx *ctx, const double *table)
{
    if (!ctx || !table) return ctx ? fail(ctx, MCRAT_B200_ERR_ARG, "set_thermal_table: null") : MCRAT_B200_ERR_ARG;
    // za[j*(N_PH_E+1)+i] = thermal_table[i][j], Src/hot_x_section.c:482-488
    std::vector<double> za((size_t)(N_PH_E + 1) * (N_T + 1));
    for (int i = 0; i <= N_PH_E; i++)
        for (int j = 0; j <= N_T; j++) za[(size_t)j * (N_PH_E + 1) + i] = table[(size_t)i * (N_T + 1) + j];
    CK(cudaMemcpyAsync(ctx->table_dev, za.data(), za.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MCRAT_B200_OK;
}

// keep = number of leading slots whose contents must survive a re-allocation (0: the list is about to be overwritten)
static int ensure_photon_capacity(mcrat_b200_ctx *ctx, int n, int keep = 0)
{
    if (n <= ctx->ph_cap_alloc) return MCRAT_B200_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    const PhotonCols old = ctx->d.ph;
    std::vector<void *> old_allocs;
    if (keep > 0)
        old_allocs.swap(ctx->ph_allocs);
    else
        free_pool(ctx->ph_allocs);
    int cap = n + n / 8 + 1024;
    PhotonCols &p = ctx->d.ph;
    double **cols[23] = {&p.r0, &p.r1, &p.r2, &p.p0, &p.p1, &p.p2, &p.p3, &p.c0, &p.c1, &p.c2, &p.c3, &p.s0,
                         &p.s1, &p.s2, &p.s3, &p.nscatt, &p.weight, &p.tau, &p.tts, &p.v0, &p.v1, &p.v2, &p.ntau};
    for (int k = 0; k < 23; ++k) CK(dev_alloc(ctx->ph_allocs, cols[k], (size_t)cap));
    CK(dev_alloc(ctx->ph_allocs, &p.safe, (size_t)cap));
    CK(dev_alloc(ctx->ph_allocs, &p.idx, (size_t)cap));
    CK(dev_alloc(ctx->ph_allocs, &p.flags, (size_t)cap));
    CK(dev_alloc(ctx->ph_allocs, &p.type, (size_t)cap));
    DevCtx &d = ctx->d;
    d.reloc_cap = cap;
    CK(dev_alloc(ctx->ph_allocs, &d.reloc_slot, (size_t)cap));
    CK(dev_alloc(ctx->ph_allocs, &d.reloc_h0, (size_t)cap));
    CK(dev_alloc(ctx->ph_allocs, &d.reloc_h1, (size_t)cap));
    CK(dev_alloc(ctx->ph_allocs, &d.reloc_h2, (size_t)cap));
    CK(dev_alloc(ctx->ph_allocs, &d.reloc_best, (size_t)cap));
    CK(dev_alloc(ctx->ph_allocs, &d.prefix_block, (size_t)(cap / 256 + 2)));
    if (ctx->aos_dev) cudaFree(ctx->aos_dev);
    CK(cudaMalloc((void **)&ctx->aos_dev, (size_t)cap * sizeof(mcrat_photon)));
    ctx->aos_cap = cap;
    ctx->ph_cap_alloc = cap;
    if (keep > 0) {
        const PhotonCols &o = old;
        double *const src[23] = {o.r0, o.r1, o.r2, o.p0, o.p1, o.p2, o.p3, o.c0, o.c1, o.c2, o.c3, o.s0,
                                 o.s1, o.s2, o.s3, o.nscatt, o.weight, o.tau, o.tts, o.v0, o.v1, o.v2, o.ntau};
        for (int k = 0; k < 23; ++k)
            CK(cudaMemcpyAsync(*cols[k], src[k], (size_t)keep * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(p.safe, o.safe, (size_t)keep * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(p.idx, o.idx, (size_t)keep * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(p.flags, o.flags, (size_t)keep, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(p.type, o.type, (size_t)keep, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        free_pool(old_allocs);
    }
    return MCRAT_B200_OK;
}

API int mcrat_b200_set_num_shards(mcrat_b200_ctx *ctx, int num_shards)
{
    if (!ctx) return MCRAT_B200_ERR_ARG;
    if (num_shards < 1 || num_shards > MAX_SHARDS) return fail(ctx, MCRAT_B200_ERR_ARG, "set_num_shards: 1..4096 sub-shards");
    if (num_shards > 1 && ctx->d.replay) return fail(ctx, MCRAT_B200_ERR_ARG, "the replay harness drives a single shard");
    if (num_shards > 1 && ctx->d.cs)
        return fail(ctx, MCRAT_B200_ERR_ARG, "with CYCLOSYNCHROTRON_SWITCH ON a context holds one rank: emission, absorption and rebinning manage "
                                             "the null slots and counters of one list (use one context per rank; contexts on different streams run side by side)");
    ctx->want_shards = num_shards;
    return MCRAT_B200_OK;
}

API int mcrat_b200_num_shards(const mcrat_b200_ctx *ctx) { return ctx ? ctx->d.nshards : 0; }

// slot ranges of the sub-shards: contiguous, equal size (the last one may be shorter)
static int layout_shards(mcrat_b200_ctx *ctx, int n)
{
    DevCtx &d = ctx->d;
    int S = ctx->want_shards;
    if (S > n) S = n > 0 ? n : 1;
    const int size = n > 0 ? (n + S - 1) / S : 1;
    S = n > 0 ? (n + size - 1) / size : 1;
    if (int rc = fetch_state(ctx)) return rc; // keep per-shard iteration counters (RNG stream positions)
    const bool relayout = (S != d.nshards) || (size != d.shard_size);
    const int old_S = d.nshards;
    d.nshards = S;
    d.shard_size = size;
    int bps = (size + PASS_THREADS - 1) / PASS_THREADS;
    int cap_sm = (ctx->num_sms * MCRAT_PASS_CTAS_PER_SM) / S;
    if (cap_sm < 1) cap_sm = 1;
    if (bps > cap_sm) bps = cap_sm;
    if (bps > BLOCKMIN_CAP / S) bps = BLOCKMIN_CAP / S;
    if (bps < 1) bps = 1;
    d.blocks_per_shard = bps;
    // The iteration number is the only per-iteration word of the Philox counters (free path: (slot, iter, 0);
    // event: (draw / 2, iter, 1)) and the key holds seed and shard only, so a shard must never see an iteration
    // number twice in the life of a context: when the list is laid out anew (injection, growth, another number of
    // sub-shards) every shard continues from the largest iteration number any shard had reached.
    unsigned long long it_max = 0;
    for (int s = 0; s < old_S; ++s)
        if (ctx->sh_host[s].iter > it_max) it_max = ctx->sh_host[s].iter;
    for (int s = 0; s < S; ++s) {
        ShardState &sh = ctx->sh_host[s];
        if (relayout) {
            memset(&sh, 0, sizeof(sh));
            sh.iter = it_max;
        }
        sh.first = s * size;
        sh.count = (s * size + size <= n) ? size : (n - s * size);
        if (sh.count < 0) sh.count = 0;
        sh.n_dt = 0;
        sh.pushed_slot = -1;
    }
    CK(cudaMemcpyAsync(d.sh, ctx->sh_host.data(), sizeof(ShardState) * S, cudaMemcpyHostToDevice, ctx->stream));
    return MCRAT_B200_OK;
}

API int mcrat_b200_set_photons(mcrat_b200_ctx *ctx, const mcrat_photon *photons, int n)
{
    if (!ctx || n < 0 || (n > 0 && !photons)) return ctx ? fail(ctx, MCRAT_B200_ERR_ARG, "set_photons: bad argument") : MCRAT_B200_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    if (int rc = ensure_photon_capacity(ctx, n)) return rc;
    ctx->d.cap = n;
    ctx->d.stream_hints = (getenv("MCRAT_B200_NO_STREAM_HINTS") == nullptr && n > PERSISTENT_MAX_PHOTONS) ? 1 : 0;
    if (int rc = layout_shards(ctx, n)) return rc;
    if (n > 0) {
        CK(cudaMemcpyAsync(ctx->aos_dev, photons, (size_t)n * sizeof(mcrat_photon), cudaMemcpyHostToDevice, ctx->stream));
        unpack_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(ctx->d, ctx->aos_dev, n);
        if (int rc = check_launch(ctx, "unpack_kernel")) return rc;
    }
    ctx->have_photons = true;
    return MCRAT_B200_OK;
}

static int flush_pushes(mcrat_b200_ctx *ctx)
{
    if (ctx->d.cap > 0) {
        flush_push_kernel<<<grid_for(ctx, ctx->d.cap, PASS_THREADS, 8), PASS_THREADS, 0, ctx->stream>>>(ctx->d);
        if (int rc = check_launch(ctx, "flush_push_kernel")) return rc;
    }
    clear_push_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d);
    return check_launch(ctx, "clear_push_kernel");
}

API int mcrat_b200_get_photons(mcrat_b200_ctx *ctx, mcrat_photon *photons, int n)
{
    if (!ctx || !photons || n < 0 || n > ctx->d.cap) return ctx ? fail(ctx, MCRAT_B200_ERR_ARG, "get_photons: bad argument") : MCRAT_B200_ERR_ARG;
    if (int rc = flush_pushes(ctx)) return rc;
    if (n > 0) {
        pack_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(ctx->d, ctx->aos_dev, 0, n);
        if (int rc = check_launch(ctx, "pack_kernel")) return rc;
        CK(cudaMemcpyAsync(photons, ctx->aos_dev, (size_t)n * sizeof(mcrat_photon), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return MCRAT_B200_OK;
}

API int mcrat_b200_get_photon(mcrat_b200_ctx *ctx, int index, mcrat_photon *out)
{
    if (!ctx || !out || index < 0 || index >= ctx->d.cap) return ctx ? fail(ctx, MCRAT_B200_ERR_ARG, "get_photon: bad index") : MCRAT_B200_ERR_ARG;
    if (int rc = flush_pushes(ctx)) return rc;
    pack_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d, ctx->aos_dev, index, 1);
    if (int rc = check_launch(ctx, "pack_kernel")) return rc;
    CK(cudaMemcpyAsync(out, ctx->aos_dev, sizeof(mcrat_photon), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MCRAT_B200_OK;
}

API int mcrat_b200_list_capacity(const mcrat_b200_ctx *ctx) { return ctx ? ctx->d.cap : 0; }

API int mcrat_b200_set_replay_uniforms(mcrat_b200_ctx *ctx, const double *u, size_t n)
{
    if (!ctx || (!u && n)) return ctx ? fail(ctx, MCRAT_B200_ERR_ARG, "set_replay_uniforms: null") : MCRAT_B200_ERR_ARG;
    if (!ctx->d.replay) return fail(ctx, MCRAT_B200_ERR_STATE, "context was not created with MCRAT_RNG_REPLAY");
    // gsl_rng_uniform_pos redraws on an exact 0 (1 in 2^24 for ranlxs0), which would shift every
    // later free-path draw; the harness must supply a zero-free stream
    for (size_t i = 0; i < n; ++i)
        if (u[i] == 0.0) return fail(ctx, MCRAT_B200_ERR_ARG, "replay stream contains an exact 0.0; pick another seed");
    if (n > ctx->replay_cap) {
        if (ctx->replay_dev) cudaFree(ctx->replay_dev);
        CK(cudaMalloc((void **)&ctx->replay_dev, (n ? n : 1) * sizeof(double)));
        ctx->replay_cap = n;
    }
    CK(cudaMemcpyAsync(ctx->replay_dev, u, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->d.replay_buf = ctx->replay_dev;
    if (int rc = fetch_global(ctx)) return rc;
    ctx->gs_host->replay_cursor = 0;
    ctx->gs_host->replay_base = 0;
    ctx->gs_host->replay_n = n;
    CK(cudaMemcpyAsync(ctx->d.gs, ctx->gs_host, sizeof(GlobalState), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MCRAT_B200_OK;
}

API long long mcrat_b200_replay_consumed(mcrat_b200_ctx *ctx)
{
    if (!ctx) return -1;
    if (fetch_global(ctx)) return -1;
    return (long long)ctx->gs_host->replay_cursor;
}

// ---- launch helpers ---------------------------------------------------------------------------
static int need_ready(mcrat_b200_ctx *ctx)
{
    if (!ctx->have_hydro) return fail(ctx, MCRAT_B200_ERR_STATE, "no hydro frame loaded (mcrat_b200_set_hydro)");
    if (!ctx->have_photons) return fail(ctx, MCRAT_B200_ERR_STATE, "no photon list loaded (mcrat_b200_set_photons)");
    if (ctx->d.cells.n <= 0) return fail(ctx, MCRAT_B200_ERR_STATE, "hydro frame has no cells");
    if (ctx->d.cap <= 0) return fail(ctx, MCRAT_B200_ERR_STATE, "photon list is empty");
    return MCRAT_B200_OK;
}

static int need_single_shard(mcrat_b200_ctx *ctx, const char *what)
{
    if (ctx->d.nshards != 1) {
        ctx->err = std::string(what) + ": the step-by-step surface works on one shard (set_num_shards(1))";
        return MCRAT_B200_ERR_STATE;
    }
    return MCRAT_B200_OK;
}

// K1 is a persistent launch: as many CTAs as fit on the device, pulling (photon chunk, cell chunk) items from a
// counter.  Items are sized for ~MCRAT_SCAN_ITEMS_PER_CTA items per resident CTA (the SMs run dry only during a CTA's
// last item), between 4 and 64 tiles of cells each (an item re-loads its photons: 27 loads against >= 4 x 256 x 9 x 7
// instructions).
static void scan_grid(mcrat_b200_ctx *ctx, int nphot, int &grid, int &tiles_per_item)
{
    const int ndim3 = ctx->d.dims == D_THREE;
    const int ntiles = ctx->d.cells.n_padded / SCAN_TILE;
    const int SCAN_P = ndim3 ? SCAN_P3 : SCAN_P2;
    long long pchunks = ((long long)nphot + SCAN_THREADS * SCAN_P - 1) / (SCAN_THREADS * SCAN_P);
    if (pchunks < 1) pchunks = 1;
    const int resident = ctx->num_sms * (ctx->occ_scan[ndim3] > 0 ? ctx->occ_scan[ndim3] : 1);
    long long tpi = ((long long)ntiles * pchunks) / ((long long)MCRAT_SCAN_ITEMS_PER_CTA * resident);
    if (tpi < 4) tpi = 4;
    if (tpi > 64) tpi = 64;
    if (tpi > ntiles) tpi = ntiles;
    tiles_per_item = (int)tpi;
    const long long items = pchunks * ((ntiles + tpi - 1) / tpi);
    grid = (int)(items < resident ? items : resident);
    if (grid < 1) grid = 1;
}

// full scan of the current relocation list with K1 (count known only on the device: sized for cap)
static int launch_scan_full(mcrat_b200_ctx *ctx, int parity, int nphot_bound)
{
    int grid, tpc;
    scan_grid(ctx, nphot_bound, grid, tpc);
    Timed t(ctx, KC_SCAN);
    if (ctx->d.dims == D_THREE)
        scan_kernel<1><<<grid, SCAN_THREADS, scan_smem_bytes(1), ctx->stream>>>(ctx->d, parity, tpc);
    else
        scan_kernel<0><<<grid, SCAN_THREADS, scan_smem_bytes(0), ctx->stream>>>(ctx->d, parity, tpc);
    return check_launch(ctx, "scan_kernel");
}

static int launch_scan_few(mcrat_b200_ctx *ctx, int parity)
{
    int g = grid_for(ctx, ctx->d.cells.n, 256, 4);
    Timed t(ctx, KC_SCAN);
    if (ctx->d.dims == D_THREE)
        scan_few_kernel<1><<<g, 256, 0, ctx->stream>>>(ctx->d, parity);
    else
        scan_few_kernel<0><<<g, 256, 0, ctx->stream>>>(ctx->d, parity);
    return check_launch(ctx, "scan_few_kernel");
}

// one locate step: pass (+ optional fused free-path draw), scan, finish.  Returns the parity used.
template <bool FUSE>
static int launch_locate(mcrat_b200_ctx *ctx, int sw, int &parity_out)
{
    const int parity = ctx->pass_parity;
    ctx->pass_parity ^= 1;
    parity_out = parity;
    const int nb_pass = ctx->d.nshards * ctx->d.blocks_per_shard;
    {
        Timed t(ctx, KC_PASS);
        pass_kernel<FUSE><<<nb_pass, PASS_THREADS, 0, ctx->stream>>>(ctx->d, sw, parity);
        if (int rc = check_launch(ctx, "pass_kernel")) return rc;
    }
    int nb_fin;
    if (ctx->cfg.scan_index) {
        Timed t(ctx, KC_SCAN);
        // one warp per photon
        // one warp per photon; in a steady-state iteration a few in every 10^4 photons change cell
        int g = grid_for(ctx, ctx->d.cap > (INT_MAX >> 5) ? INT_MAX : ctx->d.cap * 32, 128, 16);
        if (sw == 0) {
            g = ctx->d.cap / 2048;
            if (g < 32) g = 32;
            if (g > ctx->num_sms * 16) g = ctx->num_sms * 16;
        }
        scan_index_kernel<<<g, 128, 0, ctx->stream>>>(ctx->d, parity);
        if (int rc = check_launch(ctx, "scan_index_kernel")) return rc;
        nb_fin = (sw == 1) ? grid_for(ctx, ctx->d.cap, FIN_THREADS, 8) : (g / 4 < 8 ? 8 : g / 4);
    } else if (sw == 1) {
        if (int rc = launch_scan_full(ctx, parity, ctx->d.cap)) return rc;
        nb_fin = grid_for(ctx, ctx->d.cap, FIN_THREADS, 8);
    } else {
        if (int rc = launch_scan_few(ctx, parity)) return rc;
        nb_fin = 8;
    }
    {
        Timed t(ctx, KC_OTHER);
        finish_kernel<FUSE><<<nb_fin, FIN_THREADS, 0, ctx->stream>>>(ctx->d, sw, parity);
        if (int rc = check_launch(ctx, "finish_kernel")) return rc;
    }
    return MCRAT_B200_OK;
}

static int launch_mfp_unfused(mcrat_b200_ctx *ctx, int &nb)
{
    const int nblocks = (ctx->d.cap + 255) / 256;
    Timed t(ctx, KC_PASS);
    if (ctx->d.replay) {
        mfp_count_kernel<<<nblocks, 256, 0, ctx->stream>>>(ctx->d);
        mfp_scan_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d, nblocks);
        if (int rc = check_launch(ctx, "mfp_scan_kernel", 2)) return rc;
    }
    mfp_kernel<<<nblocks, 256, 0, ctx->stream>>>(ctx->d, nblocks <= BLOCKMIN_CAP ? 1 : 0);
    if (int rc = check_launch(ctx, "mfp_kernel")) return rc;
    if (nblocks > BLOCKMIN_CAP) {
        nb = grid_for(ctx, ctx->d.cap, 256, 8);
        argmin_all_kernel<<<nb, 256, 0, ctx->stream>>>(ctx->d);
        if (int rc = check_launch(ctx, "argmin_all_kernel")) return rc;
    } else {
        nb = nblocks;
    }
    return MCRAT_B200_OK;
}

static int device_error(mcrat_b200_ctx *ctx)
{
    const int e = ctx->gs_host->error;
    if (!e) return MCRAT_B200_OK;
    static const char *const sites[] = {"?", "pass: optical depth", "pass: verified re-check skip", "finish: optical depth",
                                        "calcMeanFreePath: optical depth", "calcMeanFreePath: replay stream",
                                        "photonEvent: replay stream", "photonEvent: mini-pass optical depth",
                                        "persistent loop: a block waited > 1 s for its team", "rebin", "cyclo-synchrotron emission"};
    const int site = ctx->gs_host->error_site;
    const char *what = e == MCRAT_B200_ERR_REPLAY ? "replay uniform stream exhausted"
                       : e == MCRAT_B200_ERR_TABLE ? "hot cross-section lookup outside the table"
                                                   : "device-side error";
    char buf[256];
    snprintf(buf, sizeof(buf), "%s (first raised in %s, photon slot %d)", what,
             (site >= 0 && site < (int)(sizeof(sites) / sizeof(sites[0]))) ? sites[site] : "?", ctx->gs_host->error_slot);
    return fail(ctx, e, buf);
}

__global__ void reset_loop_kernel(DevCtx d, int set_times, double time_now, double remaining, long long max_iters)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < d.nshards; s += gridDim.x * blockDim.x) {
        ShardState &st = d.sh[s];
        st.done = 0;
        st.pause_cs = 0;
        st.counted_stopped = 0;
        st.iters_done = 0;
        st.arrive = 0;
        st.gen = 0;
        st.reloc_n = 0;
        st.halt = 0;
        st.reloc_heavy = 0;
        st.mini_slot = -1;
        if (set_times) {
            st.time_now = time_now;
            st.remaining_time = remaining;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        d.gs->max_iters = max_iters;
        d.gs->n_stopped = 0;
        d.gs->reloc_heavy_any = 0;
        d.gs->stream_work = 0;
        d.gs->stream_evt_ready = 0;
        d.gs->stream_halted = 0;
        d.gs->stream_state = 0;
    }
}

static int reset_loop(mcrat_b200_ctx *ctx, int set_times, double time_now, double remaining, long long max_iters)
{
    reset_loop_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d, set_times, time_now, remaining, max_iters);
    return check_launch(ctx, "reset_loop_kernel");
}

// ---- reference function surface ------------------------------------------------------------------
API int mcrat_b200_find_containing_hydro_cell(mcrat_b200_ctx *ctx, int sw, int *num_relocated)
{
    if (!ctx) return MCRAT_B200_ERR_ARG;
    if (int rc = need_ready(ctx)) return rc;
    if (int rc = need_single_shard(ctx, "find_containing_hydro_cell")) return rc;
    if (int rc = reset_loop(ctx, 0, 0, 0, -1)) return rc;
    if (int rc = fetch_state(ctx)) return rc;
    const long long before = ctx->sh_host[0].reloc_total;
    int parity;
    if (int rc = launch_locate<false>(ctx, sw ? 1 : 0, parity)) return rc;
    clear_push_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d); // the pass consumed the pending pushes
    if (int rc = check_launch(ctx, "clear_push_kernel")) return rc;
    if (int rc = fetch_state(ctx)) return rc;
    if (num_relocated) *num_relocated = (int)(ctx->sh_host[0].reloc_total - before);
    return device_error(ctx);
}

API int mcrat_b200_calc_mean_free_path(mcrat_b200_ctx *ctx, int *first_index, double *first_tts)
{
    if (!ctx) return MCRAT_B200_ERR_ARG;
    if (int rc = need_ready(ctx)) return rc;
    if (int rc = need_single_shard(ctx, "calc_mean_free_path")) return rc;
    if (int rc = flush_pushes(ctx)) return rc;
    int nb;
    if (int rc = launch_mfp_unfused(ctx, nb)) return rc;
    head_kernel<<<1, EVT_THREADS, 0, ctx->stream>>>(ctx->d, nb);
    if (int rc = check_launch(ctx, "head_kernel")) return rc;
    if (int rc = fetch_state(ctx)) return rc;
    if (first_index) *first_index = ctx->sh_host[0].head_idx;
    if (first_tts) *first_tts = ctx->sh_host[0].head_tts;
    ctx->last_nb_mfp = nb; // block minima stay valid for a following photon_event
    return device_error(ctx);
}

__global__ void iota_kernel(int *v, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = i;
}

// The whole time order, for hosts that read more of photonList.sorted_indexes than its head (Src/mclib.c:717-729: qsort_r
// of the slot indices by time_to_scatter).  A stable radix sort of the time column on the device: equal times (the 1e12 / c of
// every photon outside the domain, Src/mclib.c:684-687) come in ascending slot order, where the reference's qsort leaves
// their order to the C library.  time_to_scatter >= 0, so the bit pattern of the doubles sorts like the values.
API int mcrat_b200_get_sorted_indexes(mcrat_b200_ctx *ctx, int *sorted_indexes, int n)
{
    if (!ctx || !sorted_indexes || n < 0 || n > ctx->d.cap) return ctx ? fail(ctx, MCRAT_B200_ERR_ARG, "get_sorted_indexes: bad argument") : MCRAT_B200_ERR_ARG;
    if (int rc = need_ready(ctx)) return rc;
    if (int rc = need_single_shard(ctx, "get_sorted_indexes")) return rc;
    if (n == 0) return MCRAT_B200_OK;
    CK(cudaSetDevice(ctx->cfg.device));
    double *keys_out = nullptr;
    int *vals_in = nullptr, *vals_out = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t e = cudaMalloc((void **)&keys_out, (size_t)n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&vals_in, (size_t)n * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void **)&vals_out, (size_t)n * sizeof(int));
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ctx->d.ph.tts, keys_out, vals_in, vals_out, n, 0, 64, ctx->stream);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1);
    if (e == cudaSuccess) {
        iota_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(vals_in, n);
        ctx->launches++;
        e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, ctx->d.ph.tts, keys_out, vals_in, vals_out, n, 0, 64, ctx->stream);
        ctx->launches++;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(sorted_indexes, vals_out, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(keys_out);
    cudaFree(vals_in);
    cudaFree(vals_out);
    cudaFree(tmp);
    CK(e);
    CK(e2);
    return MCRAT_B200_OK;
}

API int mcrat_b200_photon_event(mcrat_b200_ctx *ctx, double dt_max, double *time_step, int *scattered_ph_index,
                                int *frame_scatt_cnt, int *frame_abs_cnt)
{
    (void)frame_abs_cnt; // the reference never touches it either (Src/mclib.c:1107-1356)
    if (!ctx) return MCRAT_B200_ERR_ARG;
    if (int rc = need_ready(ctx)) return rc;
    if (int rc = need_single_shard(ctx, "photon_event")) return rc;
    if (ctx->last_nb_mfp <= 0) return fail(ctx, MCRAT_B200_ERR_STATE, "photon_event needs a preceding calc_mean_free_path");
    if (int rc = reset_loop(ctx, 0, 0, 0, -1)) return rc;
    if (int rc = fetch_state(ctx)) return rc;
    const long long before = ctx->sh_host[0].scatt_cnt;
    {
        Timed t(ctx, KC_EVENT);
        event_kernel<EVT_THREADS><<<1, EVT_THREADS, 0, ctx->stream>>>(ctx->d, 0, ctx->last_nb_mfp, 1, dt_max);
        if (int rc = check_launch(ctx, "event_kernel")) return rc;
    }
    if (int rc = fetch_state(ctx)) return rc;
    if (time_step) *time_step = ctx->sh_host[0].last_time_step;
    if (scattered_ph_index) *scattered_ph_index = ctx->sh_host[0].last_scattered_idx;
    if (frame_scatt_cnt) *frame_scatt_cnt += (int)(ctx->sh_host[0].scatt_cnt - before);
    ctx->last_nb_mfp = 0;
    return device_error(ctx);
}

API int mcrat_b200_update_photon_position(mcrat_b200_ctx *ctx, double t)
{
    if (!ctx) return MCRAT_B200_ERR_ARG;
    if (!ctx->have_photons) return fail(ctx, MCRAT_B200_ERR_STATE, "no photon list loaded");
    if (int rc = flush_pushes(ctx)) return rc;
    set_push_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d, t);
    if (int rc = check_launch(ctx, "set_push_kernel")) return rc;
    Timed tm(ctx, KC_PASS);
    return flush_pushes(ctx);
}

__global__ void clear_abs_kernel(DevCtx d)
{
    d.gs->abs_count = 0;
    d.gs->cs_scatt_count = 0;
    d.gs->abs_weight = 0;
}

API int mcrat_b200_ph_abs_cyclosynch(mcrat_b200_ctx *ctx, int *num_abs_ph, int *scatt_cyclosynch_num_ph, double *absorbed_weight)
{
    if (!ctx) return MCRAT_B200_ERR_ARG;
    if (int rc = need_ready(ctx)) return rc;
    if (int rc = flush_pushes(ctx)) return rc;
    clear_abs_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d);
    if (int rc = check_launch(ctx, "clear_abs_kernel")) return rc;
    cs_absorb_kernel<<<grid_for(ctx, ctx->d.cap, 256, 8), 256, 0, ctx->stream>>>(ctx->d);
    if (int rc = check_launch(ctx, "cs_absorb_kernel")) return rc;
    if (int rc = fetch_global(ctx)) return rc;
    if (num_abs_ph) *num_abs_ph = ctx->gs_host->abs_count;
    if (scatt_cyclosynch_num_ph) *scatt_cyclosynch_num_ph = ctx->gs_host->cs_scatt_count;
    if (absorbed_weight) *absorbed_weight = ctx->gs_host->abs_weight;
    return MCRAT_B200_OK;
}

static int run_stats(mcrat_b200_ctx *ctx, StatPartial &tot)
{
    if (!ctx->have_photons) return fail(ctx, MCRAT_B200_ERR_STATE, "no photon list loaded");
    if (int rc = flush_pushes(ctx)) return rc;
    int g = grid_for(ctx, ctx->d.cap, 256, 4);
    if (g > 1024) g = 1024;
    stats_kernel<<<g, 256, 0, ctx->stream>>>(ctx->d, ctx->stat_dev);
    if (int rc = check_launch(ctx, "stats_kernel")) return rc;
    std::vector<StatPartial> h(g);
    CK(cudaMemcpyAsync(h.data(), ctx->stat_dev, g * sizeof(StatPartial), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    tot = h[0];
    for (int k = 1; k < g; ++k) {
        tot.e_sum += h[k].e_sum; tot.w_sum += h[k].w_sum; tot.ns_sum += h[k].ns_sum; tot.r_sum += h[k].r_sum;
        tot.r_min = std::fmin(tot.r_min, h[k].r_min); tot.r_max = std::fmax(tot.r_max, h[k].r_max);
        tot.th_min = std::fmin(tot.th_min, h[k].th_min); tot.th_max = std::fmax(tot.th_max, h[k].th_max);
        tot.count += h[k].count;
        tot.ns_max = h[k].ns_max > tot.ns_max ? h[k].ns_max : tot.ns_max;
        tot.ns_min = h[k].ns_min < tot.ns_min ? h[k].ns_min : tot.ns_min;
    }
    return MCRAT_B200_OK;
}

API int mcrat_b200_ph_min_max(mcrat_b200_ctx *ctx, double *min_r, double *max_r, double *min_theta, double *max_theta)
{
    if (!ctx) return MCRAT_B200_ERR_ARG;
    StatPartial t;
    if (int rc = run_stats(ctx, t)) return rc;
    if (min_r) *min_r = t.r_min;
    if (max_r) *max_r = t.r_max;
    if (min_theta) *min_theta = t.th_min;
    if (max_theta) *max_theta = t.th_max;
    return MCRAT_B200_OK;
}

API int mcrat_b200_ph_scatt_stats(mcrat_b200_ctx *ctx, int *max_scatt, int *min_scatt, double *avg_scatt, double *avg_r)
{
    if (!ctx) return MCRAT_B200_ERR_ARG;
    StatPartial t;
    if (int rc = run_stats(ctx, t)) return rc;
    if (max_scatt) *max_scatt = t.ns_max;
    if (min_scatt) *min_scatt = t.ns_min;
    if (avg_scatt) *avg_scatt = t.ns_sum / (double)t.count;
    if (avg_r) *avg_r = t.r_sum / (double)t.count;
    return MCRAT_B200_OK;
}

API int mcrat_b200_average_photon_energy(mcrat_b200_ctx *ctx, double *avg_energy)
{
    if (!ctx) return MCRAT_B200_ERR_ARG;
    StatPartial t;
    if (int rc = run_stats(ctx, t)) return rc;
    if (avg_energy) *avg_energy = (t.e_sum * 2.99792458e10) / t.w_sum;
    return MCRAT_B200_OK;
}

// rebinCyclosynchCompPhotons (Src/mc_cyclosynch.h:86, Src/mc_cyclosynch.c:600-710) without the list leaving the device.
// The host only sizes the histograms (calculate_binning_params, :325-347, and GSL's uniform ranges).
API int mcrat_b200_rebin_cyclosynch_comp_photons(mcrat_b200_ctx *ctx, int max_photons, int *num_cyclosynch_ph_emit,
                                                 int *scatt_cyclosynch_num_ph, int *num_null_rebin_ph)
{
    if (!ctx) return MCRAT_B200_ERR_ARG;
    if (int rc = need_ready(ctx)) return rc;
    if (int rc = need_single_shard(ctx, "rebin_cyclosynch_comp_photons")) return rc;
    CK(cudaSetDevice(ctx->cfg.device));
    if (int rc = flush_pushes(ctx)) return rc;
    DevCtx &d = ctx->d;
    const int ndim3 = (d.dims == D_THREE);
    // ---- phase 1: ranges ----
    int g = grid_for(ctx, d.cap, 256, 4);
    if (g > 1024) g = 1024;
    RebinRange *part = nullptr;
    CK(cudaMalloc((void **)&part, sizeof(RebinRange) * g));
    rebin_range_kernel<<<g, 256, 0, ctx->stream>>>(d, part);
    if (int rc = check_launch(ctx, "rebin_range_kernel")) { cudaFree(part); return rc; }
    std::vector<RebinRange> hp(g);
    CK(cudaMemcpyAsync(hp.data(), part, sizeof(RebinRange) * g, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(part);
    RebinRange info = hp[0];
    for (int k = 1; k < g; ++k) {
        info.p0_min = std::fmin(info.p0_min, hp[k].p0_min); info.p0_max = std::fmax(info.p0_max, hp[k].p0_max);
        info.theta_min = std::fmin(info.theta_min, hp[k].theta_min); info.theta_max = std::fmax(info.theta_max, hp[k].theta_max);
        info.phi_min = std::fmin(info.phi_min, hp[k].phi_min); info.phi_max = std::fmax(info.phi_max, hp[k].phi_max);
        info.valid_photon_count += hp[k].valid_photon_count; info.synch_photon_count += hp[k].synch_photon_count;
    }
    if (info.valid_photon_count <= 0) return fail(ctx, MCRAT_B200_ERR_STATE, "rebin: no valid photons found for rebinning (Src/mc_cyclosynch.c:613)");
    const double log_p0_min = (info.p0_min > 0 && info.p0_max > 0) ? log10(info.p0_min) : 0.0;
    const double log_p0_max = (info.p0_min > 0 && info.p0_max > 0) ? log10(info.p0_max) : 1.0;
    // ---- phase 2: binning parameters, :325-347 ----
    const double rebin_e_perc = ctx->cs_rebin_e_perc, rebin_ang = ctx->cs_rebin_ang, rebin_ang_phi = ctx->cs_rebin_ang_phi;
    const double deg_to_rad = 3.14159265358979323846 / 180.0;
    RebinParams p;
    p.num_bins = (int)(rebin_e_perc * max_photons);
    p.num_bins_theta = (int)ceil((info.theta_max - info.theta_min) / (rebin_ang * deg_to_rad));
    p.num_bins_phi = 1;
    if (ndim3) p.num_bins_phi = (int)ceil((info.phi_max - info.phi_min) / rebin_ang_phi);
    p.total_bins = p.num_bins_theta * p.num_bins;
    if (ndim3) p.total_bins *= p.num_bins_phi;
    if (p.total_bins > max_photons) return fail(ctx, MCRAT_B200_ERR_STATE, "rebin: would create more photons than max_photons (Src/mc_cyclosynch.c:637)");
    if (p.num_bins <= 0 || p.num_bins_theta <= 0 || p.num_bins_phi <= 0) return fail(ctx, MCRAT_B200_ERR_STATE, "rebin: invalid histogram dimensions (Src/mc_cyclosynch.c:352)");
    // gsl_histogram2d_set_ranges_uniform with the reference's epsilons (:363-390)
    auto uniform = [](std::vector<double> &r, int n, double lo, double hi) {
        r.resize((size_t)n + 1);
        for (int i = 0; i <= n; i++) {
            double f1 = ((double)(n - i) / (double)n);
            double f2 = ((double)i / (double)n);
            r[(size_t)i] = f1 * lo + f2 * hi;
        }
    };
    const double e_eps = (log_p0_max - log_p0_min) * 1e-6, t_eps = (info.theta_max - info.theta_min) * 1e-6;
    const double p_eps = (info.phi_max - info.phi_min) * 1e-6;
    std::vector<double> re, rt, rp;
    uniform(re, p.num_bins, log_p0_min, log_p0_max + e_eps);
    uniform(rt, p.num_bins_theta, info.theta_min, info.theta_max + t_eps);
    if (ndim3) uniform(rp, p.num_bins_phi, info.phi_min, info.phi_max + p_eps); else rp.assign(2, 0.0);
    double *ranges = nullptr;
    int *bin_of = nullptr, *counters = nullptr;
    mcrat_photon *rebinned = nullptr;
    auto cleanup = [&]() {
        if (ranges) cudaFree(ranges);
        if (bin_of) cudaFree(bin_of);
        if (counters) cudaFree(counters);
        if (rebinned) cudaFree(rebinned);
    };
    const size_t nr = re.size() + rt.size() + rp.size();
    if (cudaMalloc((void **)&ranges, nr * sizeof(double)) != cudaSuccess || cudaMalloc((void **)&bin_of, (size_t)d.cap * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&counters, 2 * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&rebinned, (size_t)p.total_bins * sizeof(mcrat_photon)) != cudaSuccess) {
        cleanup();
        return fail(ctx, MCRAT_B200_ERR_CUDA, "rebin: cudaMalloc");
    }
    cudaMemcpyAsync(ranges, re.data(), re.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(ranges + re.size(), rt.data(), rt.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(ranges + re.size() + rt.size(), rp.data(), rp.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemsetAsync(counters, 0, 2 * sizeof(int), ctx->stream);
    p.range_e = ranges;
    p.range_theta = ranges + re.size();
    p.range_phi = ranges + re.size() + rt.size();
    // ---- phases 4-5 ----
    const int nblocks = (d.cap + 255) / 256;
    rebin_index_kernel<<<grid_for(ctx, d.cap, 256, 8), 256, 0, ctx->stream>>>(d, p, bin_of);
    rebin_accumulate_kernel<<<(p.total_bins + 127) / 128, 128, 0, ctx->stream>>>(d, p, bin_of, rebinned);
    rebin_null_kernel<<<nblocks, 256, 0, ctx->stream>>>(d);
    rebin_scan_kernel<<<1, 32, 0, ctx->stream>>>(d, nblocks, counters);
    rebin_count_kernel<<<grid_for(ctx, p.total_bins, 256, 4), 256, 0, ctx->stream>>>(rebinned, p.total_bins, counters + 1);
    if (int rc = check_launch(ctx, "rebin kernels", 5)) { cleanup(); return rc; }
    int hc[2] = {0, 0};
    cudaMemcpyAsync(hc, counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cleanup(); return fail(ctx, MCRAT_B200_ERR_CUDA, "rebin: synchronize"); }
    if (hc[0] < p.total_bins) {
        // addToPhotonList would have to grow the list (Src/photons.c:117-129): that is the host's job.  The
        // 'k' / 'c' photons are already nulled, exactly as the reference has done at this point (:588-596).
        cleanup();
        return fail(ctx, MCRAT_B200_ERR_STATE, "rebin: fewer null slots than rebinned photons; download, grow the list (addToPhotonList) and upload");
    }
    rebin_place_kernel<<<nblocks, 256, 0, ctx->stream>>>(d, rebinned, p.total_bins);
    if (int rc = check_launch(ctx, "rebin_place_kernel")) { cleanup(); return rc; }
    const int null_count = hc[1];
    // counters of the driver, :680-684
    const int scatt = p.total_bins - null_count;
    if (scatt_cyclosynch_num_ph) *scatt_cyclosynch_num_ph = scatt;
    if (num_cyclosynch_ph_emit) *num_cyclosynch_ph_emit = p.total_bins + info.synch_photon_count - null_count;
    if (num_null_rebin_ph) *num_null_rebin_ph = null_count;
    if (int rc = fetch_global(ctx)) { cleanup(); return rc; }
    ctx->gs_host->cs_scatt_num = scatt;
    CK(cudaMemcpyAsync(ctx->d.gs, ctx->gs_host, sizeof(GlobalState), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cleanup();
    return device_error(ctx);
}

// photonEmitCyclosynch with inject_single_switch == 0 (Src/mc_cyclosynch.h:90, Src/mc_cyclosynch.c:1176-1464) on the device,
// see cs_emit.cuh.  The list grows on the device when it has fewer null slots than new photons (the reference's
// addToPhotonList reallocates, Src/photons.c:117-129; new slots are null photons and sit at the end, so the slots the new
// photons land in are the reference's whatever the capacity becomes).
static int grow_list(mcrat_b200_ctx *ctx, int new_cap)
{
    const int old_cap = ctx->d.cap;
    if (new_cap <= old_cap) return MCRAT_B200_OK;
    if (int rc = ensure_photon_capacity(ctx, new_cap, old_cap)) return rc;
    ctx->d.cap = new_cap;
    init_null_kernel<<<grid_for(ctx, new_cap - old_cap, 256, 8), 256, 0, ctx->stream>>>(ctx->d, old_cap, new_cap - old_cap);
    if (int rc = check_launch(ctx, "init_null_kernel")) return rc;
    ctx->d.stream_hints = (getenv("MCRAT_B200_NO_STREAM_HINTS") == nullptr && new_cap > PERSISTENT_MAX_PHOTONS) ? 1 : 0;
    return layout_shards(ctx, new_cap);
}

static int count_null_slots(mcrat_b200_ctx *ctx, int *counter_dev, int *nulls)
{
    const int nblocks = (ctx->d.cap + 255) / 256;
    count_null_kernel<<<nblocks, 256, 0, ctx->stream>>>(ctx->d);
    rebin_scan_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d, nblocks, counter_dev);
    if (int rc = check_launch(ctx, "count_null_kernel", 2)) return rc;
    CK(cudaMemcpyAsync(nulls, counter_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MCRAT_B200_OK;
}

API int mcrat_b200_photon_emit_cyclosynch(mcrat_b200_ctx *ctx, double r_inj, double ph_weight, int maximum_photons,
                                          double theta_min, double theta_max, int *num_emitted, double *ph_weight_adjusted,
                                          int *num_cells_selected)
{
    if (!ctx) return MCRAT_B200_ERR_ARG;
    if (int rc = need_ready(ctx)) return rc;
    if (int rc = need_single_shard(ctx, "photon_emit_cyclosynch")) return rc;
    if (!(ph_weight > 0) || maximum_photons < 1) return fail(ctx, MCRAT_B200_ERR_ARG, "photon_emit_cyclosynch: ph_weight > 0, maximum_photons >= 1");
    CK(cudaSetDevice(ctx->cfg.device));
    if (int rc = flush_pushes(ctx)) return rc;
    DevCtx &d = ctx->d;
    const int n = d.cells.n, nblocks = (n + 255) / 256;
    const double rmin = mcrat_b200_calc_cyclosynch_r_limits(ctx->hydro_scatt_frame, ctx->hydro_inj_frame, ctx->hydro_fps, r_inj, "min");
    const double rmax = mcrat_b200_calc_cyclosynch_r_limits(ctx->hydro_scatt_frame, ctx->hydro_inj_frame, ctx->hydro_fps, r_inj, "max");
    const double max_photons = ctx->cs_rebin_e_perc * maximum_photons; // Src/mc_cyclosynch.c:1178
    std::vector<void *> pool;
    mcrat_photon *emitted = nullptr;
    auto cleanup = [&]() {
        free_pool(pool);
        if (emitted) cudaFree(emitted);
    };
    CsEmitWork w;
    w.n_cells = n;
    if (dev_alloc(pool, &w.flag, (size_t)n) != cudaSuccess || dev_alloc(pool, &w.block_base, (size_t)nblocks + 1) != cudaSuccess ||
        dev_alloc(pool, &w.meta, 4) != cudaSuccess || dev_alloc(pool, &w.weight, 1) != cudaSuccess) {
        cleanup();
        return fail(ctx, MCRAT_B200_ERR_CUDA, "photon_emit_cyclosynch: cudaMalloc");
    }
    w.sel_cell = nullptr; w.integ = w.vol = w.nu_c = nullptr; w.count = nullptr; w.offset = nullptr;
    cudaMemsetAsync(w.meta, 0, 4 * sizeof(int), ctx->stream);
    cs_select_kernel<<<nblocks, 256, 0, ctx->stream>>>(d, w, rmin, rmax, theta_min, theta_max);
    cs_block_scan_kernel<<<1, 1024, 0, ctx->stream>>>(w, nblocks);
    if (int rc = check_launch(ctx, "cs_select_kernel", 2)) { cleanup(); return rc; }
    int meta[4] = {0, 0, 0, 0};
    cudaMemcpyAsync(meta, w.meta, sizeof(meta), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cleanup(); return fail(ctx, MCRAT_B200_ERR_CUDA, "photon_emit_cyclosynch: synchronize"); }
    const int n_sel = meta[0];
    if (num_cells_selected) *num_cells_selected = n_sel;
    const size_t ns = (size_t)(n_sel > 0 ? n_sel : 1);
    if (dev_alloc(pool, &w.sel_cell, ns) != cudaSuccess || dev_alloc(pool, &w.integ, ns) != cudaSuccess ||
        dev_alloc(pool, &w.vol, ns) != cudaSuccess || dev_alloc(pool, &w.nu_c, ns) != cudaSuccess ||
        dev_alloc(pool, &w.count, ns) != cudaSuccess || dev_alloc(pool, &w.offset, ns + 1) != cudaSuccess) {
        cleanup();
        return fail(ctx, MCRAT_B200_ERR_CUDA, "photon_emit_cyclosynch: cudaMalloc");
    }
    ctx->emit_epoch += 1;
    cs_compact_kernel<<<nblocks, 256, 0, ctx->stream>>>(d, w);
    cs_weight_kernel<<<1, 1024, 0, ctx->stream>>>(d, w, ph_weight, max_photons, ctx->emit_epoch);
    if (int rc = check_launch(ctx, "cs_weight_kernel", 2)) { cleanup(); return rc; }
    double weight = 0;
    cudaMemcpyAsync(meta, w.meta, sizeof(meta), cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&weight, w.weight, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cleanup(); return fail(ctx, MCRAT_B200_ERR_CUDA, "photon_emit_cyclosynch: synchronize"); }
    if (meta[3] == 1) { cleanup(); return fail(ctx, MCRAT_B200_ERR_STATE, "photon_emit_cyclosynch: the black-body tail integral did not converge in 48 intervals"); }
    if (meta[3] == 2) { cleanup(); return fail(ctx, MCRAT_B200_ERR_STATE, "photon_emit_cyclosynch: the photon weight search did not settle in 400 passes"); }
    const int ph_tot = meta[1];
    if (num_emitted) *num_emitted = ph_tot;
    if (ph_weight_adjusted) *ph_weight_adjusted = weight;
    if (ph_tot > 0) {
        int nulls = 0;
        if (int rc = count_null_slots(ctx, w.meta, &nulls)) { cleanup(); return rc; }
        if (nulls < ph_tot) {
            const int need = ph_tot - nulls;
            int new_cap = d.cap * 2 > d.cap + need ? d.cap * 2 : d.cap + need;
            if (int rc = grow_list(ctx, new_cap)) { cleanup(); return rc; }
            if (int rc = count_null_slots(ctx, w.meta, &nulls)) { cleanup(); return rc; }
        }
        if (cudaMalloc((void **)&emitted, (size_t)ph_tot * sizeof(mcrat_photon)) != cudaSuccess) { cleanup(); return fail(ctx, MCRAT_B200_ERR_CUDA, "photon_emit_cyclosynch: cudaMalloc"); }
        // count_null_slots overwrote meta[0]: the fill kernel reads n_sel and ph_tot from it
        int m2[2] = {n_sel, ph_tot};
        cudaMemcpyAsync(w.meta, m2, sizeof(m2), cudaMemcpyHostToDevice, ctx->stream);
        cs_fill_kernel<<<ctx->d.replay ? 1 : grid_for(ctx, ph_tot, 256, 8), 256, 0, ctx->stream>>>(d, w, emitted, ctx->emit_epoch);
        place_photons_kernel<<<(d.cap + 255) / 256, 256, 0, ctx->stream>>>(d, emitted, ph_tot);
        if (int rc = check_launch(ctx, "cs_fill_kernel", 2)) { cleanup(); return rc; }
    }
    if (int rc = fetch_global(ctx)) { cleanup(); return rc; }
    cleanup();
    return device_error(ctx);
}

API int mcrat_b200_photon_emit_cyclosynch_single(mcrat_b200_ctx *ctx, int scatt_ph_index, int *new_photon_index)
{
    if (!ctx) return MCRAT_B200_ERR_ARG;
    if (int rc = need_ready(ctx)) return rc;
    if (int rc = need_single_shard(ctx, "photon_emit_cyclosynch_single")) return rc;
    if (scatt_ph_index < 0 || scatt_ph_index >= ctx->d.cap) return fail(ctx, MCRAT_B200_ERR_ARG, "photon_emit_cyclosynch_single: bad photon index");
    CK(cudaSetDevice(ctx->cfg.device));
    if (int rc = flush_pushes(ctx)) return rc;
    int *out = nullptr;
    CK(cudaMalloc((void **)&out, sizeof(int)));
    int slot = -1;
    for (int attempt = 0; attempt < 2; ++attempt) {
        cs_emit_single_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d, scatt_ph_index, out);
        if (int rc = check_launch(ctx, "cs_emit_single_kernel")) { cudaFree(out); return rc; }
        cudaMemcpyAsync(&slot, out, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaFree(out); return fail(ctx, MCRAT_B200_ERR_CUDA, "photon_emit_cyclosynch_single: synchronize"); }
        if (slot >= 0) break;
        // addToPhotonList doubles a list without null slots (Src/photons.c:117-129)
        if (int rc = grow_list(ctx, ctx->d.cap * 2 > ctx->d.cap + 1 ? ctx->d.cap * 2 : ctx->d.cap + 1)) { cudaFree(out); return rc; }
    }
    cudaFree(out);
    if (slot < 0) return fail(ctx, MCRAT_B200_ERR_STATE, "photon_emit_cyclosynch_single: no null slot after growing the list");
    if (new_photon_index) *new_photon_index = slot;
    if (int rc = fetch_global(ctx)) return rc;
    return device_error(ctx);
}

API int mcrat_b200_set_cs_rebin_params(mcrat_b200_ctx *ctx, double rebin_e_perc, double rebin_ang_deg, double rebin_ang_phi_deg)
{
    if (!ctx || !(rebin_e_perc > 0) || !(rebin_ang_deg > 0) || !(rebin_ang_phi_deg > 0)) return ctx ? fail(ctx, MCRAT_B200_ERR_ARG, "set_cs_rebin_params: values must be positive") : MCRAT_B200_ERR_ARG;
    ctx->cs_rebin_e_perc = rebin_e_perc;
    ctx->cs_rebin_ang = rebin_ang_deg;
    ctx->cs_rebin_ang_phi = rebin_ang_phi_deg;
    return MCRAT_B200_OK;
}

// ---- the device-resident frame loop, Src/mcrat.c:761-851 ---------------------------------------------
constexpr int MCRAT_B200_LOOP_FALLBACK = 1000; // internal: cooperative launch refused, use the streamed loop

__global__ void reset_protocol_kernel(DevCtx d)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < d.nshards; s += gridDim.x * blockDim.x) {
        ShardState &st = d.sh[s];
        st.arrive = 0;
        st.gen = 0;
        st.reloc_n = 0;
        st.halt = 0;
        st.reloc_heavy = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        d.gs->stream_work = 0;
        d.gs->stream_evt_ready = 0;
        d.gs->stream_halted = 0;
        d.gs->stream_state = 0;
    }
}

static int reset_protocol(mcrat_b200_ctx *ctx)
{
    reset_protocol_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d);
    return check_launch(ctx, "reset_protocol_kernel");
}

// geometry of the persistent launch: `bps` blocks per shard, all blocks resident
static void frame_loop_grid(const mcrat_b200_ctx *ctx, int &threads, int &bps, int &grid)
{
    const int S = ctx->d.nshards;
    const int cap256 = ctx->num_sms * ctx->occ_loop256;
    if (3 * S <= cap256) {
        // latency regime (room for at least two pass blocks per shard): a team of bps pass blocks + 1 event block
        // per sub-shard (frame_loop_kernel)
        threads = 256;
        bps = (ctx->d.shard_size + 255) / 256;
        if (bps > cap256 / S - 1) bps = cap256 / S - 1;
        if (bps > BLOCKMIN_CAP / S - 1) bps = BLOCKMIN_CAP / S - 1;
        if (bps < 1) bps = 1;
        grid = S * (bps + 1);
    } else {
        // throughput regime: one block per sub-shard (frame_loop_solo_kernel), bps = 0 marks it
        bps = 0;
        threads = (S <= cap256) ? 256 : 128;
        const int cap = (threads == 256) ? cap256 : ctx->num_sms * ctx->occ_loop128;
        grid = S < cap ? S : cap;
    }
}

// persistent loop for lists larger than L2 (frame_loop.cuh, PERSISTENT_STREAM): resident event blocks, one per sub-shard, on
// the second stream; pass blocks pulling items on the first.  All sub-shards' event blocks must be resident next to the
// pass blocks: one 128-thread block per SM beside four pass blocks.
// Event blocks of the persistent stream.  Fewer blocks leave more SMs with room for a fifth pass block, but a block serves
// its shards one after the other, and releases that queue up stall the pass stream: two shards per block for long lists
// (10^7 photons, 128 shards: 16 blocks 256 us per iteration, 32: 207, 43: 199, 64: 198, 128: 206; 5 x 10^6 photons, 64
// shards: 32 blocks 108 us, 64: 115), one for short ones, where an iteration is not much longer than an event
// (2.5 x 10^6 photons, 32 shards: 16 blocks 68 us, 32: 61; 1.25 x 10^6, 16 shards: 8 blocks 60 us, 16: 48).
static int frame_stream_evt_blocks(const mcrat_b200_ctx *ctx)
{
    const int S = ctx->d.nshards;
    int m = ctx->d.cap >= 4000000 ? 2 : 1;
    int E = (S + m - 1) / m;
    if (const char *e = getenv("MCRAT_B200_STREAM_EVT_BLOCKS"))
        if (atoi(e) > 0) E = atoi(e);
    if (E > S) E = S;
    if (E > ctx->num_sms) E = ctx->num_sms;
    while ((S + E - 1) / E > STREAM_EVT_SHARDS) ++E;
    return E;
}

// AUTO picks the persistent stream once the photon columns no longer fit in L2 beside the cells (1.2 x 10^6 photons x 97 B),
// for 16 or more sub-shards of at most 200 000 photons: with few, long shards a pass item waits for its shard's event more
// often than the missing launch boundaries save (10^7 photons, 16 shards: 272 us per iteration against 229 streamed).
// Measured against what AUTO used before: 1.25 x 10^6 photons / 16 shards 48 us (cooperative team 52), 2.5 x 10^6 / 32: 61
// (streamed 87), 5 x 10^6 / 64: 108 (127), 10^7 / 128: 197 (224).
constexpr int STREAM_AUTO_MIN_SHARDS = 16;
constexpr int STREAM_AUTO_MIN_PHOTONS = 1200000;
constexpr int STREAM_AUTO_MAX_SHARD_SIZE = 200000;

static bool frame_stream_fits(const mcrat_b200_ctx *ctx)
{
    return ctx->d.nshards >= 1 && ctx->d.nshards <= ctx->num_sms * STREAM_EVT_SHARDS && !ctx->d.cs && !ctx->d.replay &&
           frame_stream_evt_blocks(ctx) <= ctx->num_sms;
}

static int frame_stream_bps(const mcrat_b200_ctx *ctx)
{
    // Photons per thread and item.  Large items keep the item's fixed cost (wait, state, ticket: a few round trips to L2) a small
    // part of its time; small items give every resident block work when the list is short and let a shard's pass finish -- and
    // its event start -- sooner.  Measured (us per iteration; tools/gpu_r.sh, profiles/stream_coresidency_r02.txt):
    //   photons / shards   4      8      12     16     24     32
    //   10^7 / 128         224.8  210.0  202.6  199.3  197.2  198.0
    //   5 x 10^6 / 64      128.5  114.2  109.0  107.7  110.5  115
    //   2.5 x 10^6 / 32     74.1   69.9   67.8   70.9   83.6   94
    //   2.5 x 10^6 / 32 (one shard per event block):  4: 64.6   6: 61.2   8: 60.6   12: 63.6
    //   1.25 x 10^6 / 16 (one shard per event block): 3: 49.2   4: 47.6   6: 48.1    8: 50.5   12: 55.2
    int ppt = (int)(ctx->d.cap / 312500);
    if (ppt < 4) ppt = 4;
    if (ppt > 24) ppt = 24;
    if (const char *e = getenv("MCRAT_B200_STREAM_PPT"))
        if (atoi(e) > 0) ppt = atoi(e);
    int bps = (ctx->d.shard_size + PASS_THREADS * ppt - 1) / (PASS_THREADS * ppt);
    if (bps > BLOCKMIN_CAP / ctx->d.nshards - 1) bps = BLOCKMIN_CAP / ctx->d.nshards - 1;
    if (bps > EVT_THREADS_MANY - 1) bps = EVT_THREADS_MANY - 1; // the event block reads the minima with one load per thread
    if (bps < 1) bps = 1;
    return bps;
}

// Persistent loop for lists larger than L2 (frame_loop.cuh, PERSISTENT_STREAM): E resident event blocks on the second stream,
// pass blocks pulling items on the first.  The pass grid is what fits beside them: 4 blocks on an SM that holds an event
// block (48 K + 16 K registers), 5 elsewhere.  If the block scheduler places them differently some blocks simply start
// late (items are pulled, not assigned); if the two grids do not meet at all the start-up handshake calls the launch off.
static int launch_frame_stream(mcrat_b200_ctx *ctx)
{
    const int S = ctx->d.nshards;
    const int bps = frame_stream_bps(ctx);
    const int E = frame_stream_evt_blocks(ctx);
    if (getenv("MCRAT_B200_REFUSE_COOPERATIVE")) return MCRAT_B200_LOOP_FALLBACK;
    CK(cudaEventRecord(ctx->ev_join, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_join, 0));
    {
        Timed t(ctx, KC_EVENT);
        int pad_evt = 0;
        if (const char *e = getenv("MCRAT_B200_STREAM_PAD_EVT")) pad_evt = atoi(e);
        frame_stream_event_kernel<EVT_THREADS_MANY><<<E, EVT_THREADS_MANY, pad_evt, ctx->stream2>>>(ctx->d, bps);
        if (int rc = check_launch(ctx, "frame_stream_event_kernel")) return rc;
        // test hook: what a profiler that serialises kernels does to the pair (the event blocks never meet the pass blocks)
        if (getenv("MCRAT_B200_STREAM_SERIALIZE")) CK(cudaStreamSynchronize(ctx->stream2));
    }
    {
        Timed t(ctx, KC_PASS);
        int grid = E * STREAM_PASS_CTAS_PER_SM + (ctx->num_sms - E) * MCRAT_PASS_MINB;
        if (const char *e = getenv("MCRAT_B200_STREAM_PASS_BLOCKS"))
            if (atoi(e) > 0) grid = atoi(e);
        int pad_pass = STREAM_PASS_SMEM_PAD;
        if (const char *e = getenv("MCRAT_B200_STREAM_PAD_PASS")) pad_pass = atoi(e);
        frame_stream_pass_kernel<<<grid, PASS_THREADS, pad_pass, ctx->stream>>>(ctx->d, bps, E);
        if (int rc = check_launch(ctx, "frame_stream_pass_kernel")) return rc;
    }
    CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    // did the two grids meet?  (a profiler that serialises kernels, a device shared with another tenant: they may not)
    if (int rc = fetch_global(ctx)) return rc;
    if (getenv("MCRAT_B200_DEBUG"))
        fprintf(stderr, "[mcrat_b200] persistent stream: S=%d E=%d bps=%d pass grid=%d -> event blocks seen %d, state %d, items pulled %llu, shards halted %d\n",
                S, E, bps, E * STREAM_PASS_CTAS_PER_SM + (ctx->num_sms - E) * MCRAT_PASS_MINB, ctx->gs_host->stream_evt_ready,
                ctx->gs_host->stream_state, ctx->gs_host->stream_work, ctx->gs_host->stream_halted);
    if (ctx->gs_host->stream_state == 2 /* STREAM_ABORT */) return MCRAT_B200_LOOP_FALLBACK;
    return MCRAT_B200_OK;
}

// The team of a sub-shard as one thread-block cluster (frame_loop_cluster_kernel): the largest team of at most 16 blocks for
// which every sub-shard's cluster is resident at once.  Returns MCRAT_B200_OK after a launch, MCRAT_B200_LOOP_FALLBACK when the
// cooperative team kernel should run instead (no such team, clusters unavailable).
static int launch_frame_cluster(mcrat_b200_ctx *ctx, int bps_wanted)
{
    const int S = ctx->d.nshards;
    if (ctx->cluster_state < 0) return MCRAT_B200_LOOP_FALLBACK;
    int bps = bps_wanted < CLUSTER_TEAM_MAX - 1 ? bps_wanted : CLUSTER_TEAM_MAX - 1;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(256);
    cfg.stream = ctx->stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    for (; bps >= 2; --bps) {
        const int team = bps + 1;
        attr[0].val.clusterDim.x = team;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.gridDim = dim3(S * team);
        int nclusters = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, frame_loop_cluster_kernel<256>, &cfg);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            ctx->cluster_state = -1;
            return MCRAT_B200_LOOP_FALLBACK;
        }
        if (nclusters >= S) break;
    }
    if (bps < 2) return MCRAT_B200_LOOP_FALLBACK; // too many sub-shards for resident clusters
    Timed t(ctx, KC_EVENT);
    cudaError_t e = cudaLaunchKernelEx(&cfg, frame_loop_cluster_kernel<256>, ctx->d, bps);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        ctx->cluster_state = -1;
        return MCRAT_B200_LOOP_FALLBACK;
    }
    ctx->cluster_state = 1;
    return check_launch(ctx, "frame_loop_cluster_kernel");
}

// Launch geometry of the persistent stream for a list of `list_capacity` photons in `num_shards` sub-shards on a device with
// `num_sms` SMs -- the same functions the launch uses, without a device (tests/test_stream_geometry.py sweeps it for the
// bounds the kernels rely on).  fits: frame_stream_fits; auto_picks: what AUTO would do (1 = persistent stream).
API int mcrat_b200_debug_stream_geometry(int list_capacity, int num_shards, int num_sms, int *event_blocks, int *bps, int *pass_grid,
                                         int *shards_out, int *fits, int *auto_picks)
{
    if (list_capacity < 1 || num_shards < 1 || num_shards > MAX_SHARDS || num_sms < 1) return MCRAT_B200_ERR_ARG;
    mcrat_b200_ctx *ctx = new mcrat_b200_ctx();
    int S = num_shards > list_capacity ? list_capacity : num_shards; // layout_shards()
    const int size = (list_capacity + S - 1) / S;
    S = (list_capacity + size - 1) / size;
    ctx->d.nshards = S;
    ctx->d.shard_size = size;
    ctx->d.cap = list_capacity;
    ctx->d.cs = 0;
    ctx->d.replay = 0;
    ctx->num_sms = num_sms;
    const int E = frame_stream_evt_blocks(ctx);
    if (event_blocks) *event_blocks = E;
    if (bps) *bps = frame_stream_bps(ctx);
    if (pass_grid) *pass_grid = E * STREAM_PASS_CTAS_PER_SM + (num_sms - E) * MCRAT_PASS_MINB;
    if (shards_out) *shards_out = S;
    const bool f = frame_stream_fits(ctx);
    if (fits) *fits = f ? 1 : 0;
    if (auto_picks)
        *auto_picks = (f && list_capacity >= STREAM_AUTO_MIN_PHOTONS && S >= STREAM_AUTO_MIN_SHARDS && size <= STREAM_AUTO_MAX_SHARD_SIZE) ? 1 : 0;
    delete ctx;
    return MCRAT_B200_OK;
}

static int launch_frame_loop(mcrat_b200_ctx *ctx, bool stream = false)
{
    if (stream) return launch_frame_stream(ctx);
    int threads, bps, grid;
    frame_loop_grid(ctx, threads, bps, grid);
    if (grid < 1) return fail(ctx, MCRAT_B200_ERR_STATE, "frame_loop_kernel does not fit on this device");
    void *args[2] = {(void *)&ctx->d, (void *)&bps};
    if (getenv("MCRAT_B200_REFUSE_COOPERATIVE")) return MCRAT_B200_LOOP_FALLBACK; // test hook for the hand-over below
    if (bps >= 2) {
        const int rc = launch_frame_cluster(ctx, bps);
        if (rc != MCRAT_B200_LOOP_FALLBACK) return rc;
    }
    Timed t(ctx, KC_EVENT);
    cudaError_t e;
    if (bps == 0) { // independent blocks: an ordinary launch
        if (threads == 128)
            frame_loop_solo_kernel<128><<<grid, 128, 0, ctx->stream>>>(ctx->d);
        else
            frame_loop_solo_kernel<256><<<grid, 256, 0, ctx->stream>>>(ctx->d);
        return check_launch(ctx, "frame_loop_solo_kernel");
    }
    e = cudaLaunchCooperativeKernel((const void *)frame_loop_kernel<256>, dim3(grid), dim3(256), args, 0, ctx->stream);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported || e == cudaErrorLaunchOutOfResources) {
        // the device cannot hold the grid (MPS share, another tenant, no cooperative launch): the streamed loop needs nothing special
        (void)cudaGetLastError();
        return MCRAT_B200_LOOP_FALLBACK;
    }
    if (e != cudaSuccess) {
        ctx->err = std::string("frame_loop_kernel: ") + cudaGetErrorString(e);
        return MCRAT_B200_ERR_CUDA;
    }
    return check_launch(ctx, "frame_loop_kernel");
}

API int mcrat_b200_set_recheck_skip(mcrat_b200_ctx *ctx, int mode)
{
    if (!ctx) return MCRAT_B200_ERR_ARG;
    if (mode < 0 || mode > 2) return fail(ctx, MCRAT_B200_ERR_ARG, "set_recheck_skip: 0 (off), 1 (on), 2 (on, verified)");
    ctx->d.recheck_skip = mode;
    return MCRAT_B200_OK;
}

API int mcrat_b200_set_loop_mode(mcrat_b200_ctx *ctx, int mode)
{
    if (!ctx) return MCRAT_B200_ERR_ARG;
    if (mode != MCRAT_B200_LOOP_AUTO && mode != MCRAT_B200_LOOP_STREAMED && mode != MCRAT_B200_LOOP_PERSISTENT &&
        mode != MCRAT_B200_LOOP_STREAMED_GLOBAL && mode != MCRAT_B200_LOOP_PERSISTENT_STREAM)
        return fail(ctx, MCRAT_B200_ERR_ARG, "set_loop_mode: unknown mode");
    ctx->loop_mode = mode;
    return MCRAT_B200_OK;
}

static void fill_stats(const ShardState &s, const ShardState &b, mcrat_b200_frame_stats *o)
{
    o->iterations = s.iters_done;
    o->scatterings = s.scatt_cnt - b.scatt_cnt;
    o->relocations = s.reloc_total - b.reloc_total;
    o->photon_slots = s.slots - b.slots;
    o->cell_evals = 0;
    o->box_evals = 0;
    o->time_now = s.time_now;
    o->last_time_step = s.last_time_step;
    o->last_scattered_index = s.last_scattered_idx;
    o->not_found = 0;
    o->cs_host_pending = s.pause_cs;
    o->error = 0;
    o->cs_emitted = 0;
    o->scatt_cyclosynch_num_ph = 0;
    o->cs_comptonized_weight = 0;
    o->ref_equiv_evals = 0;
}

API int mcrat_b200_run_frame(mcrat_b200_ctx *ctx, double time_now, double remaining_time, long long max_iters, int sw,
                             mcrat_b200_frame_stats *stats)
{
    if (!ctx || !stats) return ctx ? fail(ctx, MCRAT_B200_ERR_ARG, "run_frame: null stats") : MCRAT_B200_ERR_ARG;
    if (int rc = need_ready(ctx)) return rc;
    CK(cudaSetDevice(ctx->cfg.device));
    if (int rc = fetch_state(ctx)) return rc;
    const int S = ctx->d.nshards;
    std::vector<ShardState> before(ctx->sh_host.begin(), ctx->sh_host.begin() + S);
    const GlobalState gbefore = *ctx->gs_host;
    if (int rc = reset_loop(ctx, 1, time_now, remaining_time, max_iters)) return rc;
    sw = sw ? 1 : 0;
    const bool fused = !ctx->d.replay;
    // one iteration of every running shard as four stream-ordered launches
    auto streamed_iteration = [&]() -> int {
        int parity = 0, nb = ctx->d.blocks_per_shard;
        if (fused) {
            if (int rc = launch_locate<true>(ctx, sw, parity)) return rc;
        } else {
            if (int rc = launch_locate<false>(ctx, sw, parity)) return rc;
            if (int rc = launch_mfp_unfused(ctx, nb)) return rc;
        }
        {
            Timed t(ctx, KC_EVENT);
            if (S >= 64)
                event_kernel<EVT_THREADS_MANY><<<S, EVT_THREADS_MANY, 0, ctx->stream>>>(ctx->d, parity, nb, 0, 0.0);
            else
                event_kernel<EVT_THREADS><<<S, EVT_THREADS, 0, ctx->stream>>>(ctx->d, parity, nb, 0, 0.0);
            if (int rc = check_launch(ctx, "event_kernel")) return rc;
        }
        sw = 0; // Src/mcrat.c:773
        return MCRAT_B200_OK;
    };
    // Streamed loop, interleaved halves (see frame_loop.cuh): the shards [0, S/2) run on the context's stream, the shards
    // [S/2, S) on a second one; the pass of a half waits for the other half's pass (so the two passes never share the HBM
    // pipe and the halves stay half a period apart) and, through stream order, for its own half's event.
    const bool local_ok = fused && !ctx->d.cs && S >= 2 && ctx->loop_mode != MCRAT_B200_LOOP_STREAMED_GLOBAL;
    const int nsh[2] = {S - S / 2, S / 2}, sh0[2] = {0, S - S / 2};
    int bps_local[2];
    for (int h = 0; h < 2; ++h) {
        int b = (MCRAT_PASS_LOCAL_CTAS_PER_SM * ctx->num_sms) / (nsh[h] > 0 ? nsh[h] : 1);
        const int need = (ctx->d.shard_size + PASS_THREADS - 1) / PASS_THREADS;
        if (b > need) b = need;
        if (b > BLOCKMIN_CAP / S) b = BLOCKMIN_CAP / S;
        if (b < 1) b = 1;
        bps_local[h] = b;
    }
    if (bps_local[1] < bps_local[0]) bps_local[0] = bps_local[1]; // one layout of the block minima for both halves
    bps_local[1] = bps_local[0];
    auto local_iterations = [&](long long count) -> int {
        const bool serial = ctx->cfg.profile != 0; // per-kernel timing: everything on one stream
        cudaStream_t st[2] = {ctx->stream, serial ? ctx->stream : ctx->stream2};
        if (!serial) {
            CK(cudaEventRecord(ctx->ev_join, ctx->stream));
            CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_join, 0));
            CK(cudaEventRecord(ctx->ev_pass[1], ctx->stream2)); // "the other half's pass is done" for the very first pass
        }
        for (long long it = 0; it < count; ++it)
            for (int h = 0; h < 2; ++h) {
                if (!serial) CK(cudaStreamWaitEvent(st[h], ctx->ev_pass[h ^ 1], 0));
                {
                    Timed t(ctx, KC_PASS);
                    pass_local_kernel<<<nsh[h] * bps_local[h], PASS_THREADS, 0, st[h]>>>(ctx->d, sh0[h], bps_local[h]);
                    if (int rc = check_launch(ctx, "pass_local_kernel")) return rc;
                }
                if (!serial) CK(cudaEventRecord(ctx->ev_pass[h], st[h]));
                {
                    Timed t(ctx, KC_EVENT);
                    if (nsh[h] >= 32)
                        event_local_kernel<EVT_THREADS_MANY><<<nsh[h], EVT_THREADS_MANY, 0, st[h]>>>(ctx->d, sh0[h], bps_local[h]);
                    else
                        event_local_kernel<EVT_THREADS><<<nsh[h], EVT_THREADS, 0, st[h]>>>(ctx->d, sh0[h], bps_local[h]);
                    if (int rc = check_launch(ctx, "event_local_kernel")) return rc;
                }
            }
        if (!serial) {
            CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
        }
        return MCRAT_B200_OK;
    };
    const bool auto_stream = ctx->loop_mode == MCRAT_B200_LOOP_AUTO && ctx->d.cap >= STREAM_AUTO_MIN_PHOTONS &&
                             S >= STREAM_AUTO_MIN_SHARDS && ctx->d.shard_size <= STREAM_AUTO_MAX_SHARD_SIZE && frame_stream_fits(ctx);
    bool persistent = fused && !ctx->cfg.profile &&
                      (ctx->loop_mode == MCRAT_B200_LOOP_PERSISTENT ||
                       (ctx->loop_mode == MCRAT_B200_LOOP_AUTO && ctx->d.cap <= PERSISTENT_MAX_PHOTONS && !auto_stream));
    // lists larger than L2: the persistent stream of pass items beside resident event blocks, if every sub-shard's event
    // block fits on the device at once; else the interleaved streamed loop
    const bool pstream = fused && !ctx->cfg.profile && !persistent && frame_stream_fits(ctx) &&
                         (ctx->loop_mode == MCRAT_B200_LOOP_PERSISTENT_STREAM || auto_stream);
    if (pstream) persistent = true;
    long long streamed_done = 0; // iterations already launched when the persistent loop hands over for good
    if (persistent) {
        // the first iteration of a new hydro frame re-locates every photon: that is K1's job
        long long launched = 0;
        if (sw == 1 && max_iters != 0) {
            if (int rc = streamed_iteration()) return rc;
            launched++;
        }
        for (;;) {
            if (int rc = launch_frame_loop(ctx, pstream)) {
                if (rc != MCRAT_B200_LOOP_FALLBACK) return rc;
                ctx->loop_mode = MCRAT_B200_LOOP_STREAMED; // for the rest of this context's life
                persistent = false;
                streamed_done = launched;
                break;
            }
            if (int rc = fetch_state(ctx)) return rc;
            if (ctx->gs_host->error) break;
            bool heavy = false, running = false;
            long long batch = 32; // iterations the slowest running shard may still do, capped
            for (int k = 0; k < S; ++k) {
                const ShardState &h = ctx->sh_host[k];
                const bool stopped = h.done || h.pause_cs || (max_iters >= 0 && h.iters_done >= max_iters);
                if (!stopped) running = true;
                if (!stopped && h.reloc_heavy) heavy = true;
                if (!stopped && max_iters >= 0 && max_iters - h.iters_done < batch) batch = max_iters - h.iters_done;
            }
            if (!running) break;
            if (!heavy) return fail(ctx, MCRAT_B200_ERR_STATE, "frame_loop_kernel returned with running shards");
            // many photons change cell per iteration (optically thin flow): K1b / K1c serve that better
            for (long long b = 0; b < batch; ++b)
                if (int rc = streamed_iteration()) return rc;
            if (int rc = reset_protocol(ctx)) return rc;
            (void)launched;
        }
    }
    if (!persistent) {
        // iterations are enqueued in batches; kernels of a shard past its stop condition return at once
        int batch = 1;
        long long launched = streamed_done;
        bool heavy = false; // many photons change cell per iteration: the grid-wide K1b / K1c serve that better
        for (;;) {
            if (local_ok && sw == 0 && !heavy) {
                if (int rc = local_iterations(batch)) return rc;
                launched += batch;
            } else {
                for (int b = 0; b < batch; ++b) {
                    if (int rc = streamed_iteration()) return rc;
                    launched++;
                }
            }
            if (int rc = fetch_global(ctx)) return rc;
            const GlobalState &g = *ctx->gs_host;
            if (g.error || g.n_stopped >= S) break;
            if (g.reloc_heavy_any) heavy = true;
            if (batch < 64) batch *= 2;
            if (max_iters >= 0 && launched + batch > max_iters) batch = (int)(max_iters - launched);
            if (batch < 1) batch = 1;
        }
    }
    if (int rc = fetch_state(ctx)) return rc;
    // aggregate over the sub-shards: counters add up, the clock reported is shard 0's
    fill_stats(ctx->sh_host[0], before[0], stats);
    for (int s = 1; s < S; ++s) {
        mcrat_b200_frame_stats t;
        fill_stats(ctx->sh_host[s], before[s], &t);
        if (t.iterations > stats->iterations) stats->iterations = t.iterations;
        stats->scatterings += t.scatterings;
        stats->relocations += t.relocations;
        stats->photon_slots += t.photon_slots;
        stats->cs_host_pending |= t.cs_host_pending;
    }
    stats->cell_evals = ctx->gs_host->cell_evals - gbefore.cell_evals;
    stats->box_evals = ctx->gs_host->box_evals - gbefore.box_evals;
    stats->ref_equiv_evals = ctx->gs_host->ref_equiv_evals - gbefore.ref_equiv_evals;
    stats->not_found = ctx->gs_host->not_found - gbefore.not_found;
    stats->error = ctx->gs_host->error;
    stats->cs_emitted = ctx->gs_host->cs_emitted - gbefore.cs_emitted;
    stats->scatt_cyclosynch_num_ph = ctx->gs_host->cs_scatt_num;
    stats->cs_comptonized_weight = ctx->gs_host->cs_comptonized_w - gbefore.cs_comptonized_w;
    ctx->last_nb_mfp = 0;
    return device_error(ctx);
}

// the driver's cyclo-synchrotron counters (Src/mcrat.c:803, 820, 857): rebin threshold and the running
// number of scattered pool photons; the loop pauses with cs_host_pending == 2 when a rebin is due
API int mcrat_b200_set_cs_limits(mcrat_b200_ctx *ctx, int max_photons, int scatt_cyclosynch_num_ph)
{
    if (!ctx) return MCRAT_B200_ERR_ARG;
    if (int rc = fetch_global(ctx)) return rc;
    ctx->gs_host->cs_max_photons = max_photons;
    ctx->gs_host->cs_scatt_num = scatt_cyclosynch_num_ph;
    CK(cudaMemcpyAsync(ctx->d.gs, ctx->gs_host, sizeof(GlobalState), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MCRAT_B200_OK;
}

// calcCyclosynchRLimits, Src/mc_cyclosynch.h:84, Src/mc_cyclosynch.c:225-242 (pure host arithmetic)
API double mcrat_b200_calc_cyclosynch_r_limits(int frame_scatt, int frame_inj, double fps, double r_inj, const char *min_or_max)
{
    const double c_light = 2.99792458e10;
    double val = r_inj;
    if (min_or_max && strcmp(min_or_max, "min") == 0)
        val += (c_light * (frame_scatt - frame_inj) / fps - 0.5 * c_light / fps);
    else
        val += (c_light * (frame_scatt - frame_inj) / fps + 0.5 * c_light / fps);
    return val;
}

API int mcrat_b200_get_shard_stats(mcrat_b200_ctx *ctx, int shard, mcrat_b200_frame_stats *stats, int *first_slot,
                                   int *num_slots)
{
    if (!ctx || !stats || shard < 0 || shard >= ctx->d.nshards) return ctx ? fail(ctx, MCRAT_B200_ERR_ARG, "get_shard_stats: bad shard") : MCRAT_B200_ERR_ARG;
    if (int rc = fetch_state(ctx)) return rc;
    ShardState zero;
    memset(&zero, 0, sizeof(zero));
    fill_stats(ctx->sh_host[shard], zero, stats);
    if (first_slot) *first_slot = ctx->sh_host[shard].first;
    if (num_slots) *num_slots = ctx->sh_host[shard].count;
    return MCRAT_B200_OK;
}

// ---- measurement -------------------------------------------------------------------------------------
// photons for which findContainingBlock found no cell since the last call (Src/geometry.c:373-388 logs each of them)
API int mcrat_b200_get_not_found(mcrat_b200_ctx *ctx, int max_entries, int *slots, double *hydro_coords, int *n_total)
{
    if (!ctx || max_entries < 0) return ctx ? fail(ctx, MCRAT_B200_ERR_ARG, "get_not_found: bad argument") : MCRAT_B200_ERR_ARG;
    if (int rc = fetch_global(ctx)) return rc;
    const GlobalState &g = *ctx->gs_host;
    const int n = std::min(std::min(g.nf_logged, NF_LOG_CAP), max_entries);
    for (int k = 0; k < n; ++k) {
        if (slots) slots[k] = g.nf_slot[k];
        if (hydro_coords) memcpy(hydro_coords + 3 * k, g.nf_h + 3 * k, 3 * sizeof(double));
    }
    if (n_total) *n_total = g.nf_logged;
    if (g.nf_logged) {
        ctx->gs_host->nf_logged = 0;
        CK(cudaMemsetAsync(&ctx->d.gs->nf_logged, 0, sizeof(int), ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return n;
}

API int mcrat_b200_get_kernel_times(mcrat_b200_ctx *ctx, mcrat_b200_kernel_times *out, int reset)
{
    if (!ctx || !out) return MCRAT_B200_ERR_ARG;
    *out = ctx->times;
    if (reset) memset(&ctx->times, 0, sizeof(ctx->times));
    return MCRAT_B200_OK;
}

API int mcrat_b200_set_profile(mcrat_b200_ctx *ctx, int on)
{
    if (!ctx) return MCRAT_B200_ERR_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->cfg.profile = on ? 1 : 0;
    return MCRAT_B200_OK;
}

#ifdef MCRAT_TIMING
API int mcrat_b200_debug_counters(mcrat_b200_ctx *ctx, long long *out32, int reset)
{
    if (int rc = fetch_global(ctx)) return rc;
    memcpy(out32, ctx->gs_host->dbg, sizeof(long long) * 32);
    if (reset) {
        memset(ctx->gs_host->dbg, 0, sizeof(long long) * 32);
        CK(cudaMemcpyAsync(ctx->d.gs, ctx->gs_host, sizeof(GlobalState), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return MCRAT_B200_OK;
}
#endif

API long long mcrat_b200_launch_count(const mcrat_b200_ctx *ctx) { return ctx ? ctx->launches : 0; }

API int mcrat_b200_rescan_all(mcrat_b200_ctx *ctx, long long *cell_evals, float *elapsed_ms)
{
    if (!ctx) return MCRAT_B200_ERR_ARG;
    if (int rc = need_ready(ctx)) return rc;
    if (int rc = flush_pushes(ctx)) return rc;
    if (int rc = reset_loop(ctx, 0, 0, 0, -1)) return rc;
    if (int rc = fetch_global(ctx)) return rc;
    const long long before = ctx->gs_host->cell_evals;
    const int parity = ctx->pass_parity;
    ctx->pass_parity ^= 1;
    const int nbp = ctx->d.nshards * ctx->d.blocks_per_shard;
    pass_kernel<false><<<nbp, PASS_THREADS, 0, ctx->stream>>>(ctx->d, 1, parity);
    if (int rc = check_launch(ctx, "pass_kernel")) return rc;
    int grid, tpc;
    scan_grid(ctx, ctx->d.cap, grid, tpc);
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    if (ctx->d.dims == D_THREE)
        scan_kernel<1><<<grid, SCAN_THREADS, scan_smem_bytes(1), ctx->stream>>>(ctx->d, parity, tpc);
    else
        scan_kernel<0><<<grid, SCAN_THREADS, scan_smem_bytes(0), ctx->stream>>>(ctx->d, parity, tpc);
    if (int rc = check_launch(ctx, "scan_kernel")) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    const int nbf = grid_for(ctx, ctx->d.cap, FIN_THREADS, 8);
    finish_kernel<false><<<nbf, FIN_THREADS, 0, ctx->stream>>>(ctx->d, 1, parity);
    if (int rc = check_launch(ctx, "finish_kernel")) return rc;
    if (int rc = fetch_global(ctx)) return rc;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (elapsed_ms) *elapsed_ms = ms;
    if (cell_evals) *cell_evals = ctx->gs_host->cell_evals - before;
    ctx->times.scan_ms += ms;
    ctx->times.scan_launches++;
    return device_error(ctx);
}

// createHotCrossSection, Src/hot_x_section.c:82-206, on the device; the table is also installed
// in the context (as mcrat_b200_set_thermal_table would)
API int mcrat_b200_build_thermal_table(mcrat_b200_ctx *ctx, long long calls, uint64_t seed, double *table_out,
                                       float *elapsed_ms)
{
    if (!ctx || calls < 1) return ctx ? fail(ctx, MCRAT_B200_ERR_ARG, "build_thermal_table: calls >= 1") : MCRAT_B200_ERR_ARG;
    CK(cudaSetDevice(ctx->cfg.device));
    const int npts = (N_PH_E + 1) * (N_T + 1);
    double *tab = nullptr;
    CK(cudaMalloc((void **)&tab, npts * sizeof(double)));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    hot_table_kernel<<<npts, 256, 0, ctx->stream>>>(tab, calls, (uint32_t)seed ^ 0x4D435261u, (uint32_t)(seed >> 32), 0);
    if (int rc = check_launch(ctx, "hot_table_kernel")) {
        cudaFree(tab);
        return rc;
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    std::vector<double> host(npts);
    CK(cudaMemcpyAsync(host.data(), tab, npts * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(tab);
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (elapsed_ms) *elapsed_ms = ms;
    for (int k = 0; k < npts; ++k)
        if (host[k] != host[k]) return fail(ctx, MCRAT_B200_ERR_STATE, "NaN in the hot cross-section table (Src/hot_x_section.c:97-103)");
    if (table_out) memcpy(table_out, host.data(), npts * sizeof(double));
    return mcrat_b200_set_thermal_table(ctx, host.data());
}

API int mcrat_b200_measure_fp64_peak(mcrat_b200_ctx *ctx, double *ginstr_per_s)
{
    if (!ctx || !ginstr_per_s) return MCRAT_B200_ERR_ARG;
    double *out = nullptr;
    CK(cudaMalloc((void **)&out, sizeof(double)));
    const int iters = 8192, blocks = ctx->num_sms * 8, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(out, iters, 1.000001 + rep);
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaFree(out);
    // per thread and iteration: 16 DFMA (one FP64-pipe instruction each)
    double instr = (double)blocks * threads * (double)iters * 16.0;
    *ginstr_per_s = instr / (best * 1e-3) / 1e9;
    return MCRAT_B200_OK;
}

// self-test of div_by_c against the hardware-rounded division: n Philox-drawn significands at each of the binary
// exponents -60 ... +60 around 1 (plus the range ends, where the true division is taken anyway)
__global__ void div_by_c_check_kernel(long long n, uint32_t seed, unsigned long long *bad)
{
    unsigned long long mine = 0;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        double a, b;
        philox_doubles((uint32_t)k, (uint32_t)(k >> 32), 0u, 7u, seed, 0x64697643u, a, b);
        const int e = (int)(k % 121) - 60;
        const double xs[4] = {ldexp(1.0 + a, e), -ldexp(1.0 + b, e), ldexp(1.0 + a, 8 * e), ldexp(1.0 + b, -1000 + e)};
        for (int q = 0; q < 4; ++q) {
            const double x = xs[q];
            if (__double_as_longlong(div_by_c(x)) != __double_as_longlong(x / C_LIGHT)) mine++;
        }
    }
    if (mine) atomicAdd(bad, mine);
}

API int mcrat_b200_selftest_div_by_c(mcrat_b200_ctx *ctx, long long n, unsigned seed, long long *mismatches)
{
    if (!ctx || !mismatches || n < 0) return MCRAT_B200_ERR_ARG;
    unsigned long long *bad = nullptr;
    CK(cudaMalloc((void **)&bad, sizeof(unsigned long long)));
    CK(cudaMemsetAsync(bad, 0, sizeof(unsigned long long), ctx->stream));
    div_by_c_check_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(n, seed, bad);
    ctx->launches++;
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, bad, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(bad);
    *mismatches = (long long)h;
    return MCRAT_B200_OK;
}

API int mcrat_b200_measure_hbm_peak(mcrat_b200_ctx *ctx, double *gb_per_s)
{
    if (!ctx || !gb_per_s) return MCRAT_B200_ERR_ARG;
    const size_t n = (size_t)1 << 26; // 2 GiB per buffer of double4: larger than L2
    double4 *a = nullptr, *b = nullptr;
    CK(cudaMalloc((void **)&a, n * sizeof(double4)));
    if (cudaMalloc((void **)&b, n * sizeof(double4)) != cudaSuccess) {
        cudaFree(a);
        return fail(ctx, MCRAT_B200_ERR_CUDA, "cudaMalloc (hbm probe)");
    }
    CK(cudaMemsetAsync(a, 0, n * sizeof(double4), ctx->stream));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        copy_kernel<<<ctx->num_sms * 16, 256, 0, ctx->stream>>>(a, b, n);
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaFree(a);
    cudaFree(b);
    *gb_per_s = 2.0 * (double)n * sizeof(double4) / (best * 1e-3) / 1e9;
    return MCRAT_B200_OK;
}

// the reference's MPI exchanges either side of the frame loop, over NCCL
#include "comm.cuh"
