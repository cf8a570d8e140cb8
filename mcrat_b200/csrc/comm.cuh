// comm.cuh -- the reference's MPI exchanges either side of the frame loop, over NCCL (NVLink / NVSwitch).
// Part of the single translation unit mcrat_b200.cu (included there, last); not a stand-alone header.
//
// The frame loop itself needs no exchange: a GPU's sub-shards are MPI ranks of the reference (Src/mcrat.c:139-164,
// 457-479; no MPI call between :609 and :924).  What the reference does exchange, and what this file provides:
//   * the hot cross-section table: rank 0 builds / reads it, MPI_Bcast to the others (Src/hot_x_section.c:717).  Here
//     either the same broadcast, or -- K7 being a device kernel -- every GPU integrates 1/N of the 17 901 points and
//     one all-gather puts the table together (the points' Philox streams are keyed by the point, so the table is the
//     one a single GPU builds, bit for bit);
//   * per-frame counters (the north star's "load-balance counts"): all-reduce of the frame statistics, all-gather of
//     the per-rank photon counts (what merge.c:784-790 gathers from the per-rank files);
//   * the merged photon output: MPI_Allgatherv of every photon column (Src/merge.c:840-876).  Here the records are
//     packed on the device and sent GPU-to-GPU into one buffer in rank order.
// NCCL is bound at run time (dlopen), so libmcrat_b200.so itself loads on hosts without it; a comm call then fails
// with MCRAT_B200_ERR_STATE.  The unique id travels over whatever the host already has (MPI_Bcast in mcrat.c, the
// torch.distributed store in bench.py).
#pragma once
#include <dlfcn.h>
#include <nccl.h> // types and prototypes only

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    const char *(*GetErrorString)(ncclResult_t);
    ncclResult_t (*GetVersion)(int *);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)(void);
    ncclResult_t (*GroupEnd)(void);
    bool ok;
    std::string why;
};

static NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    // a copy the process already holds (torch's bundled NCCL in the harness, the MPI host's own) wins: two NCCL
    // instances in one process would each open their own transport state
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen(getenv("MCRAT_B200_NCCL") ? getenv("MCRAT_B200_NCCL") : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        api.why = std::string("NCCL not found: ") + (dlerror() ? dlerror() : "dlopen failed");
        return &api;
    }
    bool all = true;
#define BIND(field, sym)                                      \
    do {                                                      \
        *(void **)(&api.field) = dlsym(h, sym);               \
        if (!api.field) {                                     \
            all = false;                                      \
            api.why = std::string("NCCL symbol missing: ") + sym; \
        }                                                     \
    } while (0)
    BIND(GetUniqueId, "ncclGetUniqueId");
    BIND(CommInitRank, "ncclCommInitRank");
    BIND(CommDestroy, "ncclCommDestroy");
    BIND(GetErrorString, "ncclGetErrorString");
    BIND(GetVersion, "ncclGetVersion");
    BIND(AllReduce, "ncclAllReduce");
    BIND(Broadcast, "ncclBroadcast");
    BIND(AllGather, "ncclAllGather");
    BIND(Send, "ncclSend");
    BIND(Recv, "ncclRecv");
    BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd");
#undef BIND
    api.ok = all;
    return &api;
}

struct mcrat_b200_comm {
    mcrat_b200_ctx *ctx;
    ncclComm_t nccl;
    int nranks, rank;
    long long *words_dev;  // staging for the small reductions: 64 words + one per rank
    long long *words_host; // pinned
    long long collectives; // NCCL calls issued
};

constexpr int COMM_WORDS = 64;

#define NK(call)                                                                                        \
    do {                                                                                                \
        ncclResult_t r__ = (call);                                                                      \
        if (r__ != ncclSuccess) {                                                                       \
            ctx->err = std::string(#call) + ": " + nccl_api()->GetErrorString(r__);                     \
            return MCRAT_B200_ERR_CUDA;                                                                 \
        }                                                                                               \
    } while (0)

// photons printPhotons would write (weight != 0, Src/mcrat_io.c:150-160) and null slots of a slot range
__global__ void count_output_kernel(DevCtx d, int n, long long *out)
{
    int cnt = 0, nulls = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        cnt += d.ph.weight[i] != 0;
        nulls += d.ph.type[i] == 'N'; // NULL_PHOTON, Src/mcrat.h:57
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        nulls += __shfl_xor_sync(0xffffffffu, nulls, off);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cnt) atomicAdd((unsigned long long *)&out[0], (unsigned long long)cnt);
        if (nulls) atomicAdd((unsigned long long *)&out[1], (unsigned long long)nulls);
    }
}

API int mcrat_b200_comm_unique_id(unsigned char *id, size_t len)
{
    NcclApi *a = nccl_api();
    if (!id || len < sizeof(ncclUniqueId)) return MCRAT_B200_ERR_ARG;
    if (!a->ok) {
        g_create_error = a->why;
        return MCRAT_B200_ERR_STATE;
    }
    ncclUniqueId u;
    ncclResult_t r = a->GetUniqueId(&u);
    if (r != ncclSuccess) {
        g_create_error = std::string("ncclGetUniqueId: ") + a->GetErrorString(r);
        return MCRAT_B200_ERR_CUDA;
    }
    memset(id, 0, len);
    memcpy(id, &u, sizeof(u));
    return MCRAT_B200_OK;
}

API int mcrat_b200_comm_nccl_version(void)
{
    NcclApi *a = nccl_api();
    int v = 0;
    if (!a->ok || a->GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

API int mcrat_b200_comm_create(mcrat_b200_ctx *ctx, int nranks, int rank, const unsigned char *id, size_t len,
                               mcrat_b200_comm **out)
{
    if (!ctx || !out || !id || len < sizeof(ncclUniqueId) || nranks < 1 || rank < 0 || rank >= nranks)
        return ctx ? fail(ctx, MCRAT_B200_ERR_ARG, "comm_create: bad argument") : MCRAT_B200_ERR_ARG;
    NcclApi *a = nccl_api();
    if (!a->ok) return fail(ctx, MCRAT_B200_ERR_STATE, a->why.c_str());
    CK(cudaSetDevice(ctx->cfg.device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    mcrat_b200_comm *c = new mcrat_b200_comm();
    c->ctx = ctx;
    c->nranks = nranks;
    c->rank = rank;
    c->collectives = 0;
    c->words_dev = nullptr;
    c->words_host = nullptr;
    ncclResult_t r = a->CommInitRank(&c->nccl, nranks, u, rank);
    if (r != ncclSuccess) {
        ctx->err = std::string("ncclCommInitRank: ") + a->GetErrorString(r);
        delete c;
        return MCRAT_B200_ERR_CUDA;
    }
    cudaError_t e = cudaMalloc((void **)&c->words_dev, (size_t)(COMM_WORDS + nranks * 2) * sizeof(long long));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&c->words_host, (size_t)(COMM_WORDS + nranks * 2) * sizeof(long long));
    if (e != cudaSuccess) {
        ctx->err = std::string("comm_create: ") + cudaGetErrorString(e);
        if (c->words_dev) cudaFree(c->words_dev);
        a->CommDestroy(c->nccl);
        delete c;
        return MCRAT_B200_ERR_CUDA;
    }
    *out = c;
    return MCRAT_B200_OK;
}

API void mcrat_b200_comm_destroy(mcrat_b200_comm *c)
{
    if (!c) return;
    cudaSetDevice(c->ctx->cfg.device);
    cudaStreamSynchronize(c->ctx->stream);
    nccl_api()->CommDestroy(c->nccl);
    cudaFree(c->words_dev);
    cudaFreeHost(c->words_host);
    delete c;
}

API int mcrat_b200_comm_rank(const mcrat_b200_comm *c) { return c ? c->rank : -1; }
API int mcrat_b200_comm_size(const mcrat_b200_comm *c) { return c ? c->nranks : 0; }
API long long mcrat_b200_comm_collectives(const mcrat_b200_comm *c) { return c ? c->collectives : 0; }

// broadcastInterpolationData, Src/hot_x_section.c:709-823: the table of `root` replaces everybody's
API int mcrat_b200_comm_bcast_thermal_table(mcrat_b200_comm *c, int root)
{
    if (!c || root < 0 || root >= c->nranks) return c ? fail(c->ctx, MCRAT_B200_ERR_ARG, "comm_bcast_thermal_table: bad root") : MCRAT_B200_ERR_ARG;
    mcrat_b200_ctx *ctx = c->ctx;
    NcclApi *a = nccl_api();
    CK(cudaSetDevice(ctx->cfg.device));
    // the device copy is already in the interpolation layout (za), which is what every rank's lookups read
    NK(a->Broadcast(ctx->table_dev, ctx->table_dev, (size_t)(N_PH_E + 1) * (N_T + 1), ncclDouble, root, c->nccl, ctx->stream));
    c->collectives++;
    CK(cudaStreamSynchronize(ctx->stream)); // MPI_Barrier, Src/hot_x_section.c:825
    return MCRAT_B200_OK;
}

// createHotCrossSection (Src/hot_x_section.c:82-206) by all GPUs of the communicator: rank r integrates the points
// [r * chunk, (r + 1) * chunk) with K7, one all-gather assembles the table on every rank.  Same (seed, point) keys as
// mcrat_b200_build_thermal_table, hence the same table whatever the number of ranks.
API int mcrat_b200_comm_build_thermal_table(mcrat_b200_comm *c, long long calls, uint64_t seed, double *table_out,
                                            float *elapsed_ms)
{
    if (!c || calls < 1) return c ? fail(c->ctx, MCRAT_B200_ERR_ARG, "comm_build_thermal_table: calls >= 1") : MCRAT_B200_ERR_ARG;
    mcrat_b200_ctx *ctx = c->ctx;
    NcclApi *a = nccl_api();
    CK(cudaSetDevice(ctx->cfg.device));
    const int npts = (N_PH_E + 1) * (N_T + 1);
    const int chunk = (npts + c->nranks - 1) / c->nranks;
    const int first = c->rank * chunk;
    const int mine = std::max(0, std::min(chunk, npts - first));
    double *tab = nullptr;
    CK(cudaMalloc((void **)&tab, (size_t)chunk * (c->nranks + 1) * sizeof(double)));
    double *part = tab + (size_t)chunk * c->nranks;
    CK(cudaMemsetAsync(part, 0, (size_t)chunk * sizeof(double), ctx->stream));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    if (mine > 0) {
        hot_table_kernel<<<mine, 256, 0, ctx->stream>>>(part, calls, (uint32_t)seed ^ 0x4D435261u, (uint32_t)(seed >> 32), first);
        if (int rc = check_launch(ctx, "hot_table_kernel")) {
            cudaFree(tab);
            return rc;
        }
    }
    ncclResult_t r = a->AllGather(part, tab, (size_t)chunk, ncclDouble, c->nccl, ctx->stream);
    c->collectives++;
    if (r != ncclSuccess) {
        cudaFree(tab);
        ctx->err = std::string("ncclAllGather: ") + a->GetErrorString(r);
        return MCRAT_B200_ERR_CUDA;
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    std::vector<double> host(npts);
    CK(cudaMemcpyAsync(host.data(), tab, (size_t)npts * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(tab);
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (elapsed_ms) *elapsed_ms = ms;
    for (int k = 0; k < npts; ++k)
        if (host[k] != host[k]) return fail(ctx, MCRAT_B200_ERR_STATE, "NaN in the hot cross-section table (Src/hot_x_section.c:97-103)");
    if (table_out) memcpy(table_out, host.data(), (size_t)npts * sizeof(double));
    return mcrat_b200_set_thermal_table(ctx, host.data());
}

// Per-frame counters over all ranks: sums of the additive counters, maxima of iterations / clocks / flags.
API int mcrat_b200_comm_reduce_frame_stats(mcrat_b200_comm *c, const mcrat_b200_frame_stats *mine, mcrat_b200_frame_stats *total)
{
    if (!c || !mine || !total) return c ? fail(c->ctx, MCRAT_B200_ERR_ARG, "comm_reduce_frame_stats: null") : MCRAT_B200_ERR_ARG;
    mcrat_b200_ctx *ctx = c->ctx;
    NcclApi *a = nccl_api();
    CK(cudaSetDevice(ctx->cfg.device));
    long long *h = c->words_host;
    double *hd = (double *)(h + 16);
    // [0, 10): sums   [10, 16): integer maxima   [16, 20) as doubles: cs weight (sum), clocks (max)
    h[0] = mine->scatterings; h[1] = mine->relocations; h[2] = mine->photon_slots; h[3] = mine->cell_evals;
    h[4] = mine->box_evals; h[5] = mine->ref_equiv_evals; h[6] = mine->not_found; h[7] = mine->cs_emitted;
    h[8] = mine->scatt_cyclosynch_num_ph; h[9] = 0;
    h[10] = mine->iterations; h[11] = mine->cs_host_pending; h[12] = -(long long)mine->error; h[13] = h[14] = h[15] = 0;
    hd[0] = mine->cs_comptonized_weight; hd[1] = mine->time_now; hd[2] = mine->last_time_step; hd[3] = 0;
    long long *w = c->words_dev;
    CK(cudaMemcpyAsync(w, h, 20 * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    NK(a->GroupStart());
    NK(a->AllReduce(w, w, 10, ncclInt64, ncclSum, c->nccl, ctx->stream));
    NK(a->AllReduce(w + 10, w + 10, 6, ncclInt64, ncclMax, c->nccl, ctx->stream));
    NK(a->AllReduce(w + 16, w + 16, 1, ncclDouble, ncclSum, c->nccl, ctx->stream));
    NK(a->AllReduce(w + 17, w + 17, 3, ncclDouble, ncclMax, c->nccl, ctx->stream));
    NK(a->GroupEnd());
    c->collectives += 4;
    CK(cudaMemcpyAsync(h, w, 20 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *total = *mine;
    total->scatterings = h[0]; total->relocations = h[1]; total->photon_slots = h[2]; total->cell_evals = h[3];
    total->box_evals = h[4]; total->ref_equiv_evals = h[5]; total->not_found = (int)h[6]; total->cs_emitted = (int)h[7];
    total->scatt_cyclosynch_num_ph = (int)h[8];
    total->iterations = h[10]; total->cs_host_pending = (int)h[11]; total->error = -(int)h[12];
    total->cs_comptonized_weight = hd[0]; total->time_now = hd[1]; total->last_time_step = hd[2];
    return MCRAT_B200_OK;
}

static int comm_gather_counts(mcrat_b200_comm *c, long long *counts3 /* nranks x (capacity, output photons, null slots) */)
{
    mcrat_b200_ctx *ctx = c->ctx;
    NcclApi *a = nccl_api();
    long long *mine = c->words_dev + 32, *all = c->words_dev + COMM_WORDS; // all: nranks x 2 (output, nulls)
    CK(cudaMemsetAsync(mine, 0, 2 * sizeof(long long), ctx->stream));
    if (ctx->have_photons && ctx->d.cap > 0) {
        count_output_kernel<<<grid_for(ctx, ctx->d.cap, 256, 8), 256, 0, ctx->stream>>>(ctx->d, ctx->d.cap, mine);
        if (int rc = check_launch(ctx, "count_output_kernel")) return rc;
    }
    NK(a->AllGather(mine, all, 2, ncclInt64, c->nccl, ctx->stream));
    c->collectives++;
    long long cap = ctx->have_photons ? ctx->d.cap : 0;
    long long *capd = c->words_dev + 34, *caps = c->words_dev + 36; // nranks <= 28 capacities fit the fixed part ...
    std::vector<long long> caph(c->nranks);
    long long *caps_dev = nullptr;
    if (c->nranks > COMM_WORDS - 36) { // ... larger communicators get their own buffer
        CK(cudaMalloc((void **)&caps_dev, (size_t)c->nranks * sizeof(long long)));
        caps = caps_dev;
    }
    CK(cudaMemcpyAsync(capd, &cap, sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    ncclResult_t r = a->AllGather(capd, caps, 1, ncclInt64, c->nccl, ctx->stream);
    c->collectives++;
    if (r == ncclSuccess) {
        cudaMemcpyAsync(caph.data(), caps, (size_t)c->nranks * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(c->words_host, all, (size_t)c->nranks * 2 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream);
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (caps_dev) cudaFree(caps_dev);
    if (r != ncclSuccess) {
        ctx->err = std::string("ncclAllGather: ") + a->GetErrorString(r);
        return MCRAT_B200_ERR_CUDA;
    }
    CK(e);
    for (int k = 0; k < c->nranks; ++k) {
        counts3[3 * k] = caph[k];
        counts3[3 * k + 1] = c->words_host[2 * k];
        counts3[3 * k + 2] = c->words_host[2 * k + 1];
    }
    return MCRAT_B200_OK;
}

// Per-rank photon counts (load balance; what Src/merge.c:784-790 gathers before the merge): list capacity, photons with
// weight != 0 (the ones printPhotons writes), null slots.  Each output array holds comm_size entries; any may be NULL.
API int mcrat_b200_comm_photon_counts(mcrat_b200_comm *c, long long *list_capacity, long long *output_photons, long long *null_slots)
{
    if (!c) return MCRAT_B200_ERR_ARG;
    mcrat_b200_ctx *ctx = c->ctx;
    CK(cudaSetDevice(ctx->cfg.device));
    if (ctx->have_photons)
        if (int rc = flush_pushes(ctx)) return rc;
    std::vector<long long> c3((size_t)c->nranks * 3);
    if (int rc = comm_gather_counts(c, c3.data())) return rc;
    for (int k = 0; k < c->nranks; ++k) {
        if (list_capacity) list_capacity[k] = c3[3 * k];
        if (output_photons) output_photons[k] = c3[3 * k + 1];
        if (null_slots) null_slots[k] = c3[3 * k + 2];
    }
    return MCRAT_B200_OK;
}

// The photon lists of all ranks, concatenated in rank order (Src/merge.c:840-876 gathers the same data column by column
// with MPI_Allgatherv): records are packed on the device and travel GPU-to-GPU; `root` receives them in `photons`
// (host memory, `capacity` records), root = -1: every rank does.  counts[rank] = records of that rank (may be NULL);
// *total = their sum.  With too small a `capacity` nothing is transferred, *total tells how much is needed and the call
// returns MCRAT_B200_ERR_ARG on the receiving ranks (all ranks skip the transfer together).
API int mcrat_b200_comm_gather_photons(mcrat_b200_comm *c, int root, mcrat_photon *photons, long long capacity, long long *counts,
                                       long long *total)
{
    if (!c || root < -1 || root >= c->nranks) return c ? fail(c->ctx, MCRAT_B200_ERR_ARG, "comm_gather_photons: bad root") : MCRAT_B200_ERR_ARG;
    mcrat_b200_ctx *ctx = c->ctx;
    NcclApi *a = nccl_api();
    CK(cudaSetDevice(ctx->cfg.device));
    const bool recv = root == -1 || root == c->rank;
    if (ctx->have_photons)
        if (int rc = flush_pushes(ctx)) return rc;
    // every rank learns every list length and whether the receivers have room (one more word: min over ranks)
    std::vector<long long> c3((size_t)c->nranks * 3);
    if (int rc = comm_gather_counts(c, c3.data())) return rc;
    long long tot = 0;
    std::vector<long long> off(c->nranks + 1, 0);
    for (int k = 0; k < c->nranks; ++k) {
        off[k] = tot;
        tot += c3[3 * k];
        if (counts) counts[k] = c3[3 * k];
    }
    off[c->nranks] = tot;
    if (total) *total = tot;
    long long *flag = c->words_dev + 24;
    long long room = (!recv || (photons && capacity >= tot)) ? 1 : 0;
    CK(cudaMemcpyAsync(flag, &room, sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    NK(a->AllReduce(flag, flag, 1, ncclInt64, ncclMin, c->nccl, ctx->stream));
    c->collectives++;
    CK(cudaMemcpyAsync(&c->words_host[0], flag, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (c->words_host[0] == 0)
        return room ? fail(ctx, MCRAT_B200_ERR_ARG, "comm_gather_photons: a receiving rank's buffer is too small")
                    : fail(ctx, MCRAT_B200_ERR_ARG, "comm_gather_photons: capacity < total number of records");
    const long long mine = c3[3 * c->rank];
    if (mine > 0) {
        pack_kernel<<<grid_for(ctx, (int)mine, 256, 8), 256, 0, ctx->stream>>>(ctx->d, ctx->aos_dev, 0, (int)mine);
        if (int rc = check_launch(ctx, "pack_kernel")) return rc;
    }
    mcrat_photon *all_dev = nullptr;
    if (recv && tot > 0) CK(cudaMalloc((void **)&all_dev, (size_t)tot * sizeof(mcrat_photon)));
    ncclResult_t r = a->GroupStart();
    for (int k = 0; k < c->nranks && r == ncclSuccess; ++k) {
        const size_t bytes = (size_t)c3[3 * k] * sizeof(mcrat_photon);
        if (!bytes) continue;
        if (root == -1) {
            r = a->Broadcast(k == c->rank ? (const void *)ctx->aos_dev : (const void *)(all_dev + off[k]), all_dev + off[k], bytes,
                             ncclUint8, k, c->nccl, ctx->stream);
            c->collectives++;
        } else {
            if (k == c->rank && c->rank != root) {
                r = a->Send(ctx->aos_dev, bytes, ncclUint8, root, c->nccl, ctx->stream);
                c->collectives++;
            }
            if (c->rank == root && k != root) {
                r = a->Recv(all_dev + off[k], bytes, ncclUint8, k, c->nccl, ctx->stream);
                c->collectives++;
            }
        }
    }
    ncclResult_t r2 = a->GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) {
        if (all_dev) cudaFree(all_dev);
        ctx->err = std::string("comm_gather_photons: ") + a->GetErrorString(r);
        return MCRAT_B200_ERR_CUDA;
    }
    cudaError_t e = cudaSuccess;
    if (recv && tot > 0) {
        if (root != -1 && mine > 0) // the root's own records
            e = cudaMemcpyAsync(all_dev + off[c->rank], ctx->aos_dev, (size_t)mine * sizeof(mcrat_photon), cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(photons, all_dev, (size_t)tot * sizeof(mcrat_photon), cudaMemcpyDeviceToHost, ctx->stream);
    }
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    if (all_dev) cudaFree(all_dev);
    CK(e);
    CK(e2);
    return MCRAT_B200_OK;
}
