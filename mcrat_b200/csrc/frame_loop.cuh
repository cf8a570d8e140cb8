// frame_loop.cuh -- persistent frame loop (team and solo kernels).
// Part of the single translation unit mcrat_b200.cu (included there, in this order); not a stand-alone header.
#pragma once

// ------------------------------------------------------------------------------------------
// Persistent frame loop: the whole while-loop of Src/mcrat.c:761-851 in ONE launch.
//
// The streamed loop above costs four dependent kernel launches per scattering; for lists that fit
// in L2 the launch boundaries, not the work, set the iteration time.  Here every sub-shard (one
// reference "rank") is owned by `bps` resident blocks that iterate on their own:
//   pass (push + re-check + free-path draw + block arg-min)  ->  arrive
//   last arriver: re-locate the few photons that left their cell (one warp per photon through the
//                 bounding-box index), shard arg-min, scattering event, publish the new clock /
//                 push list                                  ->  release
//   the other blocks spin on the shard's generation word (ld.acquire.gpu) and start the next pass.
// Shards never wait for each other (exactly like MPI ranks), so a rank that finishes its frame or
// rejects a Klein-Nishina candidate does not hold the others up.  All blocks are co-resident
// (cooperative launch); with one block per shard a block walks through its shards one after the other.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

constexpr int RELOC_HEAVY = 64; // re-locations per shard and iteration above which the grid-wide K1b / K1c serve better

// relocated photons of this iteration: new cell, comoving momentum, tau', free-path draw
template <int THREADS>
__device__ __forceinline__ void finish_reloc(DevCtx &d, ShardState &st, const int s, const int R)
{
    int missing = 0;
    for (int j = threadIdx.x; j < R; j += THREADS)
        if (finish_one<true>(d, st, s, d.reloc_slot[st.first + j], d.reloc_best[st.first + j], 0)) missing++;
    if (missing) atomicAdd(&d.gs->not_found, missing);
    __syncthreads();
    if (threadIdx.x == 0) {
        d.sh[s].reloc_n = 0;
        if (R > RELOC_HEAVY) st.reloc_heavy = 1;
    }
}

// the shard's relocation entries [first, first+R): one warp per photon through the bounding-box index
template <int THREADS>
__device__ __forceinline__ void relocate_shard(DevCtx &d, ShardState &st, const int s, const int R)
{
    long long cells_tested = 0, boxes_tested = 0;
    for (int j = threadIdx.x >> 5; j < R; j += THREADS / 32) {
        const int q = st.first + j;
        const int best = warp_locate_indexed(d, d.reloc_h0[q], d.reloc_h1[q], d.reloc_h2[q], cells_tested, boxes_tested);
        if ((threadIdx.x & 31) == 0) d.reloc_best[q] = best;
    }
    if ((threadIdx.x & 31) == 0 && (cells_tested | boxes_tested)) {
        atomicAdd((unsigned long long *)&d.gs->cell_evals, (unsigned long long)cells_tested);
        atomicAdd((unsigned long long *)&d.gs->box_evals, (unsigned long long)boxes_tested);
    }
    __syncthreads();
    finish_reloc<THREADS>(d, st, s, R);
    __syncthreads();
}

// A sub-shard is run by a team of `bps` pass blocks and one event block, all resident:
//   pass block b:  pass over its slice -> ticket on gst.arrive -> spin on gst.gen -> pull the state -> next pass
//   event block:   spin until all bps tickets of the iteration are drawn -> re-locate -> shard arg-min -> event;
//                  the state is published (gst.gen released) the moment a candidate is accepted, so the pass blocks
//                  run the next pass while the event block is still busy with the second half of the scatter; the
//                  scattered photon's own next pass is done by the event block (mini-pass, slot `bps` of the minima).
template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 2 : 4) frame_loop_kernel(DevCtx d, const int bps)
{
    __shared__ ShardState st; // this block's copy of the shard's state
    __shared__ int sh_flag;
    GlobalState &gs = *d.gs;
    const int team = bps + 1;
    const int groups = gridDim.x / team;
    const int g = blockIdx.x / team, role = blockIdx.x - g * team;
    if (g >= groups) return;
    const bool is_event_block = (role == bps);

    for (int s = g; s < d.nshards; s += groups) {
        ShardState &gst = d.sh[s];
        // global -> shared (the copied part only; the protocol words live in global memory)
        auto pull = [&]() {
            if (threadIdx.x < SHARD_STATE_WORDS)
                reinterpret_cast<unsigned long long *>(&st)[threadIdx.x] = reinterpret_cast<const unsigned long long *>(&gst)[threadIdx.x];
            __syncthreads();
        };
        // thread 0 spins until *word >= target (relaxed polls, one acquire at the end); false after ~1 s
        auto spin_until = [&](const unsigned *word, unsigned target) -> bool {
            if (threadIdx.x == 0) {
                unsigned spins = 0;
                int ok = 1;
                if (ld_acquire_u32(word) < target) { // fast path: already there, one round trip
                    while (ld_relaxed_u32(word) < target) {
                        __nanosleep(40);
                        // never in a healthy run; refuse to hang the GPU.  A hot cross-section lookup outside the table
                        // integrates for ~0.4 s inside a pass (total_thermal_cross_section_mc): allow for a few of them
                        if (++spins > (d.tau_calc == TAU_TABLE ? (1u << 28) : (1u << 24))) {
                            raise_error(&gs, MCRAT_B200_ERR_STATE, -1, ERR_SITE_LOOP_SPIN);
                            ok = 0;
                            break;
                        }
                    }
                    (void)ld_acquire_u32(word);
                }
                sh_flag = ok;
            }
            __syncthreads();
            return sh_flag != 0;
        };
        __syncthreads();
        pull();
        // stop test at entry: shard state only, so that all blocks of the team decide alike
        bool halt = st.done | st.pause_cs | (gs.max_iters >= 0 && st.iters_done >= gs.max_iters);
        unsigned k = 0; // iterations of this launch; gst.arrive / gst.gen were zeroed before it

        if (!is_event_block) {
            // ---------------- pass block ----------------
            const int b = role;
            while (!halt) {
                double best_t = DBL_MAX;
                int best_i = INT_MAX;
                pass_body<true, true, THREADS>(d, st, s, b, bps, 0, 0, best_t, best_i);
                block_argmin<THREADS>(best_t, best_i);
                __syncthreads();
                if (threadIdx.x == 0) {
                    int bidx = -1;
                    double btemp = 0;
                    if (best_i != INT_MAX) {
                        bidx = d.ph.idx[best_i];
                        if (bidx >= 0) btemp = d.cells.temp[bidx];
                    }
                    d.bm_t[s * team + b] = best_t;
                    d.bm_i[s * team + b] = best_i;
                    d.bm_idx[s * team + b] = bidx;
                    d.bm_temp[s * team + b] = btemp;
                    __threadfence();
                    atomicAdd(&gst.arrive, 1u);
                }
                ++k;
                if (!spin_until(&gst.gen, k)) break;
                pull();
                halt = st.halt != 0;
                __syncthreads();
            }
        } else {
            // ---------------- event block ----------------
            if (threadIdx.x == 0) {
                d.bm_t[s * team + bps] = DBL_MAX;
                d.bm_i[s * team + bps] = INT_MAX;
                st.mini_slot = -1;
            }
            __syncthreads();
            while (!halt) {
                if (!spin_until(&gst.arrive, (k + 1) * (unsigned)bps)) break;
                ++k;
                // this thread's entry of the block minima and the relocation count: one round trip
                double pre_t = DBL_MAX;
                int pre_i = INT_MAX;
                const bool have_pre = (team <= THREADS);
                int pre_idx = -2;
                double pre_temp = 0;
                if (have_pre && (int)threadIdx.x < team) {
                    pre_t = *(volatile double *)&d.bm_t[s * team + threadIdx.x];
                    pre_i = *(volatile int *)&d.bm_i[s * team + threadIdx.x];
                    pre_idx = *(volatile int *)&d.bm_idx[s * team + threadIdx.x];
                    pre_temp = *(volatile double *)&d.bm_temp[s * team + threadIdx.x];
                }
                const int R = *(volatile int *)&gst.reloc_n;
                if (R > 0) relocate_shard<THREADS>(d, st, s, R);
                const bool released = event_body<THREADS>(d, s, st.first, R, team, 0, 0.0, st, &gst, k, s * team + bps, have_pre,
                                                          pre_t, pre_i, pre_idx, pre_temp);
                if (!released) {
                    // frame end, Klein-Nishina walk exhausted, cyclo-synchrotron run: publish now
                    if (threadIdx.x == 0) {
                        st.halt = (loop_stopped(gs, st) || st.reloc_heavy) ? 1 : 0;
                        st.mini_slot = -1;
                        d.bm_t[s * team + bps] = DBL_MAX;
                        d.bm_i[s * team + bps] = INT_MAX;
                    }
                    __syncthreads();
                    if (threadIdx.x < SHARD_STATE_WORDS)
                        reinterpret_cast<unsigned long long *>(&gst)[threadIdx.x] = reinterpret_cast<const unsigned long long *>(&st)[threadIdx.x];
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        __threadfence();
                        st_release_u32(&gst.gen, k);
                    }
                } else if (threadIdx.x == 0 && st.mini_slot < 0) {
                    d.bm_t[s * team + bps] = DBL_MAX; // released with a halt: no mini-pass ran
                    d.bm_i[s * team + bps] = INT_MAX;
                }
                __syncthreads();
                halt = st.halt != 0;
            }
            // the state proper is current in global memory (published with every release)
        }
    }
}

// ------------------------------------------------------------------------------------------
// The same team as ONE THREAD-BLOCK CLUSTER (round 2): bps pass blocks (cluster ranks 0 .. bps-1) + the event block (rank
// bps), at most 16 blocks.  The protocol is frame_loop_kernel's, but its three hand-overs no longer go through L2:
//   pass block -> event block   the block's minimum (time, slot, cell, temperature, "some photon left its cell") is stored
//                               into the event block's shared memory (st.shared::cluster) and the block arrives on the event
//                               block's mbarrier (mbarrier.arrive.release.cluster on the remote address) -- instead of four
//                               global stores, a fence and an atomicAdd that the event block polls through L2;
//   event block -> pass blocks  on the early release the helper warp pushes the new shard state into every pass block's
//                               shared memory and arrives on each one's generation mbarrier -- instead of a global copy the
//                               pass blocks poll for and then read back;
//   waiting                     mbarrier.try_wait on the block's OWN shared memory (acquire at cluster scope).
// Photon data still travels through global memory; __threadfence() on the producing side before the arrival and the acquire
// on the waiting side order it, as before.  A cluster needs no cooperative launch (its blocks are co-scheduled by
// definition), clusters never talk to each other, and a team that cannot get a cluster (bps + 1 > 16 is clamped; a device
// without free cluster slots) runs frame_loop_kernel instead.
// MEASURED (tools/gpu_o.sh, gpurun_out/ab_o.log -> profiles/ncu_r02_summary.md): photons bit-identical to the cooperative team
// kernel's on every workload, iteration time NOT better -- C2 16 ranks 21.5 against 20.8 us, C1 20.5 / 19.3, C5 24.4 / 23.8,
// C3 33.8 / 33.9, C2 64 ranks 23.1 / 24.2, C5 10^6 photons 47.0 / 40.0 (15 instead of 17 pass blocks per rank).  The six to
// seven microseconds between two scatterings of a rank are not signalling latency: they are the pass itself (one dependent
// L2 round trip for the photon columns, the draw, the stores) and the event's first loads.  The kernel therefore runs only
// on request (MCRAT_B200_CLUSTER_TEAM=1); the cooperative team kernel stays the default.
// ------------------------------------------------------------------------------------------
constexpr int CLUSTER_TEAM_MAX = 16;

struct TeamMail { // one per pass block, in the event block's shared memory; written remotely
    double t, temp;
    int i, idx;
    int reloc, pad;
};

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// thread 0 of a block: wait for phase `parity` of its own barrier; false after ~1 s (never in a healthy run)
__device__ __forceinline__ bool mbar_wait_cluster(GlobalState &gs, uint64_t *bar, uint32_t parity, unsigned limit)
{
    unsigned spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (++spins > limit) {
            raise_error(&gs, MCRAT_B200_ERR_STATE, -1, ERR_SITE_LOOP_SPIN);
            return false;
        }
        if ((spins & 4095u) == 0 && *(volatile int *)&gs.error != 0) return false;
    }
    return true;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2) frame_loop_cluster_kernel(DevCtx d, const int bps)
{
    __shared__ ShardState st;  // pass blocks: written remotely by the event block at every release
    __shared__ int sh_flag, sh_reloc;
    __shared__ __align__(8) uint64_t bar_gen; // pass blocks: one arrival (the event block's) per iteration
    __shared__ __align__(8) uint64_t bar_arr; // event block: bps arrivals per iteration
    __shared__ TeamMail mail[CLUSTER_TEAM_MAX];
    GlobalState &gs = *d.gs;
    const int team = bps + 1;
    const int groups = gridDim.x / team;
    const int g = blockIdx.x / team;
    const int role = (int)cluster_ctarank();
    const bool is_event_block = (role == bps);
    const unsigned limit = d.tau_calc == TAU_TABLE ? (1u << 30) : (1u << 26); // try_wait suspends for a while by itself
    uint32_t ph_gen = 0, ph_arr = 0; // phases completed of this block's own barriers
    if (threadIdx.x == 0) {
        mbar_init(&bar_gen, 1);
        mbar_init(&bar_arr, bps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();
    const uint32_t st_addr = smem_u32(&st), gen_addr = smem_u32(&bar_gen);
    const uint32_t arr_remote = cluster_map(smem_u32(&bar_arr), (uint32_t)bps);

    for (int s = g; s < d.nshards; s += groups) {
        ShardState &gst = d.sh[s];
        __syncthreads();
        if (threadIdx.x < SHARD_STATE_WORDS)
            reinterpret_cast<unsigned long long *>(&st)[threadIdx.x] = reinterpret_cast<const unsigned long long *>(&gst)[threadIdx.x];
        __syncthreads();
        // stop test at entry: shard state only, so that all blocks of the team decide alike
        bool halt = st.done | st.pause_cs | (gs.max_iters >= 0 && st.iters_done >= gs.max_iters);
        if (!is_event_block) {
            // ---------------- pass block ----------------
            const int b = role;
            const uint32_t mail_remote = cluster_map(smem_u32(&mail[b]), (uint32_t)bps);
            while (!halt) {
                if (threadIdx.x == 0) sh_reloc = 0;
                __syncthreads();
                double best_t = DBL_MAX;
                int best_i = INT_MAX;
                pass_body<true, true, THREADS>(d, st, s, b, bps, 0, 0, best_t, best_i, &sh_reloc);
                block_argmin<THREADS>(best_t, best_i);
                __syncthreads();
                if (threadIdx.x == 0) {
                    int bidx = -1;
                    double btemp = 0;
                    if (best_i != INT_MAX) {
                        bidx = d.ph.idx[best_i];
                        if (bidx >= 0) btemp = d.cells.temp[bidx];
                    }
                    __threadfence(); // this block's photon columns and relocation entries before the arrival
                    cluster_st_u64(mail_remote, (unsigned long long)__double_as_longlong(best_t));
                    cluster_st_u64(mail_remote + 8, (unsigned long long)__double_as_longlong(btemp));
                    cluster_st_u64(mail_remote + 16, ((unsigned long long)(unsigned)bidx << 32) | (unsigned long long)(unsigned)best_i);
                    cluster_st_u32(mail_remote + 24, (unsigned)sh_reloc);
                    cluster_mbar_arrive(arr_remote);
                    sh_flag = mbar_wait_cluster(gs, &bar_gen, ph_gen & 1u, limit) ? 1 : 0; // the new state is in `st`
                }
                ++ph_gen;
                __syncthreads();
                if (!sh_flag) break;
                halt = st.halt != 0;
            }
        } else {
            // ---------------- event block ----------------
            if (threadIdx.x == 0) {
                d.bm_t[s * team + bps] = DBL_MAX; // the mini-pass slot stays in global memory (written and read by this block)
                d.bm_i[s * team + bps] = INT_MAX;
                st.mini_slot = -1;
                sh_reloc = 0;
            }
            __syncthreads();
            unsigned k = 0;
            while (!halt) {
                // this block's own mini-pass result of the last event, requested before the wait
                double pre_t = DBL_MAX, pre_temp = 0;
                int pre_i = INT_MAX, pre_idx = -2;
                if ((int)threadIdx.x == bps) {
                    pre_t = *(volatile double *)&d.bm_t[s * team + bps];
                    pre_i = *(volatile int *)&d.bm_i[s * team + bps];
                    pre_idx = *(volatile int *)&d.bm_idx[s * team + bps];
                    pre_temp = *(volatile double *)&d.bm_temp[s * team + bps];
                }
                if (threadIdx.x == 0) sh_flag = mbar_wait_cluster(gs, &bar_arr, ph_arr & 1u, limit) ? 1 : 0;
                ++ph_arr;
                __syncthreads();
                if (!sh_flag) break;
                ++k;
                int any_reloc = sh_reloc; // the last mini-pass
                if ((int)threadIdx.x < bps) {
                    const volatile TeamMail &m = mail[threadIdx.x];
                    pre_t = m.t;
                    pre_i = m.i;
                    pre_idx = m.idx;
                    pre_temp = m.temp;
                    any_reloc |= m.reloc;
                }
                any_reloc = __syncthreads_or(any_reloc);
                int R = 0;
                if (any_reloc) {
                    R = *(volatile int *)&gst.reloc_n;
                    if (R > 0) relocate_shard<THREADS>(d, st, s, R);
                }
                const bool released = event_body<THREADS>(d, s, st.first, R, team, 0, 0.0, st, &gst, k, s * team + bps, true, pre_t,
                                                          pre_i, pre_idx, pre_temp, bps, st_addr, gen_addr, &sh_reloc);
                if (!released) {
                    // frame end, Klein-Nishina walk exhausted, cyclo-synchrotron run: publish now
                    if (threadIdx.x == 0) {
                        st.halt = (loop_stopped(gs, st) || st.reloc_heavy) ? 1 : 0;
                        st.mini_slot = -1;
                        d.bm_t[s * team + bps] = DBL_MAX;
                        d.bm_i[s * team + bps] = INT_MAX;
                    }
                    __syncthreads();
                    if (threadIdx.x < SHARD_STATE_WORDS)
                        reinterpret_cast<unsigned long long *>(&gst)[threadIdx.x] = reinterpret_cast<const unsigned long long *>(&st)[threadIdx.x];
                    __syncthreads();
                    if (threadIdx.x < 32) cluster_publish_warp(st, bps, st_addr, gen_addr, (int)threadIdx.x);
                } else if (threadIdx.x == 0 && st.mini_slot < 0) {
                    d.bm_t[s * team + bps] = DBL_MAX; // released with a halt: no mini-pass ran
                    d.bm_i[s * team + bps] = INT_MAX;
                }
                __syncthreads();
                halt = st.halt != 0;
            }
            // the state proper is current in global memory (written with every release)
        }
    }
    cluster_sync_all(); // nobody leaves while its shared memory may still be written remotely
}

// ------------------------------------------------------------------------------------------
// Persistent loop for lists LARGER than L2 (PERSISTENT_STREAM): the same team protocol -- tickets on gst.arrive, early
// release on gst.gen, mini-pass by the event block -- but the pass blocks are not tied to a shard.  Two kernels run
// side by side (the pass needs 48 registers and 40 warps per SM to keep HBM busy, the event 128 registers and three
// warps; one kernel could not have both):
//   frame_stream_event_kernel  resident blocks, each serving a few sub-shards in turn; per shard exactly the event block of
//                              frame_loop_kernel
//   frame_stream_pass_kernel   resident blocks pull pass items (iteration k, shard s, block b) from a counter, in that
//                              order: the item waits until shard s has released generation k (its event of iteration
//                              k - 1 has accepted a candidate), runs pass_body over its slice and draws a ticket.
// With S shards the stream reaches shard s again one whole list later, so an event (tens of microseconds) has the time
// the other S - 1 shards' photons take to stream past (hundreds): the HBM pipe never waits for a scattering and never
// meets a launch boundary.  An item depends only on items before it in the order, and every block that holds an item
// is running (items are pulled, not assigned), so the protocol cannot deadlock as long as the event blocks are resident:
// they are launched first and counted (stream_evt_ready) before the first item is pulled; a block that waits longer
// than ~1 s raises MCRAT_B200_ERR_STATE instead of hanging the device and the host falls back to the streamed loop.
// A halted shard (frame end, max_iters, pause, many re-locations) releases generation UINT_MAX: its items are skipped.
// ------------------------------------------------------------------------------------------
constexpr unsigned STREAM_SPIN_LIMIT = 1u << 24;
constexpr unsigned STREAM_START_SPINS = 1u << 19; // x >= 40 ns: the other grid has ~30 ms to show up before the launch is called off
enum { STREAM_UNDECIDED = 0, STREAM_GO = 1, STREAM_ABORT = 2 };

// Start-up handshake of the two grids (thread 0 of every block).  Pass blocks: wait until all event blocks are resident,
// then vote GO; event blocks: wait for the vote.  Whoever waits too long votes ABORT; the first vote stands (atomicCAS)
// and is taken before any photon is touched, so an aborted launch leaves the lists exactly as they were.
__device__ __forceinline__ int stream_handshake(GlobalState &gs, const bool is_pass, const int evt_blocks)
{
    unsigned spins = 0;
    for (;;) {
        const int state = *(volatile int *)&gs.stream_state;
        if (state != STREAM_UNDECIDED) return state;
        if (is_pass && *(volatile int *)&gs.stream_evt_ready >= evt_blocks) {
            atomicCAS(&gs.stream_state, STREAM_UNDECIDED, STREAM_GO);
            continue;
        }
        __nanosleep(40);
        if (++spins > STREAM_START_SPINS) atomicCAS(&gs.stream_state, STREAM_UNDECIDED, STREAM_ABORT);
    }
}
constexpr int STREAM_PASS_CTAS_PER_SM = 4; // pass blocks on an SM that also holds an event block (48 + 16 K registers); 5 elsewhere
// Blocks of the two grids share SMs, and an SM runs with ONE split of its 256 KB between L1 and shared memory: a block of a
// kernel the driver has configured for another split waits until the SM has drained.  Both kernels therefore (a) ask for the
// same carve-out (cudaFuncAttributePreferredSharedMemoryCarveout, 32 KB) and (b) actually need it at full occupancy -- the
// pass blocks carry 3 KB of unused dynamic shared memory for that -- or the driver overrides the preference.  Measured:
// without (a) or without (b) the pass grid starts only when the event grid has left (the start-up handshake calls the launch
// off); with the maximum carve-out for both they meet, but the pass loses its spills' L1 (281 against 207 us per iteration).
constexpr int STREAM_PASS_SMEM_PAD = 3072;

__device__ __forceinline__ bool stream_wait_ge(GlobalState &gs, const unsigned *word, unsigned target, unsigned limit)
{
    if (ld_acquire_u32(word) >= target) return true;
    unsigned spins = 0;
    while (ld_relaxed_u32(word) < target) {
        __nanosleep(40);
        if (++spins > limit) {
            raise_error(&gs, MCRAT_B200_ERR_STATE, -1, ERR_SITE_LOOP_SPIN);
            return false;
        }
        if ((spins & 1023u) == 0 && *(volatile int *)&gs.error != 0) return false; // somebody else gave up
    }
    (void)ld_acquire_u32(word);
    return true;
}

__global__ void __launch_bounds__(PASS_THREADS, MCRAT_PASS_MINB) frame_stream_pass_kernel(DevCtx d, const int bps, const int evt_blocks)
{
    __shared__ ShardState st;
    __shared__ unsigned long long sh_item;
    __shared__ int sh_flag;
    GlobalState &gs = *d.gs;
    const int S = d.nshards, team = bps + 1;
    const unsigned long long per_iter = (unsigned long long)S * (unsigned long long)bps;
    const unsigned limit = d.tau_calc == TAU_TABLE ? (1u << 28) : STREAM_SPIN_LIMIT;
    if (threadIdx.x == 0) {
        sh_flag = (stream_handshake(gs, true, evt_blocks) == STREAM_GO) ? 1 : 0;
        if (sh_flag) sh_item = atomicAdd(&gs.stream_work, 1ull);
    }
    __syncthreads();
    if (!sh_flag) return;
    for (;;) {
        const unsigned long long item = sh_item;
        const unsigned long long k64 = item / per_iter;
        const unsigned rem = (unsigned)(item - k64 * per_iter);
        const int s = (int)(rem / (unsigned)bps), b = (int)(rem - (unsigned)s * (unsigned)bps);
        const unsigned k = (unsigned)k64;
        ShardState &gst = d.sh[s];
        // wait until shard s has released generation k (halted shards release UINT_MAX); stop when every shard has halted
        if (threadIdx.x == 0) {
            int ok = (*(volatile int *)&gs.stream_halted >= S || *(volatile int *)&gs.error != 0) ? 0 : 1;
            if (ok && k > 0) ok = stream_wait_ge(gs, &gst.gen, k, limit) ? 1 : 0;
            sh_flag = ok;
        }
        __syncthreads();
        if (!sh_flag) return;
        if (threadIdx.x < SHARD_STATE_WORDS)
            reinterpret_cast<unsigned long long *>(&st)[threadIdx.x] = reinterpret_cast<const unsigned long long *>(&gst)[threadIdx.x];
        __syncthreads();
        // the stop test at entry, shard state only (the event block decides alike), later the published halt flag
        // (a shard that was stopped at entry never publishes: its state says so by itself)
        const bool halt = (st.halt != 0) | st.done | st.pause_cs | (gs.max_iters >= 0 && st.iters_done >= gs.max_iters);
        double best_t = DBL_MAX;
        int best_i = INT_MAX;
        if (!halt) {
            pass_body<true, true, PASS_THREADS>(d, st, s, b, bps, 0, 0, best_t, best_i);
            block_argmin<PASS_THREADS>(best_t, best_i);
        }
        __syncthreads(); // everybody is done with sh_item and st
        if (threadIdx.x == 0) {
            // the next item is requested before this one's result is published: one round trip instead of two
            const unsigned long long next = atomicAdd(&gs.stream_work, 1ull);
            if (!halt) {
                int bidx = -1;
                double btemp = 0;
                if (best_i != INT_MAX) {
                    bidx = d.ph.idx[best_i];
                    if (bidx >= 0) btemp = d.cells.temp[bidx];
                }
                d.bm_t[s * team + b] = best_t;
                d.bm_i[s * team + b] = best_i;
                d.bm_idx[s * team + b] = bidx;
                d.bm_temp[s * team + b] = btemp;
                __threadfence();
                atomicAdd(&gst.arrive, 1u);
            }
            sh_item = next;
        }
        __syncthreads();
    }
}

// One event block serves `stride`-spaced sub-shards s = blockIdx.x, blockIdx.x + gridDim.x, ... (at most STREAM_EVT_SHARDS of
// them), one after the other within an iteration -- the order the pass stream reaches them in, gridDim.x shards (tens
// of microseconds of streaming) apart, so a block is rarely asked for two events at once.  Fewer event blocks leave
// more SMs with room for a fifth pass block.  Every shard keeps its own state in shared memory between iterations,
// exactly as the event block of frame_loop_kernel does for its one shard.
constexpr int STREAM_EVT_SHARDS = 8;

template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 2 : 4) frame_stream_event_kernel(DevCtx d, const int bps)
{
    __shared__ ShardState sts[STREAM_EVT_SHARDS];
    __shared__ int sh_flag;
    __shared__ int sh_halted[STREAM_EVT_SHARDS];
    __shared__ unsigned sh_k[STREAM_EVT_SHARDS];
    GlobalState &gs = *d.gs;
    const int team = bps + 1;
    const int S = d.nshards, E = gridDim.x;
    const int mine = (S - (int)blockIdx.x + E - 1) / E; // shards of this block
    const unsigned limit = d.tau_calc == TAU_TABLE ? (1u << 28) : STREAM_SPIN_LIMIT;
    if (threadIdx.x == 0) {
        atomicAdd(&gs.stream_evt_ready, 1); // resident
        sh_flag = (stream_handshake(gs, false, 0) == STREAM_GO) ? 1 : 0;
    }
    __syncthreads();
    if (!sh_flag) return; // called off: nothing touched
    int running = 0;
    for (int j = 0; j < mine; ++j) {
        const int s = blockIdx.x + j * E;
        ShardState &gst = d.sh[s];
        ShardState &st = sts[j];
        if (threadIdx.x < SHARD_STATE_WORDS)
            reinterpret_cast<unsigned long long *>(&st)[threadIdx.x] = reinterpret_cast<const unsigned long long *>(&gst)[threadIdx.x];
        __syncthreads();
        const bool halt = st.done | st.pause_cs | (gs.max_iters >= 0 && st.iters_done >= gs.max_iters);
        if (threadIdx.x == 0) {
            d.bm_t[s * team + bps] = DBL_MAX;
            d.bm_i[s * team + bps] = INT_MAX;
            st.mini_slot = -1;
            sh_halted[j] = halt ? 1 : 0;
            sh_k[j] = 0;
            if (halt) { // stopped at entry: every item of this shard is skipped
                __threadfence();
                st_release_u32(&gst.gen, 0xFFFFFFFFu);
                atomicAdd(&gs.stream_halted, 1);
            }
        }
        if (!halt) running++;
        __syncthreads();
    }
    bool gave_up = false;
    while (running > 0 && !gave_up) {
        for (int j = 0; j < mine; ++j) {
            if (sh_halted[j]) continue;
            const int s = blockIdx.x + j * E;
            ShardState &gst = d.sh[s];
            ShardState &st = sts[j];
            unsigned k = sh_k[j];
            if (threadIdx.x == 0) sh_flag = stream_wait_ge(gs, &gst.arrive, (k + 1) * (unsigned)bps, limit) ? 1 : 0;
            __syncthreads();
            if (!sh_flag) {
                gave_up = true;
                break;
            }
            ++k;
            double pre_t = DBL_MAX;
            int pre_i = INT_MAX;
            const bool have_pre = (team <= THREADS);
            int pre_idx = -2;
            double pre_temp = 0;
            if (have_pre && (int)threadIdx.x < team) {
                pre_t = *(volatile double *)&d.bm_t[s * team + threadIdx.x];
                pre_i = *(volatile int *)&d.bm_i[s * team + threadIdx.x];
                pre_idx = *(volatile int *)&d.bm_idx[s * team + threadIdx.x];
                pre_temp = *(volatile double *)&d.bm_temp[s * team + threadIdx.x];
            }
            const int R = *(volatile int *)&gst.reloc_n;
            if (R > 0) relocate_shard<THREADS>(d, st, s, R);
            const bool released = event_body<THREADS>(d, s, st.first, R, team, 0, 0.0, st, &gst, k, s * team + bps, have_pre,
                                                      pre_t, pre_i, pre_idx, pre_temp);
            if (!released) {
                // frame end, Klein-Nishina walk exhausted: publish now
                if (threadIdx.x == 0) {
                    st.halt = (loop_stopped(gs, st) || st.reloc_heavy) ? 1 : 0;
                    st.mini_slot = -1;
                    d.bm_t[s * team + bps] = DBL_MAX;
                    d.bm_i[s * team + bps] = INT_MAX;
                }
                __syncthreads();
                if (threadIdx.x < SHARD_STATE_WORDS)
                    reinterpret_cast<unsigned long long *>(&gst)[threadIdx.x] = reinterpret_cast<const unsigned long long *>(&st)[threadIdx.x];
                __syncthreads();
                if (threadIdx.x == 0) {
                    __threadfence();
                    st_release_u32(&gst.gen, k);
                }
            } else if (threadIdx.x == 0 && st.mini_slot < 0) {
                d.bm_t[s * team + bps] = DBL_MAX; // released with a halt: no mini-pass ran
                d.bm_i[s * team + bps] = INT_MAX;
            }
            __syncthreads();
            if (threadIdx.x == 0) sh_k[j] = k;
            if (st.halt != 0) { // halted: every later item of this shard is skipped
                running--;
                if (threadIdx.x == 0) {
                    sh_halted[j] = 1;
                    __threadfence();
                    st_release_u32(&gst.gen, 0xFFFFFFFFu);
                    atomicAdd(&gs.stream_halted, 1);
                }
            }
            __syncthreads();
        }
    }
    if (gave_up && threadIdx.x == 0) { // an error has been raised; let the pass blocks go
        for (int j = 0; j < mine; ++j)
            if (!sh_halted[j]) {
                st_release_u32(&d.sh[blockIdx.x + j * E].gen, 0xFFFFFFFFu);
                atomicAdd(&gs.stream_halted, 1);
            }
    }
}

// With more sub-shards than resident teams the GPU is busy anyway (many events in flight per SM) and throughput,
// not the latency of one shard, is what counts: one block per sub-shard does pass and event in turn, and walks through
// its shards one after the other if there are more shards than resident blocks.  No inter-block protocol at all.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 2 : 4) frame_loop_solo_kernel(DevCtx d)
{
    __shared__ ShardState st;
    GlobalState &gs = *d.gs;
    for (int s = blockIdx.x; s < d.nshards; s += gridDim.x) {
        ShardState &gst = d.sh[s];
        __syncthreads();
        if (threadIdx.x < SHARD_STATE_WORDS)
            reinterpret_cast<unsigned long long *>(&st)[threadIdx.x] = reinterpret_cast<const unsigned long long *>(&gst)[threadIdx.x];
        __syncthreads();
        bool halt = st.done | st.pause_cs | (gs.max_iters >= 0 && st.iters_done >= gs.max_iters);
        while (!halt) {
            double best_t = DBL_MAX;
            int best_i = INT_MAX;
            pass_body<true, true, THREADS>(d, st, s, 0, 1, 0, 0, best_t, best_i);
            block_argmin<THREADS>(best_t, best_i);
            if (threadIdx.x == 0) {
                d.bm_t[s] = best_t;
                d.bm_i[s] = best_i;
            }
            __syncthreads();
            const int R = *(volatile int *)&gst.reloc_n;
            if (R > 0) relocate_shard<THREADS>(d, st, s, R);
            event_body<THREADS>(d, s, st.first, R, 1, 0, 0.0, st);
            if (threadIdx.x == 0) st.halt = (loop_stopped(gs, st) || st.reloc_heavy) ? 1 : 0;
            __syncthreads();
            halt = st.halt != 0;
        }
        if (threadIdx.x < SHARD_STATE_WORDS)
            reinterpret_cast<unsigned long long *>(&gst)[threadIdx.x] = reinterpret_cast<const unsigned long long *>(&st)[threadIdx.x];
    }
}

// ------------------------------------------------------------------------------------------
// Streamed loop for lists larger than L2, two launches per iteration and half of the sub-shards:
//   pass_local_kernel   the fused pass over the shards [shard0, shard0 + n) (photons that left their cell go to the
//                       shard's own region of the relocation list, as in the persistent loop)
//   event_local_kernel  one block per shard: re-locate those few photons through the bounding-box index (one warp per
//                       photon), finish them, shard arg-min, scattering event
// The host puts the two halves of the shards on two streams, half a period apart: while the blocks of one half run
// their events (latency-bound, 64 KB of registers per SM left free for them) the other half's pass streams its
// photons through the SMs, so the HBM pipe never waits for a scattering (mcrat_b200_run_frame).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PASS_THREADS, MCRAT_PASS_MINB) pass_local_kernel(DevCtx d, const int shard0, const int bps)
{
    const int s = shard0 + blockIdx.x / bps;
    const int b = blockIdx.x - (blockIdx.x / bps) * bps;
    double best_t = DBL_MAX;
    int best_i = INT_MAX;
    if (!loop_stopped(*d.gs, d.sh[s])) pass_body<true, true, PASS_THREADS>(d, d.sh[s], s, b, bps, 0, 0, best_t, best_i);
    block_argmin<PASS_THREADS>(best_t, best_i);
    if (threadIdx.x == 0) {
        d.bm_t[s * bps + b] = best_t;
        d.bm_i[s * bps + b] = best_i;
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 2 : 4) event_local_kernel(DevCtx d, const int shard0, const int bps)
{
    const int s = shard0 + blockIdx.x;
    ShardState &st = d.sh[s];
    if (loop_stopped(*d.gs, st)) return;
    const int R = *(volatile int *)&st.reloc_n;
    if (R > 0) {
        relocate_shard<THREADS>(d, st, s, R);
        if (threadIdx.x == 0 && R > RELOC_HEAVY) d.gs->reloc_heavy_any = 1; // the host hands the next batch to K1b / K1c
    }
    event_body<THREADS>(d, s, st.first, R, bps, 0, 0.0, st);
}

// head of the time order only (step API calcMeanFreePath; single shard)
__global__ void __launch_bounds__(EVT_THREADS) head_kernel(DevCtx d, int nb)
{
    double bt = DBL_MAX;
    int bi = INT_MAX;
    for (int k = threadIdx.x; k < nb; k += EVT_THREADS)
        if (lex_less(d.bm_t[k], d.bm_i[k], bt, bi)) {
            bt = d.bm_t[k];
            bi = d.bm_i[k];
        }
    block_argmin<EVT_THREADS>(bt, bi);
    if (threadIdx.x == 0) {
        d.sh[0].head_idx = bi;
        d.sh[0].head_tts = bt;
    }
}
