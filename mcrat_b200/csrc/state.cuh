// state.cuh -- shard / global loop state, photon and cell columns, DevCtx.
// Part of the single translation unit mcrat_b200.cu (included there, in this order); not a stand-alone header.
#pragma once

// ------------------------------------------------------------------------------------------
// device-side data
// ------------------------------------------------------------------------------------------
enum : unsigned char { F_MOVABLE = 1, F_RECALC = 2 };

constexpr int MAX_DT = 16;         // pushes recorded by one event (1 + Klein-Nishina rejections)
constexpr int BLOCKMIN_CAP = 8192; // per-block arg-min slots (persistent loop: two per sub-shard at 4096 sub-shards)
constexpr int MAX_SHARDS = 4096;
constexpr int PERSISTENT_MAX_PHOTONS = 1 << 21; // above this the list no longer fits in L2: streamed loop, streaming cache hints
#ifndef MCRAT_SCAN_THREADS
#define MCRAT_SCAN_THREADS 128
#endif
// photons per thread / resident CTAs per SM (round 2, persistent work-counter launch + PTX containment test; 10^5
// photons x 2^20 cells, CUDA events; % = of 148 SMs x 64 FP64 lanes x 1965 MHz): 3-D 9 / 5: 35.34 ms = 95.6 % (10^6
// photons: 96.0 %), 9 / 6: 35.48, 8 / 6: 35.73, 7 / 7: 37.06, 12 / 4: 35.41 (10^6: 96.9 %); 2-D 8 / 6: 23.95 ms = 94.1 %,
// 7 / 5: 24.01, 7 / 6: 24.45, 6 / 7: 24.29, 12 / 4: 24.11.  Round 1 (static grid, C++ test): 3-D 39.85 ms, 2-D 24.89 ms.
#ifndef MCRAT_SCAN_P
#define MCRAT_SCAN_P 8
#endif
#ifndef MCRAT_SCAN_P3
#define MCRAT_SCAN_P3 9
#endif
#ifndef MCRAT_SCAN_TILE
#define MCRAT_SCAN_TILE 256
#endif
#ifndef MCRAT_SCAN_UNROLL
#define MCRAT_SCAN_UNROLL 8
#endif
#ifndef MCRAT_SCAN_ITEMS_PER_CTA
#define MCRAT_SCAN_ITEMS_PER_CTA 40 // work items per resident CTA the scan is cut into (tail <= 1 / this)
#endif
#ifndef MCRAT_SCAN_MINB
#define MCRAT_SCAN_MINB 6 // 2-D: resident CTAs per SM the register allocation aims at
#endif
#ifndef MCRAT_SCAN_MINB3
#define MCRAT_SCAN_MINB3 5 // 3-D
#endif
// 1: the containment test is written in PTX (chained DSETP + one predicated minimum); 0: plain C++ (A/B)
#ifndef MCRAT_SCAN_PTX_TEST
#define MCRAT_SCAN_PTX_TEST 1
#endif
#define MCRAT_PRAGMA_STR2(x) #x
#define MCRAT_PRAGMA_STR(x) MCRAT_PRAGMA_STR2(x)
constexpr int SCAN_THREADS = MCRAT_SCAN_THREADS;
constexpr int SCAN_P2 = MCRAT_SCAN_P;      // photons per thread held in registers, 2-D
constexpr int SCAN_P3 = MCRAT_SCAN_P3;     // ... 3-D
constexpr int SCAN_TILE = MCRAT_SCAN_TILE; // cells per shared-memory stage
#ifndef MCRAT_PASS_LOCAL_CTAS_PER_SM
#define MCRAT_PASS_LOCAL_CTAS_PER_SM 4 // pass blocks per SM of the interleaved streamed loop: leaves room for one event block
#endif
constexpr int FEW_RMAX = 128;  // relocating photons handled per pass of the cell-parallel scan
constexpr int RELOC_LIST_SCAN_MAX = 2048;

struct PhotonCols {
    double *r0, *r1, *r2, *p0, *p1, *p2, *p3, *c0, *c1, *c2, *c3, *s0, *s1, *s2, *s3, *nscatt, *weight, *tau, *tts;
    // derived columns, rewritten whenever p / tau are (store_momentum, store_tau): what the pass reads instead of
    // them.  v_k = (p_k * (1/p0)) * C_LIGHT is the reference's own intermediate of the push (Src/mclib.c:1074-1080),
    // ntau = -1/tau that of the free path (Src/mclib.c:683): same roundings, one division each per *change* of the
    // photon instead of per photon-iteration, and 24 + 8 bytes per photon-iteration instead of 32 + 8.
    double *v0, *v1, *v2, *ntau;
    // safe[i]: the value of the shard's path counter (ShardState.path) up to which photon i provably cannot have left
    // its cell or the domain, so that the pass may skip its containment re-check (0: always re-check); see safe_path()
    unsigned long long *safe;
    int *idx;
    unsigned char *flags;
    char *type;
};

struct CellCols {
    int n, n_padded;
    const double4 *geoA; // 2-D: (c0, c1, h0, h1); 3-D: (c0, c1, c2, h0)     h = 0.5 * size
    const double2 *geoB; // 3-D: (h1, h2)
    const double *r0, *r1, *r2, *v0, *v1, *v2, *dens, *dens_lab, *temp, *gamma, *B0, *B1, *B2;
    const double *s0, *s1, *s2; // r0_size, r1_size, r2_size as uploaded (cyclo-synchrotron emission: corners, volumes)
    double *k2; // K_2(1/theta) of the cell's temperature (Src/electron.c:215), filled on first use; 0 = not yet
    double dom[6];
    // optional two-level bounding-box index over consecutive cells (BOX_T cells per level-1 box,
    // BOX_T level-1 boxes per level-2 box): 2 doubles (lo, hi) per dimension, 3 dimensions stored
    const double *box1, *box2;
    int nbox1, nbox2;
};

constexpr int BOX_T = 32;

// One sub-shard = one "rank" of the reference: a contiguous range of photon slots with its own
// clock, its own time-ordered event sequence (shard-local arg-min, exactly as per MPI rank,
// Src/mcrat.c:139-164) and its own Philox streams.
struct ShardState {
    double time_now, remaining_time, last_time_step;
    double dt_list[MAX_DT];
    double head_tts;
    unsigned long long iter;
    unsigned long long path; // length of all pushes before the pending ones, in 1/PATH_SCALE cm, rounded up (path_units)
    long long scatt_cnt, reloc_total, slots, iters_done;
    unsigned long long last_event_draw; // draws the last event took from its Philox stream (step API: the pool replacement continues there)
    int n_dt, pushed_slot; // pushed_slot: global slot index
    int done, pause_cs, counted_stopped;
    int last_scattered_idx, head_idx; // global slot indices
    int first, count;                 // slot range
    int mini_slot;                    // persistent loop: the photon whose next pass the event block does itself, or -1
    int pad1_;
    int halt;                         // persistent loop: loop_stopped() as evaluated by the publishing block
    int reloc_heavy;                  // the last iteration re-located many photons: better served by K1b / K1c
    int pad0_;
    // ---- everything above is the shard's state proper (the persistent loop keeps a shared-memory copy of it
    // and writes it back once per iteration); below: words other blocks update with atomics, never copied ----
    unsigned int arrive;              // tickets drawn by blocks that finished their pass (monotonic)
    unsigned int gen;                 // iterations completed, published by the last arriver (monotonic)
    int reloc_n;                      // entries of this shard's region of the relocation list
    int pad_;
};
constexpr int SHARD_STATE_WORDS = offsetof(ShardState, arrive) / 8; // 8-byte words of the copied part
static_assert(offsetof(ShardState, arrive) % 8 == 0, "ShardState: copied part must be a whole number of 8-byte words");

constexpr int NF_LOG_CAP = 32;

struct GlobalState {
    int reloc_count[2];
    int error, not_found, n_stopped;
    int reloc_heavy_any;    // streamed loop (local re-location): a shard re-located more than RELOC_HEAVY photons in one iteration
    unsigned int scan_work; // K1: next work item (photon chunk, cell chunk); zeroed by the pass kernel that feeds the scan
    // persistent loop for lists larger than L2 (frame_stream_*_kernel): next pass item, event blocks resident, shards halted
    unsigned long long stream_work;
    int stream_evt_ready, stream_halted;
    int stream_state; // 0 undecided, 1 go (every event block is resident and the pass blocks have seen it), 2 abort (the two
                      // grids did not become co-resident in time -- a profiler serialising kernels, a shared device --:
                      // nothing has been touched, the host runs the streamed loop instead)
    int error_slot, error_site; // photon slot (or -1) and ERR_SITE_* of the first error raised (raise_error)
    long long cell_evals, box_evals, max_iters;
    long long ref_equiv_evals; // first-hit index + 1 summed over the photons of full rescans (what the reference's loop executes)
    unsigned long long replay_cursor, replay_base, replay_n;
    int abs_count, cs_scatt_count;
    double abs_weight;
    // cyclo-synchrotron bookkeeping of the driver, Src/mcrat.c:792-831
    int cs_max_photons;        // rebin threshold (max_photons of mc.par); INT_MAX: never
    int cs_scatt_num;          // scatt_cyclosynch_num_ph
    int cs_emitted;            // pool photons replaced on the device
    double cs_comptonized_w;   // n_comptonized
    // the first NF_LOG_CAP photons for which no containing cell exists since the host last read the log: slot and hydro
    // coordinates, what findContainingBlock writes to the rank's log file (Src/geometry.c:373-388)
    int nf_logged;
    int nf_slot[NF_LOG_CAP];
    double nf_h[3 * NF_LOG_CAP];
#ifdef MCRAT_TIMING
    long long dbg[32];         // SM-cycle accumulators of shard 0 (tools/loop_timing.py; not in the product build)
#endif
};

#ifdef MCRAT_TIMING
#define TSTAMP_DECL long long t_last__ = clock64()
#define TSTAMP(gsref, k)                                 \
    do {                                                 \
        long long now__ = clock64();                     \
        (gsref).dbg[k] += now__ - t_last__;              \
        t_last__ = now__;                                \
    } while (0)
#else
#define TSTAMP_DECL
#define TSTAMP(gsref, k)
#endif

// where a device-side error was raised (GlobalState.error_site)
enum { ERR_SITE_NONE = 0, ERR_SITE_PASS_TABLE = 1, ERR_SITE_PASS_VERIFY = 2, ERR_SITE_FINISH_TABLE = 3, ERR_SITE_MFP_TABLE = 4,
       ERR_SITE_MFP_REPLAY = 5, ERR_SITE_EVENT_REPLAY = 6, ERR_SITE_EVENT_MINIPASS = 7, ERR_SITE_LOOP_SPIN = 8,
       ERR_SITE_REBIN = 9, ERR_SITE_CS_EMIT = 10 };

// first error wins: many threads may fail in the same launch, the host reports the first one with its photon slot
__device__ __forceinline__ void raise_error(GlobalState *gs, int code, int slot, int site)
{
    if (atomicCAS(&gs->error, 0, code) == 0) {
        gs->error_slot = slot;
        gs->error_site = site;
    }
}

struct DevCtx {
    int dims, geom, stokes, tau_calc, cs, b_calc;
    double epsilon_b;
    uint32_t k0, k1; // k1 is XORed with the sub-shard's global id
    uint32_t shard_base;
    int replay;
    int cap;
    int nshards, shard_size, blocks_per_shard;
    int mj_rounds;    // warp-wide Maxwell-Juttner sampling: rounds of 64 trials before the sequential loop takes over
    int recheck_skip; // 0: every photon re-checks its cell in every pass; 1: skip while provably inside (safe_path);
                      // 2: decide as in 1 but re-check anyway and raise an error if a skipped photon had left (tests)
    const double *dom_dev; // cells.dom in global memory, for safe_path()
    double path_pad;  // bound on the rounding error of one push of a photon inside the domain, cm
    int stream_hints; // the list is larger than L2: photon columns are streamed past it (ld/st.global.cs) so that the
                      // cell geometry the pass gathers from stays resident
    PhotonCols ph;
    CellCols cells;
    HotTable table;
    ShardState *sh;
    GlobalState *gs;
    // relocation scratch
    int *reloc_slot;
    double *reloc_h0, *reloc_h1, *reloc_h2;
    int *reloc_best;
    int reloc_cap;
    // arg-min scratch, one entry per pass block
    double *bm_t;
    int *bm_i;
    // team kernel: cell index and cell temperature of the block's best candidate, so that the event block gets
    // what the scattering lane needs first together with the minima (the pass blocks' time is hidden, the event's is not)
    int *bm_idx;
    double *bm_temp;
    // replay
    const double *replay_buf;
    int *prefix_block;
};
