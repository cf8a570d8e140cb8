// event.cuh -- K3 scattering event (three warps), cyclo-synchrotron single emission.
// Part of the single translation unit mcrat_b200.cu (included there, in this order); not a stand-alone header.
#pragma once

// ------------------------------------------------------------------------------------------
// K3: event kernel -- shard-local arg-min, then photonEvent (Src/mclib.c:1107-1356) and the
// driver's bookkeeping (Src/mcrat.c:777-846).  One block per sub-shard; lane 0 runs the scatter.
// ------------------------------------------------------------------------------------------
constexpr int EVT_THREADS = 256;   // one shard / few shards: wide block for the list scans
constexpr int EVT_THREADS_MANY = 128; // many sub-shards: the three event warps + one, 4 events resident per SM

// Mailbox of the three-warp scattering event (shared memory).
//   warp 0 = the scattering lane: electron sampling, Klein-Nishina draws, the boosts -- the only
//            consumer of random numbers and the critical path;
//   warp 1 = Stokes chain: every rotation angle of stokesRotation (Src/mcrat_scattering.c:103-149)
//            is a function of momenta only, so they are evaluated here, two to four at a time, one
//            per lane on the same instruction stream; the Stokes vector itself enters warp 0 only
//            through (q, u) in the azimuth draw;
//   warp 2 = helper: everything that depends on velocities alone and would otherwise sit on the
//            critical path -- the candidate's pushed position and fluid velocity, the Lorentz
//            matrices of the boosts back (lorentzBoost's matrix depends on beta only), the
//            alignment rotation and the Fano matrix.
// Each piece is the reference's statement block, operation for operation: results are
// bit-identical to the single-lane form (single_scatter in device_math.cuh; the round-1 single-lane event is the
// reference build of the A/B harness, tools/ab_compare.py).
struct ScatterMail {
    double pre[64];         // first 64 uniforms of the event's Philox stream
    double zhat[3];
    double fb[3], nfb[3];   // fluid velocity (Src/mclib.c:1151-1174) and its negative
    double r[3];            // candidate position after this event's pushes
    double p[4];            // lab 4-momentum of the candidate
    double pc[4];           // fluid-frame 4-momentum (photon.comv_p*)
    double pcb[4];          // the same after lorentzBoost had it as input (renormalised in place if beta = 0)
    double el_v[3], nel_v[3];
    double php[4];          // electron rest frame, before the scatter (`orig`)
    double out[4];          // electron rest frame, after
    double outb[4];         // `out` after lorentzBoost had it as input
    double pc_new[4];       // fluid frame, after (as the last stokesRotation of singleScatter sees it)
    double pc_fin[4];       // fluid frame, after the lab boost had it as input
    double p_new[4];        // lab frame, after
    double fano[5];
    double q, u;
    BoostMat Lf, Le;        // boosts by -fluid_beta and by -el_v
    ScatterRot rot;
    ElRot erot;             // rotateElectron's angles (functions of the comoving photon only)
    int occurred;
    unsigned char flags;
};

constexpr int SCATTER_THREADS = 96; // warps 0..2 of the event block

// Early hand-over (persistent loop): once a candidate is accepted by the Klein-Nishina test, everything the
// other photons' next pass needs -- the pushes of this event, the new clock, the iteration number -- is final,
// while half of the event (azimuth, outgoing photon, boosts back, Stokes chain) still lies ahead and touches only
// the scattered photon.  The helper warp therefore does the driver's bookkeeping (Src/mcrat.c:781-846) right there,
// publishes the shard state and releases the pass blocks; the event block finishes the scatter and runs the
// scattered photon's next pass itself ("mini-pass").
struct EarlyRelease {
    int enabled;        // set by thread 0 before the scatter
    ShardState *gst;    // global copy of the shard state
    unsigned gen_value; // value to release on gst->gen
    int bm_index;       // slot of d.bm_t / d.bm_i that receives the mini-pass result
    int step_mode;
    int n_dt, ph_index;
    double scatt_time;
    int released;       // out: the state has been published
    // cluster team (frame_loop_cluster_kernel): the state is pushed into the pass blocks' shared memory and their
    // generation barriers are arrived on remotely; 0 pass blocks = not a cluster team
    int cl_bps;
    uint32_t cl_st_addr, cl_bar_addr; // shared-space addresses of `st` and the generation mbarrier (the same in every block)
    int mini_reloc;     // out: the mini-pass put its photon on the shard's relocation list
};

// distributed shared memory of a thread-block cluster
__device__ __forceinline__ uint32_t cluster_map(uint32_t smem_addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_st_u64(uint32_t addr, unsigned long long v)
{
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void cluster_st_u32(uint32_t addr, unsigned v)
{
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_mbar_arrive(uint32_t bar_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}

// one warp: the shard state into the shared memory of pass blocks 0 .. bps-1 of this cluster, then an arrival on each one's
// generation barrier (release at cluster scope: the pushes are visible to whoever wakes up on it)
__device__ __forceinline__ void cluster_publish_warp(const ShardState &st, const int bps, const uint32_t st_addr, const uint32_t bar_addr,
                                                     const int lane)
{
    for (int b = 0; b < bps; ++b)
        for (int k = lane; k < SHARD_STATE_WORDS; k += 32)
            cluster_st_u64(cluster_map(st_addr + 8u * (uint32_t)k, (uint32_t)b), reinterpret_cast<const unsigned long long *>(&st)[k]);
    __syncwarp();
    __threadfence(); // the event block's global writes (scattered photon, relocation list) before the wake-up
    if (lane < bps) cluster_mbar_arrive(cluster_map(bar_addr, (uint32_t)lane));
    __syncwarp();
}

// the driver's bookkeeping after photonEvent returned (Src/mcrat.c:783-787, 834-846), without the cyclo-synchrotron part
__device__ __forceinline__ void event_bookkeeping(ShardState &st, int n_dt, int ph_index, double scatt_time, int step_mode)
{
    st.n_dt = n_dt;
    st.last_scattered_idx = ph_index;
    st.last_time_step = scatt_time;
    st.iter += 1;
    st.iters_done += 1;
    if (step_mode == 0) {
        st.time_now += scatt_time;
        st.remaining_time -= scatt_time;
        if (!(st.remaining_time > 0)) st.done = 1;
    }
}

__device__ __forceinline__ void event_count_stopped(GlobalState &gs, ShardState &st)
{
    if ((st.done | st.pause_cs | (gs.max_iters >= 0 && st.iters_done >= gs.max_iters)) && !st.counted_stopped) {
        st.counted_stopped = 1;
        atomicAdd(&gs.n_stopped, 1);
    }
}

__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v);

__device__ __forceinline__ void trio_bar() { asm volatile("bar.sync 1, 96;" ::: "memory"); }
__device__ __forceinline__ void duo_bar() { asm volatile("bar.sync 2, 64;" ::: "memory"); } // warps 1 and 2
// warp 1 hands rotateElectron's angles to warp 0 without waiting for it
__device__ __forceinline__ void erot_arrive() { asm volatile("bar.arrive 3, 64;" ::: "memory"); }
__device__ __forceinline__ void erot_wait() { asm volatile("bar.sync 3, 64;" ::: "memory"); }

// up to four Stokes angles at once, one per lane; returns sin/cos(2 phi) of this lane's angle
__device__ __forceinline__ void lane_angle(const double *k1, const double *a, const double *k2, const double *b, bool active,
                                           double &sn, double &cs)
{
    sn = 0;
    cs = 1;
    if (active) {
        double phi = stokes_angle4(k1, a, k2, b);
        sincos(2 * phi, &sn, &cs);
    }
}

__device__ __forceinline__ void rot_from_lane(double sn, double cs, int src, double *s)
{
    double a = __shfl_sync(0xffffffffu, sn, src), c = __shfl_sync(0xffffffffu, cs, src);
    muller_rotation_sc(a, c, s);
}

// photonEvent's body for one candidate (Src/mclib.c:1138-1333); threads 0..95 of the block call
// this together (STOKES_SWITCH ON).
// cand_idx: the candidate's cell (-2: not known yet); cand_temp: that cell's temperature (< 0: not known yet)
__device__ void scatter_candidate_3w(DevCtx &d, ShardState &st, EventRng &rng_sh, ScatterMail &m, const int i, int cand_idx,
                                     const double cand_temp, int n_dt, int *event_did_occur, EarlyRelease &early)
{
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stokes = d.stokes;
    // warp-0 lane-0 state carried across stages
    double theta = 0;
    KnTheta kn;
    EventRng rng;
    double s[4] = {0, 0, 0, 0}; // warp 1, replicated in its lanes
    double sn = 0, cs = 1;
    CellState cell; // warp 2 lane 0: the candidate's cell, kept for the mini-pass
    cell.v0 = cell.v1 = cell.v2 = cell.r0 = cell.r1 = cell.r2 = cell.gamma = cell.dens_lab = cell.temp = 0;
    int cell_idx = -1;
#ifdef MCRAT_TIMING
    const bool tm__ = (w == 0 && lane == 0 && st.first == 0);
    GlobalState &gsr__ = *d.gs;
#define T2W(k) if (tm__) TSTAMP(gsr__, k)
#else
#define T2W(k)
#endif
    TSTAMP_DECL;

    // ---- stage A/B: electron + boost into its rest frame | position, fluid velocity, lab -> fluid rotation ----
    if (w == 0) {
        if (!rng_sh.replay) {
            double a, b;
            philox_doubles((uint32_t)lane, (uint32_t)rng_sh.iter, (uint32_t)(rng_sh.iter >> 32), 1u, rng_sh.k0, rng_sh.k1, a, b);
            m.pre[2 * lane] = a;
            m.pre[2 * lane + 1] = b;
        }
        __syncwarp();
        // Maxwellian branch: the three gaussians at once, one candidate pair per lane
        double temp = 0;
        if (lane == 0) temp = (cand_temp >= 0) ? cand_temp : d.cells.temp[cand_idx == -2 ? d.ph.idx[i] : cand_idx];
        temp = __shfl_sync(0xffffffffu, temp, 0);
        double g3[3] = {0, 0, 0};
        int used = 0;
        if (!rng_sh.replay && temp < 1e7) used = warp_gaussians3(m.pre, 64, rng_sh.draw, sqrt(K_B * temp / M_EL), g3);
        // Maxwell-Juttner branch: K_2(1/theta) from the per-cell cache, then 64 rejection trials per round
        double gamma = 1, k2 = 0;
        uint64_t used_mj = 0;
        const int MJ_ROUNDS = d.mj_rounds; // x 64 trials, then sequentially (never in practice; MCRAT_B200_MJ_ROUNDS for tests)
        if (temp >= 1e7) {
            const double factor = K_B * temp / (M_EL * C_LIGHT * C_LIGHT);
            if (lane == 0) {
                const int cell = (cand_idx == -2) ? d.ph.idx[i] : cand_idx;
                k2 = d.cells.k2[cell];
                if (!(k2 > 0)) { // not yet evaluated for this cell in this hydro frame (or underflowed: evaluated again)
                    k2 = bessel_K2(1.0 / factor);
                    d.cells.k2[cell] = k2;
                }
            }
            k2 = __shfl_sync(0xffffffffu, k2, 0);
            if (!rng_sh.replay) {
                used_mj = warp_mj_gamma(rng_sh.k0, rng_sh.k1, rng_sh.iter, rng_sh.draw, factor, k2, MJ_ROUNDS, gamma);
                if (!used_mj) gamma = 1;
            }
        }
        if (lane == 0) {
            rng = rng_sh;
            rng.pre = m.pre;
            rng.npre = rng_sh.replay ? 0 : 64;
            T2W(8);
            if (used) {
                rng.draw += (uint64_t)used;
                gamma = maxwellian_gamma(g3);
            } else if (used_mj) {
                rng.draw += used_mj;
            } else {
                if (temp >= 1e7 && !rng_sh.replay) rng.draw += 128ull * (uint64_t)MJ_ROUNDS; // those trials were all rejected
                gamma = sample_thermal_electron(temp, rng, k2);
            }
        }
        erot_wait(); // warp 1 has the rotation angles ready long before
        if (lane == 0) {
            double pc[4] = {d.ph.c0[i], d.ph.c1[i], d.ph.c2[i], d.ph.c3[i]};
            double el[4], el_v[3], php[4];
            thermal_electron_from_gamma(el, gamma, m.erot, rng);
            T2W(9);
            scatter_stage_boost(el, pc, el_v, php);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                m.el_v[k] = el_v[k];
                m.nel_v[k] = (-1 * el_v[k]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                m.pcb[k] = pc[k];
                m.php[k] = php[k];
            }
        }
        __syncwarp();
    } else if (w == 2) {
        if (lane == 0) {
            const unsigned char flags = d.ph.flags[i];
            double p[4] = {d.ph.p0[i], d.ph.p1[i], d.ph.p2[i], d.ph.p3[i]};
            double r0 = d.ph.r0[i], r1 = d.ph.r1[i], r2 = d.ph.r2[i];
            if (flags & F_MOVABLE) apply_pushes(st, n_dt, p[0], p[1], p[2], p[3], r0, r1, r2);
            cell_idx = (cand_idx == -2) ? d.ph.idx[i] : cand_idx;
            cell = load_cell_state(d.cells, cell_idx);
            double fb[3];
            fluid_beta_of(d, cell, r0, r1, fb);
            m.zhat[0] = 0; m.zhat[1] = 0; m.zhat[2] = 1;
            m.r[0] = r0; m.r[1] = r1; m.r[2] = r2;
            m.flags = flags;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                m.fb[k] = fb[k];
                m.nfb[k] = -1 * fb[k];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) m.p[k] = p[k];
            m.pc[0] = d.ph.c0[i]; m.pc[1] = d.ph.c1[i]; m.pc[2] = d.ph.c2[i]; m.pc[3] = d.ph.c3[i];
        }
        __syncwarp();
        duo_bar();
        if (lane == 0) {
            double nfb[3] = {m.nfb[0], m.nfb[1], m.nfb[2]};
            boost_matrix(nfb, m.Lf);
        }
        __syncwarp();
    } else {
        if (lane == 0) {
            double pc[4] = {d.ph.c0[i], d.ph.c1[i], d.ph.c2[i], d.ph.c3[i]};
            ElRot er;
            electron_rot_angles(pc, er);
            m.erot = er;
        }
        __syncwarp();
        erot_arrive();
        if (stokes) { s[0] = d.ph.s0[i]; s[1] = d.ph.s1[i]; s[2] = d.ph.s2[i]; s[3] = d.ph.s3[i]; }
        duo_bar();
        if (stokes) {
            // stokesRotation(fluid_beta, p, comv_p), Src/mclib.c:1190-1196
            lane_angle(lane == 0 ? m.p + 1 : m.pc + 1, lane == 0 ? m.zhat : m.fb, lane == 0 ? m.p + 1 : m.pc + 1,
                       lane == 0 ? m.fb : m.zhat, lane < 2, sn, cs);
            rot_from_lane(sn, cs, 0, s);
            rot_from_lane(sn, cs, 1, s);
        }
    }
    T2W(10);
    trio_bar();
    T2W(11);
    // ---- stage C: Klein-Nishina accept / polar angle | fluid -> electron-frame rotation | alignment, boost matrix ----
    if (w == 0) {
        if (lane == 0) m.occurred = kn_accept_theta(theta, m.php[0], kn, rng);
        __syncwarp();
    } else if (w == 1) {
        if (stokes) {
            // stokesRotation(el_v, ph_comov, ph_p_prime), Src/mcrat_scattering.c:245-253
            lane_angle(lane == 0 ? m.pcb + 1 : m.php + 1, lane == 0 ? m.zhat : m.el_v, lane == 0 ? m.pcb + 1 : m.php + 1,
                       lane == 0 ? m.el_v : m.zhat, lane < 2, sn, cs);
            rot_from_lane(sn, cs, 0, s);
            rot_from_lane(sn, cs, 1, s);
        }
        if (lane == 0) {
            m.q = s[1];
            m.u = s[2];
        }
    } else {
        if (lane == 0) {
            double php[4] = {m.php[0], m.php[1], m.php[2], m.php[3]};
            ScatterRot rot;
            scatter_stage_align(php, rot);
            m.rot = rot;
            double nel_v[3] = {m.nel_v[0], m.nel_v[1], m.nel_v[2]};
            boost_matrix(nel_v, m.Le);
        }
        __syncwarp();
    }
    T2W(12);
    trio_bar();
    T2W(13);
    if (!m.occurred) { // Klein-Nishina rejection: the draws are spent, nothing else changes
        if (w == 0 && lane == 0) {
            rng.pre = nullptr;
            rng.npre = 0;
            rng_sh = rng;
        }
        return;
    }
    // ---- stage D: azimuth + outgoing photon | -- | bookkeeping and early release of the pass blocks ----
    if (w == 0) {
        if (lane == 0) {
            double phi = kn_phi(stokes, kn, m.q, m.u, rng);
            double out[4];
            ScatterRot rot = m.rot;
            scatter_stage_out(m.php[0], kn.st, kn.ct, phi, rot, out);
#pragma unroll
            for (int k = 0; k < 4; ++k) m.out[k] = out[k];
        }
        __syncwarp();
    } else if (w == 2 && early.enabled) {
        if (lane == 0) {
            st.pushed_slot = i; // the accepted candidate is at its pushed position already (Src/mclib.c:1138)
            st.scatt_cnt += 1;
            event_bookkeeping(st, early.n_dt, early.ph_index, early.scatt_time, early.step_mode);
            event_count_stopped(*d.gs, st);
            st.halt = (loop_stopped(*d.gs, st) || st.reloc_heavy) ? 1 : 0;
            st.mini_slot = st.halt ? -1 : i; // a halted shard leaves the photon as photonEvent left it
        }
        __syncwarp();
        for (int k = lane; k < SHARD_STATE_WORDS; k += 32)
            reinterpret_cast<unsigned long long *>(early.gst)[k] = reinterpret_cast<const unsigned long long *>(&st)[k];
        __syncwarp();
        if (early.cl_bps > 0) {
            cluster_publish_warp(st, early.cl_bps, early.cl_st_addr, early.cl_bar_addr, lane);
            if (lane == 0) early.released = 1;
        } else if (lane == 0) {
            __threadfence();
            st_release_u32(&early.gst->gen, early.gen_value);
            early.released = 1;
        }
        __syncwarp();
    }
    T2W(14);
    trio_bar();
    T2W(15);
    // ---- stage E: boosts back to the fluid and lab frames | scattering-plane angles | Fano matrix ----
    if (w == 0) {
        if (lane == 0) {
            double out[4] = {m.out[0], m.out[1], m.out[2], m.out[3]};
            double pcn[4], pn[4];
            boost_apply(m.Le, out, pcn, true); // Src/mcrat_scattering.c:455-463
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                m.outb[k] = out[k];
                m.pc_new[k] = pcn[k];
            }
            boost_apply(m.Lf, pcn, pn, true); // Src/mclib.c:1262-1265
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                m.pc_fin[k] = pcn[k];
                m.p_new[k] = pn[k];
            }
        }
        __syncwarp();
    } else if (w == 1) {
        // lane 0: into the scattering plane (Src/mcrat_scattering.c:402-405); lane 1: back out of it (:438-447)
        if (stokes)
            lane_angle(lane == 0 ? m.php + 1 : m.out + 1, lane == 0 ? m.zhat : m.php + 1, m.out + 1,
                       lane == 0 ? m.php + 1 : m.zhat, lane < 2, sn, cs);
    } else {
        if (lane == 0 && stokes) {
            double f[5];
            scatter_stage_fano(m.php, m.out, f);
#pragma unroll
            for (int k = 0; k < 5; ++k) m.fano[k] = f[k];
        }
        __syncwarp();
    }
    T2W(16);
    trio_bar();
    T2W(17);
    // ---- stage F: the remaining angles, the Stokes chain applied in order, write-back ----
    if (w == 1 && !stokes) {
        // unpolarised run (STOKES_SWITCH OFF): no Stokes chain; warp 1 only supplied rotateElectron's angles
    } else if (w == 1) {
        // lane 0 / 1: stokesRotation(-el_v, out, pc_new), Src/mcrat_scattering.c:465-473 (`out` as lorentzBoost
        // left it); lane 2 / 3: stokesRotation(-fluid_beta, pc_fin, p_new), Src/mclib.c:1267-1287
        double sn2, cs2;
        const double *k = lane == 0 ? m.outb + 1 : (lane == 1 ? m.pc_new + 1 : (lane == 2 ? m.pc_fin + 1 : m.p_new + 1));
        const double *a = lane == 0 ? m.zhat : (lane == 1 ? m.nel_v : (lane == 2 ? m.zhat : m.nfb));
        const double *b = lane == 0 ? m.nel_v : (lane == 1 ? m.zhat : (lane == 2 ? m.nfb : m.zhat));
        lane_angle(k, a, k, b, lane < 4, sn2, cs2);
        rot_from_lane(sn, cs, 0, s);
        {
            double f[5] = {m.fano[0], m.fano[1], m.fano[2], m.fano[3], m.fano[4]};
            fano_apply(f, s);
        }
        rot_from_lane(sn, cs, 1, s);
        rot_from_lane(sn2, cs2, 0, s);
        rot_from_lane(sn2, cs2, 1, s);
        rot_from_lane(sn2, cs2, 2, s);
        rot_from_lane(sn2, cs2, 3, s);
        if (lane == 0) {
            d.ph.s0[i] = s[0];
            d.ph.s1[i] = s[1];
            d.ph.s2[i] = s[2];
            d.ph.s3[i] = s[3];
        }
    } else if (w == 0 && lane == 0) {
        store_momentum(d.ph, i, m.p_new[0], m.p_new[1], m.p_new[2], m.p_new[3]);
        d.ph.c0[i] = m.pc_fin[0]; d.ph.c1[i] = m.pc_fin[1]; d.ph.c2[i] = m.pc_fin[2]; d.ph.c3[i] = m.pc_fin[3];
        d.ph.nscatt[i] = d.ph.nscatt[i] + 1;
        d.ph.flags[i] = m.flags | F_RECALC;
        // this photon is already at its pushed position: the next pass must not push it again
        d.ph.r0[i] = m.r[0];
        d.ph.r1[i] = m.r[1];
        d.ph.r2[i] = m.r[2];
        d.ph.safe[i] = 0;
        if (!early.released) {
            st.pushed_slot = i;
            st.scatt_cnt += 1;
        }
        *event_did_occur = 1;
        rng.pre = nullptr;
        rng.npre = 0;
        rng_sh = rng;
        T2W(18);
    }
    // ---- mini-pass: the scattered photon's share of the next pass (pass_body for one photon), by the helper warp ----
    if (early.released && st.mini_slot == i) {
        double t_next = 1e12 / C_LIGHT, tau_next = 0, h0 = 0, h1 = 0, h2 = 0;
        int state = 0; // 0: out of the domain, 1: still in its cell (t_next, tau_next valid), 2: left its cell
        if (w == 2 && lane == 0) {
            const int ndim3 = (d.dims == D_THREE);
            coord_to_hydro(d.dims, d.geom, m.r[0], m.r[1], m.r[2], h0, h1, h2);
            bool in_domain;
            if (!ndim3)
                in_domain = ((h1 < d.cells.dom[3]) && (h1 > d.cells.dom[2]) && (h0 < d.cells.dom[1]) && (h0 > d.cells.dom[0]));
            else
                in_domain = ((h2 < d.cells.dom[5]) && (h2 > d.cells.dom[4]) && (h1 < d.cells.dom[3]) && (h1 > d.cells.dom[2]) &&
                             (h0 < d.cells.dom[1]) && (h0 > d.cells.dom[0]));
            if (in_domain) {
                if (in_cell(ndim3, d.cells, cell_idx, h0, h1, h2)) {
                    int terr = 0;
                    const FallbackRng fr = {d.k0, d.k1 ^ (d.shard_base + (uint32_t)(early.gst - d.sh)), (uint32_t)(i - st.first),
                                            st.iter, d.replay};
                    tau_next = optical_depth(d.dims, d.geom, d.tau_calc, d.table, cell, m.r[0], m.r[1], m.p_new[1], m.p_new[2],
                                             m.p_new[3], m.pc_fin[0], &terr, &fr);
                    if (terr) raise_error(d.gs, MCRAT_B200_ERR_TABLE, i, ERR_SITE_EVENT_MINIPASS);
                    const uint32_t k1 = d.k1 ^ (d.shard_base + (uint32_t)(early.gst - d.sh));
                    const double xi = philox_mfp_uniform(d.k0, k1, st.iter, (uint32_t)(i - st.first));
                    t_next = free_path_time(tau_next, xi);
                    state = 1;
                } else {
                    state = 2;
                }
            }
        }
        trio_bar(); // warp 0 has written the photon's new columns
        if (w == 2 && lane == 0) {
            double bt = DBL_MAX;
            int bi = INT_MAX;
            if (state == 1) {
                store_tau(d.ph, i, tau_next);
                d.ph.flags[i] = m.flags & ~F_RECALC;
                d.ph.tts[i] = t_next;
                bt = t_next;
                bi = i;
            } else if (state == 2) {
                early.mini_reloc = 1;
                const int pos = st.first + atomicAdd(&early.gst->reloc_n, 1);
                d.reloc_slot[pos] = i;
                d.reloc_h0[pos] = h0;
                d.reloc_h1[pos] = h1;
                d.reloc_h2[pos] = h2;
                d.reloc_best[pos] = INT_MAX;
            } else {
                d.ph.idx[i] = -1; // Src/mclib.c:589-595 (safe[i] is 0 since the write-back)
                d.ph.tts[i] = t_next;
                bt = t_next;
                bi = i;
            }
            d.bm_t[early.bm_index] = bt;
            d.bm_i[early.bm_index] = bi;
            d.bm_idx[early.bm_index] = (state == 1) ? cell_idx : -1;
            d.bm_temp[early.bm_index] = cell.temp;
        }
    }
}

// getMagneticFieldMagnitude, Src/mc_cyclosynch.c:78-92
__device__ __forceinline__ double cell_b_field(const DevCtx &d, int idx)
{
    if (d.b_calc == B_TOTAL_E || d.b_calc == B_INTERNAL_E) {
        double el_dens = d.cells.dens[idx] / M_P;
        return calc_b(d.b_calc, d.epsilon_b, el_dens, d.cells.temp[idx]);
    }
    if (d.dims == D_TWO) {
        double b0 = d.cells.B0[idx], b1 = d.cells.B1[idx];
        return sqrt(b0 * b0 + b1 * b1);
    }
    double b0 = d.cells.B0[idx], b1 = d.cells.B1[idx], b2 = d.cells.B2[idx];
    return sqrt(b0 * b0 + b1 * b1 + b2 * b2);
}

// photonEmitCyclosynch with inject_single_switch == 1 (Src/mc_cyclosynch.c:1465-1555): a pool
// photon that scattered is replaced by a fresh one at the cyclotron frequency of its cell, placed
// into the first null slot of the list (addToPhotonList, Src/photons.c:132-160), and the scattered
// photon is re-positioned at random inside the cell (:1541-1553).
__device__ void cs_emit_single(DevCtx &d, EventRng &rng, int scatt, int slot)
{
    const int i = d.ph.idx[scatt];
    const int ndim3 = (d.dims == D_THREE);
    const double nu_c = calc_cyclotron_freq(cell_b_field(d, i));
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
    const double cr0 = d.cells.r0[i], cr1 = d.cells.r1[i], cr2 = d.cells.r2[i];
    if (ndim3)
        hydro_vector_to_cartesian(d.dims, d.geom, boost, d.cells.v0[i], d.cells.v1[i], d.cells.v2[i], cr0, cr1, cr2);
    else if (d.dims == D_TWO_POINT_FIVE)
        hydro_vector_to_cartesian(d.dims, d.geom, boost, d.cells.v0[i], d.cells.v1[i], d.cells.v2[i], cr0, cr1, position_phi);
    else
        hydro_vector_to_cartesian(d.dims, d.geom, boost, d.cells.v0[i], d.cells.v1[i], 0, cr0, cr1, position_phi);
    boost[0] *= -1;
    boost[1] *= -1;
    boost[2] *= -1;
    lorentz_boost(boost, p_comv, l_boost, true);
    if (ndim3)
        hydro_coord_to_mcrat(d.dims, d.geom, pos, cr0, cr1, cr2);
    else
        hydro_coord_to_mcrat(d.dims, d.geom, pos, cr0, cr1, position_phi);
    store_momentum(d.ph, slot, l_boost[0], l_boost[1], l_boost[2], l_boost[3]);
    d.ph.c0[slot] = p_comv[0]; d.ph.c1[slot] = p_comv[1]; d.ph.c2[slot] = p_comv[2]; d.ph.c3[slot] = p_comv[3];
    d.ph.r0[slot] = pos[0]; d.ph.r1[slot] = pos[1]; d.ph.r2[slot] = pos[2];
    d.ph.safe[slot] = 0;
    d.ph.s0[slot] = 1; d.ph.s1[slot] = 0; d.ph.s2[slot] = 0; d.ph.s3[slot] = 0;
    d.ph.nscatt[slot] = 0;
    d.ph.weight[slot] = d.ph.weight[scatt];
    d.ph.idx[slot] = i;
    d.ph.type[slot] = 'p';
    d.ph.flags[slot] = F_RECALC; // pool photons do not move (Src/mclib.c:1070)
    d.ph.tts[slot] = 0;
    store_tau(d.ph, slot, 0);
    // new random position of the scattered photon inside its cell
    const double4 a = d.cells.geoA[i];
    double size0, size1, size2 = 0;
    if (!ndim3) {
        size0 = 2 * a.z;
        size1 = 2 * a.w;
    } else {
        const double2 b = d.cells.geoB[i];
        size0 = 2 * a.w;
        size1 = 2 * b.x;
        size2 = 2 * b.y;
    }
    const double pr = rng.uniform_pos() * (size0) - (size0) / 2.0;
    const double pr2 = rng.uniform_pos() * (size1) - (size1) / 2.0;
    if (ndim3) {
        const double pr3 = rng.uniform_pos() * (size2) - (size2) / 2.0;
        hydro_coord_to_mcrat(d.dims, d.geom, pos, cr0 + pr, cr1 + pr2, cr2 + pr3);
    } else {
        hydro_coord_to_mcrat(d.dims, d.geom, pos, cr0 + pr, cr1 + pr2, position_phi);
    }
    d.ph.safe[scatt] = 0;
    d.ph.r0[scatt] = pos[0];
    d.ph.r1[scatt] = pos[1];
    d.ph.r2[scatt] = pos[2];
}

// step_mode 0: frame loop (driver bookkeeping included); 1: photonEvent only (dt_max given)
// blockmin_valid: the pass wrote per-block minima for this iteration (fused path / step API)
// `early_gst` != nullptr (persistent loop): publish the state to *early_gst and release `early_gen` on its generation
// word as soon as a candidate is accepted; d.bm_*[early_bm] receives the scattered photon's mini-pass.  Returns
// whether that happened (else the caller publishes after the event).
// R is the relocation count of the whole list in the streamed loop: the entries of this shard are found by reading all
// R slots (4 B each, L2-resident) -- cheaper than reading the shard's own time column (8 B per photon) unless the list is
// long compared with the shard.  (A fixed cap of 2048 made every event block of a 10^7-photon list re-read its whole
// column once 2e-4 of the photons changed cell in an iteration.)
__device__ __forceinline__ bool reloc_list_cheaper(int R, int shard_count)
{
    return R <= RELOC_LIST_SCAN_MAX || R <= shard_count / 2;
}

template <int EVT_THREADS>
__device__ __forceinline__ bool event_body(DevCtx &d, const int s, const int reloc_base, const int R, int nb_per_shard,
                                           int step_mode, double dt_max_arg, ShardState &st, ShardState *early_gst = nullptr,
                                           unsigned early_gen = 0, int early_bm = 0, const bool have_pre = false,
                                           const double pre_t = DBL_MAX, const int pre_i = INT_MAX, const int pre_idx = -2,
                                           const double pre_temp = 0, const int cl_bps = 0, const uint32_t cl_st_addr = 0,
                                           const uint32_t cl_bar_addr = 0, int *mini_reloc_out = nullptr)
{
    GlobalState &gs = *d.gs;
    __shared__ EarlyRelease early;

    __shared__ double sh_cand_t;
    __shared__ double sh_cand_temp;
    __shared__ int sh_cand_i, sh_cand_idx, sh_cand_known, sh_finished; // sh_cand_known: idx and temperature are in shared memory
    // ---- head of this shard's time order ----
    double bt = DBL_MAX;
    int bi = INT_MAX;
    if (nb_per_shard > 0) {
        if (have_pre) { // the caller requested this thread's entry of the block minima together with other loads
            bt = pre_t;
            bi = pre_i;
        } else {
            for (int k = threadIdx.x; k < nb_per_shard; k += EVT_THREADS) {
                const int q = s * nb_per_shard + k;
                if (lex_less(d.bm_t[q], d.bm_i[q], bt, bi)) {
                    bt = d.bm_t[q];
                    bi = d.bm_i[q];
                }
            }
        }
        // photons relocated in this iteration got their time in finish
        if (R > 0 && reloc_list_cheaper(R, st.count)) {
            for (int j = threadIdx.x; j < R; j += EVT_THREADS) {
                const int i = d.reloc_slot[reloc_base + j];
                if (i >= st.first && i < st.first + st.count) {
                    double t = d.ph.tts[i];
                    if (lex_less(t, i, bt, bi)) {
                        bt = t;
                        bi = i;
                    }
                }
            }
        } else if (R > 0) {
            for (int j = threadIdx.x; j < st.count; j += EVT_THREADS) {
                const int i = st.first + j;
                double t = d.ph.tts[i];
                if (lex_less(t, i, bt, bi)) {
                    bt = t;
                    bi = i;
                }
            }
        }
    } else {
        for (int j = threadIdx.x; j < st.count; j += EVT_THREADS) {
            const int i = st.first + j;
            double t = d.ph.tts[i];
            if (lex_less(t, i, bt, bi)) {
                bt = t;
                bi = i;
            }
        }
    }
    block_argmin<EVT_THREADS>(bt, bi);

    __shared__ EventRng rng_sh;
    __shared__ ScatterMail mail;
    __shared__ double old_scatt_time, scatt_time, dt_max;
    __shared__ int n_dt, ph_index, sh_try, sh_event;
    if (threadIdx.x == 0) {
        // the candidate's cell index: delivered with the block minima (team kernel), else requested first so that
        // the loads of the set-up below travel with it
        sh_cand_known = 0;
        sh_cand_idx = have_pre ? -2 : ((bi != INT_MAX) ? d.ph.idx[bi] : -1);
        sh_cand_t = bt;
        sh_cand_i = bi;
        sh_finished = 0;
        st.head_idx = bi;
        st.head_tts = bt;
        dt_max = (step_mode == 0) ? st.remaining_time : dt_max_arg;
        old_scatt_time = 0;
        scatt_time = 0;
        n_dt = 0;
        ph_index = bi;
        fold_path(st, d.path_pad); // the pushes of the last event have been applied by the pass that led here
        st.n_dt = 0;
        st.pushed_slot = -1;
        early.enabled = 0;
        early.released = 0;
        early.cl_bps = cl_bps;
        early.cl_st_addr = cl_st_addr;
        early.cl_bar_addr = cl_bar_addr;
        early.mini_reloc = 0;
        early.gst = early_gst;
        early.gen_value = early_gen;
        early.bm_index = early_bm;
        early.step_mode = step_mode;
        rng_sh.replay = d.replay;
        rng_sh.k0 = d.k0;
        rng_sh.k1 = d.k1 ^ (d.shard_base + (uint32_t)s);
        rng_sh.iter = st.iter;
        rng_sh.draw = 0;
        rng_sh.buf = d.replay_buf;
        rng_sh.pos = d.replay ? gs.replay_cursor : 0; // global loads only the parity harness needs
        rng_sh.n = d.replay ? gs.replay_n : 0;
        rng_sh.exhausted = 0;
        rng_sh.pre = nullptr;
        rng_sh.npre = 0;
        if (step_mode == 0) st.slots += st.count;
        if (step_mode == 0 && !(bt < dt_max)) {
            // Src/mcrat.c:834-846: nothing scatters before the next hydro frame
            st.time_now += st.remaining_time;
            st.dt_list[0] = st.remaining_time;
            st.n_dt = 1;
            st.last_time_step = st.remaining_time;
            st.remaining_time = 0;
            st.done = 1;
            st.iter += 1;
            st.iters_done += 1;
            if (!st.counted_stopped) {
                st.counted_stopped = 1;
                atomicAdd(&gs.n_stopped, 1);
            }
            sh_finished = 1;
        }
    }
    __syncthreads();
    if (sh_finished) return false;
    if (have_pre && pre_i == sh_cand_i && pre_idx >= 0) { // the thread whose entry won hands over what came with it
        sh_cand_idx = pre_idx;
        sh_cand_temp = pre_temp;
        sh_cand_known = 1;
    }

    // ---- photonEvent: walk candidates in ascending time, Src/mclib.c:1128-1339 ----
    while (true) {
        // The reference's walk has no bound on the number of rejected candidates (Src/mclib.c:1128-1339), the push
        // list has MAX_DT entries: when it is full, the recorded pushes are applied here to every photon of the shard
        // (one by one, in order -- exactly what the next pass would have done) and recording starts over.  No other
        // block touches this shard's photons at this point (streamed loop: stream order; persistent loop: the pass
        // blocks wait on the generation word).
        if (n_dt == MAX_DT) {
            for (int j = threadIdx.x; j < st.count; j += EVT_THREADS) {
                const int i = st.first + j;
                if (d.ph.flags[i] & F_MOVABLE) {
                    double r0 = d.ph.r0[i], r1 = d.ph.r1[i], r2 = d.ph.r2[i];
                    apply_pushes_v(st, MAX_DT, d.ph.v0[i], d.ph.v1[i], d.ph.v2[i], r0, r1, r2);
                    d.ph.r0[i] = r0;
                    d.ph.r1[i] = r1;
                    d.ph.r2[i] = r2;
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                st.n_dt = MAX_DT;
                fold_path(st, d.path_pad);
                st.n_dt = 0;
                n_dt = 0;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const int i = sh_cand_i;
            const double t = sh_cand_t;
            bool event = false, attempt = false;
            ph_index = i;
            scatt_time = t;
            if (t < dt_max) {
                st.dt_list[n_dt] = t - old_scatt_time;
                n_dt++;
                attempt = true;
            } else {
                scatt_time = dt_max;
                st.dt_list[n_dt] = scatt_time - old_scatt_time;
                n_dt++;
                event = true;
            }
            old_scatt_time = scatt_time;
            sh_try = attempt ? 1 : 0;
            sh_event = event ? 1 : 0;
            // if this candidate is accepted, these are the event's final numbers
            early.enabled = (early_gst != nullptr && attempt && !d.cs && step_mode == 0 && !d.replay) ? 1 : 0;
            early.n_dt = n_dt;
            early.ph_index = i;
            early.scatt_time = scatt_time;
        }
        __syncthreads();
        if (sh_try) {
            // three warps: scattering lane | Stokes chain (idle when STOKES_SWITCH is OFF) | helper
            if (threadIdx.x < SCATTER_THREADS)
                scatter_candidate_3w(d, st, rng_sh, mail, sh_cand_i, sh_cand_idx, sh_cand_known ? sh_cand_temp : -1.0, n_dt,
                                     &sh_event, early);
            __syncthreads();
        }
        if (sh_event) break;
        // Klein-Nishina rejection (rare): next entry of this shard's time order after (cand_t, cand_i)
        {
            const double pt = sh_cand_t;
            const int pi = sh_cand_i;
            double nt = DBL_MAX;
            int ni = INT_MAX;
            // streamed loop: the pass blocks' minima are still there.  A block whose minimum comes after (pt, pi) offers
            // exactly that minimum; a block whose minimum has been consumed (at most one per rejection) is read again;
            // photons re-located in this iteration are not in any minimum and come from the re-location list.  Same
            // result as reading every time of the shard, without the 80 MB read by one block at 10^7 photons.
            const bool two_level = (early_gst == nullptr) && !have_pre && step_mode == 0 && !d.replay && nb_per_shard > 1 &&
                                   reloc_list_cheaper(R, st.count);
            if (two_level) {
                for (int k = threadIdx.x; k < nb_per_shard; k += EVT_THREADS) {
                    const int q = s * nb_per_shard + k;
                    const double t = d.bm_t[q];
                    const int ti = d.bm_i[q];
                    if (lex_less(pt, pi, t, ti) && lex_less(t, ti, nt, ni)) {
                        nt = t;
                        ni = ti;
                    }
                }
                for (int k = 0; k < nb_per_shard; ++k) { // uniform over the block
                    const int q = s * nb_per_shard + k;
                    if (lex_less(pt, pi, d.bm_t[q], d.bm_i[q])) continue;
                    // pass block k's photons: j = k * PASS_THREADS + u + m * nb_per_shard * PASS_THREADS (pass_body)
                    for (int base = k * PASS_THREADS; base < st.count; base += nb_per_shard * PASS_THREADS)
                        for (int u = threadIdx.x; u < PASS_THREADS; u += EVT_THREADS) {
                            const int j = base + u;
                            if (j >= st.count) break;
                            const int kk = st.first + j;
                            const double t = d.ph.tts[kk];
                            if (lex_less(pt, pi, t, kk) && lex_less(t, kk, nt, ni)) {
                                nt = t;
                                ni = kk;
                            }
                        }
                }
                for (int j = threadIdx.x; j < R; j += EVT_THREADS) {
                    const int kk = d.reloc_slot[reloc_base + j];
                    if (kk >= st.first && kk < st.first + st.count) {
                        const double t = d.ph.tts[kk];
                        if (lex_less(pt, pi, t, kk) && lex_less(t, kk, nt, ni)) {
                            nt = t;
                            ni = kk;
                        }
                    }
                }
            } else {
                for (int j = threadIdx.x; j < st.count; j += EVT_THREADS) {
                    const int k = st.first + j;
                    double t = d.ph.tts[k];
                    if (lex_less(pt, pi, t, k) && lex_less(t, k, nt, ni)) {
                        nt = t;
                        ni = k;
                    }
                }
            }
            block_argmin<EVT_THREADS>(nt, ni);
            if (threadIdx.x == 0) {
                if (ni == INT_MAX) { // list exhausted (Src/mclib.c:1128 loop bound)
                    sh_finished = 1;
                } else {
                    sh_cand_t = nt;
                    sh_cand_i = ni;
                    sh_cand_idx = d.ph.idx[ni];
                    sh_cand_known = 0;
                }
            }
            __syncthreads();
            if (sh_finished) break;
        }
    }

    // ---- cyclo-synchrotron pool replacement, Src/mcrat.c:791-808 (the list is one shard here) ----
    __shared__ int cs_need, cs_slot;
    if (d.cs) {
        if (threadIdx.x == 0) cs_need = (step_mode == 0 && d.ph.type[ph_index] == 'p') ? 1 : 0;
        __syncthreads();
    } else if (threadIdx.x == 0) {
        cs_need = 0;
    }
    if (d.cs && cs_need) {
        // first null slot of the list (Src/photons.c:143-150)
        int first_null = INT_MAX;
        for (int j = threadIdx.x; j < st.count; j += EVT_THREADS)
            if (d.ph.type[st.first + j] == 'N') {
                first_null = st.first + j;
                break;
            }
        double dummy = 0;
        block_argmin<EVT_THREADS>(dummy, first_null);
        if (threadIdx.x == 0) cs_slot = first_null;
        __syncthreads();
    }

    if (threadIdx.x == 0 && !early.released) {
        event_bookkeeping(st, n_dt, ph_index, scatt_time, step_mode);
        if (step_mode == 0) {
            if (cs_need) {
                gs.cs_comptonized_w += d.ph.weight[ph_index];
                d.ph.type[ph_index] = 'k'; // COMPTONIZED_PHOTON
                if (d.ph.weight[ph_index] != 0) d.ph.flags[ph_index] |= F_MOVABLE;
                if (cs_slot == INT_MAX) {
                    // no null slot: the host must grow the list and emit (Src/photons.c:117-129)
                    st.pause_cs = 1;
                } else {
                    EventRng rng = rng_sh;
                    cs_emit_single(d, rng, ph_index, cs_slot);
                    rng_sh = rng;
                    gs.cs_emitted += 1;
                    gs.cs_scatt_num += 1;
                }
            }
            // Src/mcrat.c:810-831: every 1000 scatterings the driver may have to rebin on the host
            if (d.cs && !st.pause_cs && (st.scatt_cnt % 1000 == 0) && (st.scatt_cnt != 0) && gs.cs_scatt_num > gs.cs_max_photons)
                st.pause_cs = 2;
            event_count_stopped(gs, st);
        }
        if (d.replay) {
            gs.replay_cursor = rng_sh.pos;
            if (rng_sh.exhausted) raise_error(&gs, MCRAT_B200_ERR_REPLAY, ph_index, ERR_SITE_EVENT_REPLAY);
        }
        st.last_event_draw = rng_sh.draw;
    }
    __syncthreads();
    if (mini_reloc_out && threadIdx.x == 0) *mini_reloc_out = early.mini_reloc;
    return early.released != 0;
}

template <int EVT_THREADS>
__global__ void __launch_bounds__(EVT_THREADS) event_kernel(DevCtx d, int parity, int nb_per_shard, int step_mode,
                                                            double dt_max_arg)
{
    const int s = blockIdx.x;
    if (loop_stopped(*d.gs, d.sh[s])) return;
    event_body<EVT_THREADS>(d, s, 0, d.gs->reloc_count[parity], nb_per_shard, step_mode, dt_max_arg, d.sh[s]);
}
