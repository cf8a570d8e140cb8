#!/usr/bin/env python
"""bench.py -- MCRaT hot path on B200: scatterings/s, photon-iterations/s and photon-cell evals/s vs roofline.

The job (default): BASELINE.json configs[4], the north-star configuration -- the 3-D spherical PLUTO-shape jet
(256 x 64 x 64 cells), ONE list of 1e7 photons, decomposed into ``--ranks`` (128) reference ranks: contiguous slot
ranges, each with its own clock, Philox key and time-ordered scatter sequence, no exchange inside the frame loop
(the reference's MPI decomposition, Src/mcrat.c:139-164, 457-479).  ``--gpus N`` splits THAT job over N GPUs
(mcrat_b200.shard.rank_slice over the ranks; GPU g owns ranks [a, b) and their photons): strong scaling, the same
128 ranks and -- bit for bit -- the same photons whatever N is.  ``--weak`` gives every GPU its own copy of the job.

A *step* is one scatter-frame slice of the driver loop (Src/mcrat.c:754-851): a new hydro frame arrives, so every
photon is re-located by the full photon x cell containment scan (find_nearest_grid_switch = 1, Src/mcrat.c:756; K1),
then ``--iters`` iterations of the while-loop run in every rank (fused push / re-check / free-path draw / arg-min
pass over the rank's photons + one scattering event).  Steps continue the same simulation.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference [...]                         the reference's own CPU code (oracle/_ref)
  python bench.py --workload C2                                  BASELINE configs[1] (round-1 default)
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scatterings/sec"
UNIT = "scatterings/s"
NUM_SMS_FP64_LANES = 64  # FP64 lanes per SM: 2 warp-instructions per clock
BYTES_PER_PHOTON_ITERATION = 100.0

_T0 = time.time()


def log(msg):
    print("[bench %6.1fs] %s" % (time.time() - _T0, msg), file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--iters", type=int, default=5000, help="while-loop iterations per rank and step (frame slice)")
    ap.add_argument("--workload", default="C5", choices=["C2", "C5"],
                    help="C5 (default) = 3-D spherical PLUTO-shape jet of the photon-count sweep, the north-star config; "
                         "C2 = BASELINE configs[1], 2-D cylindrical FLASH-shape jet")
    ap.add_argument("--photons", type=float, default=0, help="photons of the whole job (default: 1e7 for C5, 1e5 for C2)")
    ap.add_argument("--scale", type=float, default=1.0, help="grid scale (1.0 = the BASELINE grid, 1 048 576 cells)")
    ap.add_argument("--ranks", type=int, default=0,
                    help="reference ranks the job is decomposed into (default 128 for C5, 16 for C2); both arms use the "
                         "same decomposition, and with --gpus N every GPU owns ranks/N of them")
    ap.add_argument("--weak", action="store_true", help="every GPU runs its own copy of the job (weak scaling)")
    ap.add_argument("--index", action="store_true",
                    help="re-locate a new hydro frame through the bounding-box index (K1c: same cells, ~2000x fewer "
                         "tests) instead of the reference's full photon x cell scan (K1, the default and the kernel "
                         "the FP64 roofline is quoted on)")
    ap.add_argument("--loop", default="auto", choices=["auto", "streamed", "persistent", "persistent_stream", "streamed_global"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the s_sweep block (scatterings/s against the number of ranks)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-scan-photons", type=int, default=0, help="photons per CPU rank sample of the rescan (0: auto)")
    ap.add_argument("--cpu-iters", type=int, default=0, help="loop iterations per CPU rank sample (0: auto)")
    ap.add_argument("--cpu-cores", type=int, default=0, help="CPU arm: host cores to use (0: all the process may run on)")
    ap.add_argument("--comm-gather", action="store_true",
                    help="with --gpus N > 1: also time one gather of the whole job's photons to rank 0 through the product's communicator "
                         "(always done at N = 1; profiles/bench_r02_2gpu.json has it for two GPUs)")
    ap.add_argument("--cpu-build", default="o3", choices=["o3", "o2"],
                    help="CPU arm: o3 = the reference's sources built -O3 -march=x86-64-v3 (default); o2 = the -O2 -ffp-contract=off "
                         "build the parity oracle uses (-O2 is what the reference's authors advise)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        # "under load": samples taken while the GPU drew clearly more than idle power
        load = [s for s, p in zip(sm, pw) if p > 300.0] if len(pw) == len(sm) else []
        return {"sm_mhz": float(np.median(load if load else sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(load),
                "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------------
# the job and its decomposition
# ------------------------------------------------------------------------------------------------
def job_of(args):
    """-> dict describing the whole job (independent of --gpus unless --weak)."""
    photons = int(args.photons) if args.photons else (10_000_000 if args.workload == "C5" else 100_000)
    ranks = args.ranks if args.ranks else (128 if args.workload == "C5" else 16)
    ranks = max(1, min(ranks, photons // 256 if photons >= 256 else 1))
    size = -(-photons // ranks)          # photons per rank (mcrat_b200.shard.sub_shard_ranges)
    ranks = -(-photons // size)
    return dict(workload=args.workload, photons=photons, ranks=ranks, rank_size=size)


def my_share(job, rank, world, weak):
    """Ranks [a, b) and photon slots [lo, hi) of GPU `rank`."""
    from mcrat_b200 import shard as shardlib
    if weak or world == 1:
        return 0, job["ranks"], 0, job["photons"]
    sl = shardlib.rank_slice(job["ranks"], rank, world)
    a, b = sl.start, sl.stop
    return a, b, a * job["rank_size"], min(b * job["rank_size"], job["photons"])


def workload_name(job, args):
    if job["workload"] == "C5":
        base = "C5: 3-D spherical PLUTO-shape jet, 256x64x64 cells, %d photons in one list, Stokes on" % job["photons"]
    else:
        base = "C2: 2-D cylindrical FLASH-shape GRB jet, 1024x1024 cells, %d photons, Stokes on" % job["photons"]
    if args.scale != 1.0:
        base += " (REDUCED grid: scale=%g)" % args.scale
    return base


class stdout_to_stderr:
    """NCCL prints its version banner on stdout when a communicator is created; stdout is reserved for the one JSON
    line, so file descriptor 1 points at stderr while communicators come up."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own sources (oracle/_ref), bounded sample + extrapolation
# ------------------------------------------------------------------------------------------------
def _cpu_core(core, ncores, job, cfg, hydro, photons, cells, frame, n_scan, k_iters, nsamples, refname, timing, kind,
              barrier, out):
    """One host core.  Its share of the job's ranks is core, core + ncores, ...; per sample it runs ONE of them:
    (a) the rescan (findContainingHydroCell, switch = 1) of n_scan of that rank's photons against the whole grid,
    (b) k_iters iterations of the frame loop over the rank's whole list (switch = 0, photons start located)."""
    from oracle import api
    mine = list(range(core, job["ranks"], ncores))
    if kind == "reference":
        eng = api.RefLib(refname, timing=timing)
    else:
        eng = api.Oracle(cfg)
    eng.set_hydro(hydro)
    res = []
    for k in range(nsamples):
        r = mine[k % len(mine)] if mine else None
        barrier.wait()
        if r is None:
            res.append(None)
            continue
        lo, hi = r * job["rank_size"], min((r + 1) * job["rank_size"], job["photons"])
        lst = photons[lo:hi].copy()
        rng = eng.new_rng(seed=r + 1)[0] if kind == "reference" else api.OracleRng("ranlxs0", seed=r + 1)
        # (a) rescan sample: photons spread evenly over the rank's list
        pick = np.unique(np.linspace(0, lst.size - 1, min(n_scan, lst.size)).astype(np.int64))
        eng.set_photons(lst[pick])
        t0 = time.perf_counter()
        eng.find_containing_hydro_cell(1, rng)
        t_scan = (time.perf_counter() - t0) / pick.size
        # (b) loop sample over the whole rank, photons located by the grid's own structure
        lst["nearest_block_index"] = cells[lo:hi]
        eng.set_photons(lst)
        t0 = time.perf_counter()
        st = eng.run_frame(rng, frame["time_now"], 1.0 / frame["fps"], max_iters=k_iters, switch=0)
        t_loop = time.perf_counter() - t0
        it = max(int(st["iterations"]), 1)
        res.append((t_scan, t_loop / it, st["scatterings"] / it, hi - lo, len(mine), t_scan * pick.size + t_loop))
    out.put((core, res))


def run_cpu_arm(job, cfg, hydro, photons, frame, iters, nsamples, refname, n_scan=0, k_iters=0, build="o3", cores=0):
    """The reference's own CPU code on all host cores, one rank per core at a time (the reference's own way of using
    more cores).  Returns one entry per sample with the step time of the WHOLE job extrapolated from it:
        t_step(core) = sum over the core's ranks of [ photons(rank) x t_scan_per_photon + iters x t_iteration ]
        t_step(job)  = max over cores;   scatterings(job) = ranks x iters x (scatterings per iteration, measured)."""
    from oracle import api
    from mcrat_b200 import synth
    timing = build == "o3" and api.ref_available(refname, timing=True) and api.host_runs_timing_build()
    kind = "reference" if api.ref_available(refname, timing=timing) else "port"
    if kind == "port":
        api.build_oracle()
    else:
        api.RefLib(refname, timing=timing)  # dlopen in the parent too: the forked workers inherit the mapping
    ncores = len(os.sched_getaffinity(0))
    if cores > 0:
        ncores = min(ncores, cores)
    ncores = max(1, min(ncores, job["ranks"]))
    ncells = int(hydro["num_elements"])
    if not k_iters:  # ~1 s of loop per sample: one iteration costs ~0.3 us per photon of the rank (draw, push, qsort)
        k_iters = int(max(8, min(iters, 1.0 / (3e-7 * job["rank_size"]))))
    if not n_scan:   # the step's own mix: rank_size photons rescanned per `iters` iterations, scaled to k_iters
        n_scan = int(max(16, min(job["rank_size"], round(job["rank_size"] * k_iters / max(iters, 1)))))
    log("cpu arm: %d cores, %s build%s, %d sample(s): rescan of %d photons + %d loop iterations of a %d-photon rank"
        % (ncores, kind, " (-O3 -march=x86-64-v3)" if timing else " (-O2)", nsamples, n_scan, k_iters, job["rank_size"]))
    cells = synth.locate_photons(hydro, photons)
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(ncores)
    out = ctx.Queue()
    procs = [ctx.Process(target=_cpu_core, args=(c, ncores, job, cfg, hydro, photons, cells, frame, n_scan, k_iters,
                                                 nsamples, refname, timing, kind, barrier, out)) for c in range(ncores)]
    for p in procs:
        p.start()
    res = dict(out.get() for _ in procs)
    for p in procs:
        p.join()
    samples = []
    for k in range(nsamples):
        t_step, t_sample, spi, t_scan_all, t_iter_all = 0.0, 0.0, [], [], []
        for c in range(ncores):
            e = res[c][k]
            if e is None:
                continue
            t_scan, t_iter, s_per_it, nph, nmine, t_meas = e
            t_step = max(t_step, nmine * (nph * t_scan + iters * t_iter))
            t_sample = max(t_sample, t_meas)
            spi.append(s_per_it)
            t_scan_all.append(t_scan)
            t_iter_all.append(t_iter)
        scatt = job["ranks"] * iters * float(np.mean(spi))
        samples.append(dict(step_s=t_step, sample_s=t_sample, scatterings=scatt, value=scatt / t_step,
                            scan_us_per_photon=1e6 * float(np.mean(t_scan_all)), loop_ms_per_iteration=1e3 * float(np.mean(t_iter_all)),
                            photon_iterations_per_s=job["photons"] * iters / t_step))
    flags = "-O3 -march=x86-64-v3 -fopenmp" if timing else "-O2 -fopenmp -ffp-contract=off"
    return dict(samples=samples, cores=ncores, kind=kind, nproc=os.cpu_count(), cpu_model=cpu_model(), flags=flags,
                sample="per step and core: one of the core's ranks runs a scaled-down step with the job's own mix -- rescan "
                       "(findContainingHydroCell, switch=1) of %d of its photons against all %d cells + %d loop iterations over "
                       "its whole %d-photon list (full step: %d-photon rescan + %d iterations per rank, %d ranks, %d per core); "
                       "value = the job's scatterings per step / its step time, both composed linearly from the two measured "
                       "unit costs (BASELINE.md section 2); ms_per_step is the time the sample itself took"
                       % (n_scan, ncells, k_iters, job["rank_size"], job["rank_size"], iters, job["ranks"], -(-job["ranks"] // ncores)))


def cpu_line_fields(r):
    v = float(np.mean([s["value"] for s in r["samples"]]))
    return {"value": v, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
            "nproc": r["nproc"], "cpu_model": r["cpu_model"], "flags": r["flags"],
            "scan_us_per_photon": float(np.mean([s["scan_us_per_photon"] for s in r["samples"]])),
            "loop_ms_per_iteration": float(np.mean([s["loop_ms_per_iteration"] for s in r["samples"]])),
            "photon_iterations_per_sec": float(np.mean([s["photon_iterations_per_s"] for s in r["samples"]])),
            "extrapolated_step_s": float(np.mean([s["step_s"] for s in r["samples"]])),
            "sample_s": float(np.mean([s["sample_s"] for s in r["samples"]]))}


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0

    from mcrat_b200 import synth
    job = job_of(args)
    log("building the job: %s, %d photons, %d ranks of %d" % (job["workload"], job["photons"], job["ranks"], job["rank_size"]))
    # strong scaling: every GPU builds the same list (same seed) and takes its ranks' slice; weak: its own list
    seed = 1234 + (rank if args.weak else 0)
    cfg, hydro, photons_all, frame = synth.workload(job["workload"], scale=args.scale, n_photons=job["photons"], seed=seed)
    refname = {"C2": "c2_2d_cyl_stokes", "C5": "c5_3d_sph"}[job["workload"]]
    ra, rb, lo, hi = my_share(job, rank, world, args.weak)
    scaling = "weak" if args.weak else "strong"
    config = {"workload": workload_name(job, args), "cells": int(hydro["num_elements"]), "photons_total": job["photons"] * (world if args.weak else 1),
              "reference_ranks": job["ranks"] * (world if args.weak else 1), "photons_per_rank": job["rank_size"],
              "loop_iterations_per_rank_and_step": args.iters, "gpus": world,
              "decomposition": "one list split into %d reference ranks (contiguous slot ranges; own clock, Philox key and "
                               "time-ordered scatter sequence each; no exchange inside the frame loop); %s"
                               % (job["ranks"], "every GPU runs its own copy of the job" if args.weak else
                                  "GPU g of N owns ranks rank_slice(%d, g, N) and their photons -- same job, same photons for every N" % job["ranks"]),
              "step": "full photon x cell rescan of the GPU's photons (new hydro frame) + loop iterations in every rank; steps continue one simulation",
              "l2": "photon list and cell arrays exceed L2 at N <= 4 (1e7 x 100 B per pass); additionally flushed between timed steps (256 MiB write)",
              "loop": args.loop + " (auto: from 1.2e6 photons and 16 ranks of <= 2e5 photons per GPU on, a pair of co-resident grids per frame "
                      "(resident event blocks beside pass blocks streaming the photon columns); otherwise one cooperative launch per "
                      "frame up to 2^21 photons and two launches per iteration and half of the ranks above)",
              "relocation": "bounding-box index over the cells in array order (K1c, identical first-match results)"
              if args.index else "full photon x cell scan of every new hydro frame (K1), as the reference does; "
              "steady-state re-locations inside the loop go through the bounding-box index"}

    if args.impl == "reference":
        r = run_cpu_arm(job, cfg, hydro, photons_all, frame, args.iters, args.warmup + args.steps, refname,
                        args.cpu_scan_photons, args.cpu_iters, build=args.cpu_build, cores=args.cpu_cores)
        r["samples"] = r["samples"][args.warmup:]
        f = cpu_line_fields(r)
        line = {"impl": "reference", "metric": METRIC, "value": f["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * f["sample_s"],
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "photon_iterations_per_sec": f["photon_iterations_per_sec"],
                "cpu_baseline": f,
                "e2e": {"value": f["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return 0

    # ---------------- our arm ----------------
    import torch
    import torch.distributed as dist
    from mcrat_b200 import HotPath
    from mcrat_b200.lib import HYDRO_FIELDS, PHOTON_DTYPE

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.all_reduce(torch.zeros(1, device="cuda"))
            torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    photons = photons_all[lo:hi]
    nshards = rb - ra
    shard_base = ra + (rank * job["ranks"] if args.weak else 0)
    stream = torch.cuda.current_stream().cuda_stream
    hp = HotPath(cfg, device=local_rank, seed=20261018, shard=shard_base, stream=stream, num_shards=nshards,
                 scan_index=args.index, loop_mode=args.loop)
    hp.set_hydro(hydro)
    hp.set_photons(photons)
    # product-side communicator (include/mcrat_b200.h, mcrat_b200_comm_*): per-frame counter reduction inside the timed
    # steps, photon-count gather and the photon gather for the merged output measured once below
    from mcrat_b200 import Comm
    comm, comm_err = None, None
    try:
        with stdout_to_stderr():
            comm = Comm.from_torch_dist(hp, dist if world > 1 else None)
            from mcrat_b200.lib import FrameStats
            comm.reduce_frame_stats(FrameStats().as_dict())  # first collective: the channels come up outside the timed steps
    except Exception as exc:
        if world > 1:
            raise
        comm_err = str(exc)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    dt_frame = 1.0 / frame["fps"]
    config["photons_this_gpu"] = int(photons.size)
    config["ranks_this_gpu"] = int(nshards)

    log("context ready (%d photons, %d ranks on this GPU); warm-up" % (photons.size, nshards))
    # ---- device-resident arm: `value` ----
    time_now = frame["time_now"]
    for _ in range(args.warmup):
        st = hp.run_frame(time_now, dt_frame, max_iters=args.iters, switch=1)
        time_now = st["time_now"]
    log("timed steps")
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = hp.launch_count()
    tot_ms = 0.0
    scatt = evals = slots = ref_evals = job_scatt = 0
    barrier()
    for _ in range(args.steps):
        flush_buf.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        st = hp.run_frame(time_now, dt_frame, max_iters=args.iters, switch=1)
        if comm is not None:
            job_st = comm.reduce_frame_stats(st)   # the frame's counters over all GPUs (NCCL, on the context's stream)
        else:
            job_st = st
        e1.record()
        torch.cuda.synchronize()
        tot_ms += e0.elapsed_time(e1)
        time_now = st["time_now"]
        scatt += st["scatterings"]
        evals += st["cell_evals"]
        slots += st["photon_slots"]
        ref_evals += st["ref_equiv_evals"]
        job_scatt += job_st["scatterings"]
    barrier()
    launches = hp.launch_count() - launches0

    log("timed steps done: %.1f ms/step; loop alone" % (tot_ms / args.steps))
    # ---- the loop alone (no rescan), CUDA events around one call ----
    probe_iters = max(64, min(args.iters, 2000))
    flush_buf.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    st = hp.run_frame(time_now, dt_frame, max_iters=probe_iters, switch=0)
    e1.record()
    torch.cuda.synchronize()
    time_now = st["time_now"]
    loop_us = 1e3 * e0.elapsed_time(e1) / max(st["iterations"], 1)
    loop_slots_per_s = st["photon_slots"] / (e0.elapsed_time(e1) * 1e-3)
    loop_scatt_per_s = st["scatterings"] / (e0.elapsed_time(e1) * 1e-3)
    clk = clocks.stop() if rank == 0 else None

    # ---- per-kernel times of the streamed loop (profile mode: every launch bracketed by CUDA events) ----
    hp.set_profile(True)
    for _ in range(2):  # the first round loads the streamed loop's kernels (the timed steps ran another driver); the second counts
        hp.kernel_times(reset=True)
        st = hp.run_frame(time_now, dt_frame, max_iters=24, switch=0)
        time_now = st["time_now"]
        kt = hp.kernel_times(reset=True)
    hp.set_profile(False)
    pass_ms = kt["pass_ms"] / max(kt["pass_launches"], 1)
    event_ms = kt["event_ms"] / max(kt["event_launches"], 1)
    # photons one pass launch covers: the interleaved streamed loop launches the pass per half of the ranks
    pass_photons = st["photon_slots"] / max(kt["pass_launches"], 1)

    log("loop %.1f us/iteration (pass %.1f us, event %.1f us in profile mode); scan roofline" % (loop_us, 1e3 * pass_ms, 1e3 * event_ms))
    # ---- roofline of the kernel that owns the step (K1 scan), timed alone with CUDA events on its stream ----
    scan_ms, scan_evals = [], 0
    for _ in range(2 if photons.size > 2_000_000 else 3):
        flush_buf.fill_(1)
        ev, ms = hp.rescan_all()
        scan_ms.append(ms)
        scan_evals = ev
    scan_ms_avg = float(np.mean(scan_ms))
    dfma_peak = hp.measure_fp64_peak()  # G FP64-pipe instr/s (16 independent DFMA chains per thread), same GPU, same run
    props = torch.cuda.get_device_properties(local_rank)
    num_sms = props.multi_processor_count
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    sm_mhz = (clk or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    hw_peak = num_sms * NUM_SMS_FP64_LANES * sm_mhz * 1e6 / 1e9  # G thread-instr/s: 2 warp-instr / clk / SM
    ndim3 = cfg["dimensions"] == synth.THREE
    instr_per_eval = 6 if ndim3 else 4  # one DADD + one DSETP per dimension
    achieved = scan_evals * instr_per_eval / (scan_ms_avg * 1e-3) / 1e9
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    tkey = "scan_kernel_%s_%d_bytes" % (job["workload"], photons.size)
    traffic = json.load(open(tpath)).get(tkey) if os.path.exists(tpath) and args.scale == 1.0 else None
    alg_bytes = (48 if ndim3 else 32) * int(hydro["num_elements"]) + 28 * photons.size
    k1_share = scan_ms_avg / (tot_ms / args.steps)
    roofline = {"kernel": "scan_kernel<%d> (K1 photon x cell containment scan, %s)" % (1 if ndim3 else 0, "3-D" if ndim3 else "2-D"),
                "bound": "fp64", "achieved": achieved, "peak": hw_peak, "unit": "G FP64-pipe instr/s",
                "frac": achieved / hw_peak, "traffic": traffic,
                "traffic_note": "DRAM bytes per launch (ncu --set full, profiles/ncu_traffic.json key %s); algorithmic minimum "
                                "%d B x cells + 28 B x photons = %d" % (tkey, 48 if ndim3 else 32, alg_bytes),
                "peak_source": "hardware issue peak: %d SMs x 64 FP64 lanes (2 warp-instr/clk) x %.0f MHz (median SM clock "
                               "under load in this run, nvidia-smi); MEASURED_PEAKS.json holds no FP64 figure" % (num_sms, sm_mhz),
                "dfma_microbenchmark": dfma_peak, "frac_of_dfma_microbenchmark": achieved / dfma_peak,
                "algorithmic": "%d FP64-pipe instr per photon-cell eval x %d evals per launch" % (instr_per_eval, scan_evals),
                "evals_per_s": scan_evals / (scan_ms_avg * 1e-3), "ms_per_launch": scan_ms_avg, "share_of_step": k1_share}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    pass_gbs = pass_photons * BYTES_PER_PHOTON_ITERATION / (pass_ms * 1e-3) / 1e9 if pass_ms > 0 else None
    pass_roofline = {"kernel": "pass kernel (K4+K2: push + re-check + free-path draw + block arg-min), %d photons per launch "
                               "(%d launches per iteration of the %d-photon list)" % (pass_photons, round(photons.size / max(pass_photons, 1)), photons.size),
                     "bound": "hbm", "achieved": pass_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": pass_gbs / hbm_peak if pass_gbs else None,
                     "traffic": (json.load(open(tpath)).get("pass_local_kernel_C5_%d_bytes" % round(pass_photons)) if os.path.exists(tpath) else None),
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s",
                     "algorithmic": "100 B per photon-iteration (SURVEY 8d) x %d photons per launch; columns actually moved: 97 B" % pass_photons,
                     "ms_per_launch": pass_ms, "event_kernel_ms_per_launch": event_ms,
                     "note": "meaningful while the list exceeds L2 (> 2^21 photons per GPU); below that the pass runs out of L2"}
    loop_roofline = {"kernel": "whole loop iteration (pass over %d photons + one scattering event per rank, as the loop driver in use overlaps them)" % photons.size,
                     "bound": "hbm", "achieved": loop_slots_per_s * BYTES_PER_PHOTON_ITERATION / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": loop_slots_per_s * BYTES_PER_PHOTON_ITERATION / 1e9 / hbm_peak, "us_per_iteration": loop_us}

    # ---- s_sweep: scatterings/s of the loop against the number of ranks the same list is split into ----
    s_sweep = None
    if not args.no_sweep and rank == 0:
        log("s_sweep")
        s_sweep = []
        try:
            hs = HotPath(cfg, device=local_rank, seed=7, shard=0, stream=stream, scan_index=True, loop_mode=args.loop)
            hs.set_hydro(hydro)
            for S in (1, 16, 64, 296):
                if S > max(1, photons.size // 256):
                    continue
                hs.set_num_shards(S)
                hs.set_photons(photons)
                w = hs.run_frame(frame["time_now"], dt_frame, max_iters=8, switch=1)  # locate through the index, warm up
                n_it = 100 if photons.size > 2_000_000 else 1000
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                s2 = hs.run_frame(w["time_now"], dt_frame, max_iters=n_it, switch=0)
                e1.record()
                torch.cuda.synchronize()
                sec = e0.elapsed_time(e1) * 1e-3
                s_sweep.append({"ranks_on_this_gpu": hs.num_shards(), "scatterings_per_sec": s2["scatterings"] / sec,
                                "photon_iterations_per_sec": s2["photon_slots"] / sec,
                                "us_per_iteration": 1e6 * sec / max(s2["iterations"], 1)})
            hs.close()
        except Exception as exc:  # measurement extra; never fail the bench line over it
            s_sweep.append({"error": str(exc)})

    # ---- e2e: the same step through the C ABI with host buffers (H2D + D2H inside the timed region) ----
    e2e = None
    e2e_s, e2e_scatt, h2d, d2h = 0.0, 0, 0, 0
    if not args.no_e2e:
        log("e2e")
        host_ph = torch.empty(photons.size * PHOTON_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
        host_np = host_ph.numpy().view(PHOTON_DTYPE)
        host_np[:] = hp.get_photons()
        h2d = sum(np.asarray(hydro[f]).nbytes for f in HYDRO_FIELDS) + host_ph.numel()
        d2h = host_ph.numel()
        # the hydro frame sits in pinned host memory, as a reader that fills the device-bound arrays directly would leave it
        hydro_pinned = dict(hydro)
        pinned_keep = []
        for f in HYDRO_FIELDS:
            if f in hydro:
                t = torch.from_numpy(np.ascontiguousarray(hydro[f], dtype=np.float64)).pin_memory()
                pinned_keep.append(t)
                hydro_pinned[f] = t.numpy()
        for k in range(1 + args.steps):
            if k == 1:
                barrier()
                t0 = time.perf_counter()
            hp.set_hydro(hydro_pinned)                            # the frame the driver just read (Src/mcrat.c:721)
            hp.set_photons_ptr(host_ph.data_ptr(), photons.size)  # host list -> device
            st = hp.run_frame(time_now, dt_frame, max_iters=args.iters, switch=1)
            hp.get_photons_ptr(host_ph.data_ptr(), photons.size)  # device -> host list (checkpoint / mc_proc output)
            time_now = st["time_now"]
            if k >= 1:
                e2e_scatt += st["scatterings"]
        barrier()
        e2e_s = time.perf_counter() - t0

    # ---- the exchanges either side of the frame loop, measured once: photon counts, photon gather to rank 0 ----
    comm_info = {"error": comm_err} if comm is None else None
    if comm is not None:
        log("comm: counts + gather")
        ver = hp.L.mcrat_b200_comm_nccl_version()
        barrier()
        t0 = time.perf_counter()
        cnt = comm.photon_counts()
        t_counts = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(20):
            comm.reduce_frame_stats(st)
        t_reduce = (time.perf_counter() - t0) / 20
        gather_ms, gathered = None, None
        total_records = int(cnt["list_capacity"].sum())
        if not args.no_e2e and (world == 1 or args.comm_gather):
            out = None
            if rank == 0:
                out_t = torch.empty(total_records * PHOTON_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
                out = out_t.numpy().view(PHOTON_DTYPE)
            barrier()
            t0 = time.perf_counter()
            allp, counts = comm.gather_photons(root=0, out=out)
            barrier()
            gather_ms = 1e3 * (time.perf_counter() - t0)
            gathered = int(counts.sum())
            if rank == 0:  # rank order = slot order of the one list the job started from
                assert allp.size == total_records
                mine_now = hp.get_photons()
                assert allp[:mine_now.size].tobytes() == mine_now.tobytes(), "gathered list does not start with rank 0's photons"
        comm_info = {"library": "NCCL %d.%d.%d (bound at run time)" % (ver // 10000, (ver // 100) % 100, ver % 100), "ranks": comm.size,
                     "per_step_inside_timed_region": "all-reduce of the frame counters (4 NCCL reductions in one group, 20 words)",
                     "frame_stats_allreduce_us": 1e6 * t_reduce, "photon_counts_allgather_us": 1e6 * t_counts,
                     "photons_per_gpu": [int(x) for x in cnt["list_capacity"]],
                     "output_photons_per_gpu": [int(x) for x in cnt["output_photons"]],
                     "gather_photons_to_rank0_ms": gather_ms, "gather_records": gathered,
                     "gather_bytes": None if gathered is None else gathered * PHOTON_DTYPE.itemsize,
                     "scatterings_per_step_from_the_allreduce": job_scatt / args.steps,
                     "nccl_operations": comm.collectives()}

    # ---- aggregate over ranks ----
    t_max, scatt_all, evals_all, slots_all, e2e_max, e2e_all = tot_ms, scatt, evals, slots, e2e_s, e2e_scatt
    ref_evals_all, loop_slots_all, loop_scatt_all, scan_evals_rate = ref_evals, loop_slots_per_s, loop_scatt_per_s, roofline["evals_per_s"]
    per_gpu_ms = [tot_ms / args.steps]
    if world > 1:
        mine_ms = torch.tensor([tot_ms / args.steps], dtype=torch.float64, device="cuda")
        all_ms = [torch.zeros_like(mine_ms) for _ in range(world)]
        dist.all_gather(all_ms, mine_ms)
        per_gpu_ms = [float(x) for x in all_ms]
        t = torch.tensor([tot_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c = torch.tensor([scatt, evals, slots, e2e_scatt, ref_evals, loop_slots_per_s, loop_scatt_per_s, roofline["evals_per_s"]],
                         dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        t_max, e2e_max = float(t[0]), float(t[1])
        scatt_all, evals_all, slots_all, e2e_all, ref_evals_all, loop_slots_all, loop_scatt_all, scan_evals_rate = [float(x) for x in c]

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            log("cpu baseline")
            r = run_cpu_arm(job, cfg, hydro, photons_all, frame, args.iters, 1, refname,
                            args.cpu_scan_photons or 0, args.cpu_iters or 0, build=args.cpu_build, cores=args.cpu_cores)
            cpu = cpu_line_fields(r)
        line = {"metric": METRIC, "value": scatt_all / (t_max * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_max / args.steps,
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "photon_iterations_per_sec": slots_all / (t_max * 1e-3),
                "photon_cell_evals_per_sec": scan_evals_rate,
                "photon_cell_evals_executed_per_step": evals_all / args.steps,
                "reference_equivalent_evals_per_step": ref_evals_all / args.steps,
                "k1_full_scan_ms": scan_ms_avg,
                "ms_per_step_per_gpu": per_gpu_ms,
                "loop_us_per_iteration": loop_us,
                "loop_only": {"scatterings_per_sec": loop_scatt_all, "photon_iterations_per_sec": loop_slots_all,
                              "note": "frame loop without the per-step rescan, %d iterations, summed over GPUs" % probe_iters},
                "roofline": roofline, "pass_roofline": pass_roofline, "loop_roofline": loop_roofline,
                "s_sweep": s_sweep, "comm": comm_info, "cpu_baseline": cpu,
                "e2e": None if args.no_e2e else {"value": e2e_all / e2e_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                                                 "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_max / args.steps},
                "gpu_launches": int(launches), "clocks": clk}
        print(json.dumps(line), flush=True)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
