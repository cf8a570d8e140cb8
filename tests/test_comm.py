"""Multi-GPU plumbing of the product (include/mcrat_b200.h, mcrat_b200_comm_*): the reference's MPI exchanges either side
of the frame loop -- table broadcast (Src/hot_x_section.c:717), per-frame counters, photon gather for the merged output
(Src/merge.c:784-876) -- over NCCL.  CPU: argument checks and the point partition; GPU: a one-rank communicator in
process, and two ranks on two GPUs when the box has them."""
import ctypes as C
import os

import numpy as np
import pytest

from mcrat_b200 import lib, shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NPTS = 221 * 81


def test_comm_entry_points_reject_bad_arguments_without_a_gpu():
    L = lib.load()
    small = C.create_string_buffer(16)
    assert L.mcrat_b200_comm_unique_id(small, C.c_size_t(16)) == -2          # ERR_ARG: the id needs 128 bytes
    assert L.mcrat_b200_comm_unique_id(None, C.c_size_t(128)) == -2
    out = C.c_void_p()
    assert L.mcrat_b200_comm_create(None, 2, 0, b"\0" * 128, C.c_size_t(128), C.byref(out)) == -2  # no context
    assert L.mcrat_b200_comm_rank(None) == -1 and L.mcrat_b200_comm_size(None) == 0
    assert L.mcrat_b200_comm_collectives(None) == 0
    assert L.mcrat_b200_comm_bcast_thermal_table(None, 0) == -2
    assert L.mcrat_b200_comm_photon_counts(None, None, None, None) == -2
    assert L.mcrat_b200_comm_gather_photons(None, 0, None, C.c_longlong(0), None, None) == -2
    L.mcrat_b200_comm_destroy(None)  # no-op


def test_nccl_is_bound_at_run_time():
    # this image ships NCCL: the library finds it without linking against it
    assert lib.load().mcrat_b200_comm_nccl_version() >= 22000
    import subprocess
    needed = subprocess.run(["readelf", "-d", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libnccl" not in needed


def test_table_points_are_partitioned_like_the_device_does():
    # mcrat_b200_comm_build_thermal_table: rank r integrates [r * chunk, (r + 1) * chunk), chunk = ceil(npts / nranks)
    for n in (1, 2, 3, 4, 8, 16):
        chunk = -(-NPTS // n)
        seen = np.zeros(NPTS, dtype=np.int32)
        for r in range(n):
            first = r * chunk
            mine = max(0, min(chunk, NPTS - first))
            seen[first:first + mine] += 1
        assert (seen == 1).all()


class _FakeCollectiveLib:
    """Stands in for libmcrat_b200.so in the wrapper test below: `mcrat_b200_comm_gather_photons` behaves like the C function --
    every rank must enter it (a barrier with a time-out stands for the NCCL calls inside), the ranks agree on whether all
    receivers have room, and either all of them fail with ERR_ARG and the total, or the data moves."""

    def __init__(self, nranks, sizes):
        import threading
        self.n, self.sizes = nranks, sizes
        self.barrier = threading.Barrier(nranks, timeout=5)
        self.room = [1] * nranks
        self.calls = [0] * nranks

    def gather(self, rank, root, ptr, cap, counts, tot):
        total = sum(self.sizes)
        recv = root == -1 or root == rank
        self.room[rank] = 1 if (not recv or (ptr is not None and cap >= total)) else 0
        self.calls[rank] += 1
        self.barrier.wait()                      # the all-gather of the counts
        ok = min(self.room)
        self.barrier.wait()                      # the all-reduce of the flag
        for k in range(self.n):
            counts[k] = self.sizes[k]
        tot.value = total
        return 0 if ok else -2


@pytest.mark.parametrize("root,buffers", [(0, "root_only"), (0, "none"), (-1, "none"), (0, "too_small"), (-1, "mixed")])
def test_gather_wrapper_keeps_the_ranks_in_step(root, buffers):
    """Comm.gather_photons must issue exactly the same sequence of collectives on every rank whatever each rank passes
    (rank 0 with a buffer and the others without is how bench.py calls it).  Four fake ranks on threads; a rank that enters a
    collective the others do not enter trips the barrier's time-out."""
    import threading
    n, sizes = 4, [5, 0, 7, 3]
    fake = _FakeCollectiveLib(n, sizes)
    results, errors = [None] * n, []

    def rank_main(r):
        class L:
            @staticmethod
            def mcrat_b200_comm_gather_photons(c, root_, ptr, cap, counts_ptr, tot_ref):
                counts = np.ctypeslib.as_array(counts_ptr, shape=(n,))
                return fake.gather(r, root_.value, ptr, cap.value, counts, tot_ref._obj)

            @staticmethod
            def mcrat_b200_comm_size(c):
                return n

            @staticmethod
            def mcrat_b200_comm_rank(c):
                return r

            @staticmethod
            def mcrat_b200_comm_photon_counts(*a):
                raise AssertionError("the wrapper must not issue collectives of its own")
        comm = lib.Comm.__new__(lib.Comm)
        comm.L, comm.c, comm.hp = L, None, None
        out = None
        if buffers == "root_only" and r == 0:
            out = np.zeros(sum(sizes), dtype=lib.PHOTON_DTYPE)
        if buffers == "too_small" and r == 0:
            out = np.zeros(2, dtype=lib.PHOTON_DTYPE)
        if buffers == "mixed" and r % 2 == 0:
            out = np.zeros(sum(sizes) + r, dtype=lib.PHOTON_DTYPE)
        try:
            results[r] = comm.gather_photons(root=root, out=out)
        except Exception as exc:  # a broken barrier = ranks out of step
            errors.append((r, repr(exc)))

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=30)
    assert not errors, errors
    assert len(set(fake.calls)) == 1, "ranks made different numbers of collective calls: %s" % fake.calls
    assert fake.calls[0] == (1 if buffers in ("root_only",) else 2)
    for r in range(n):
        allp, counts = results[r]
        assert counts.tolist() == sizes
        if root == -1 or root == r:
            assert allp is not None and allp.size == sum(sizes)
        else:
            assert allp is None


# ------------------------------------------------------------------------------------------------
def _frame(hp, photons, hydro, frame, iters=60):
    hp.set_hydro(hydro)
    hp.set_photons(photons)
    return hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=iters, switch=1)


@pytest.mark.gpu
def test_one_rank_communicator_is_the_identity():
    from mcrat_b200 import Comm, HotPath, comm_unique_id
    cfg, hydro, photons, frame = synth.workload("C2", scale=1.0 / 16, n_photons=3000, seed=3)
    photons["weight"][::7] = 0.0
    photons["type"][::7] = b"N"  # null photons, Src/mcrat.h:57 (setNullPhoton zeroes the weight too, Src/photons.c:230-250)
    hp = HotPath(cfg, device=0, seed=99, shard=0, num_shards=4)
    st = _frame(hp, photons, hydro, frame)
    comm = Comm(hp, 1, 0, comm_unique_id())
    assert comm.rank == 0 and comm.size == 1
    tot = comm.reduce_frame_stats(st)
    for k, v in st.items():
        assert tot[k] == v, k
    got = hp.get_photons()
    cnt = comm.photon_counts()
    assert cnt["list_capacity"].tolist() == [photons.size]
    assert cnt["output_photons"].tolist() == [int((got["weight"] != 0).sum())]
    assert cnt["null_slots"].tolist() == [int((got["type"] == b"N").sum())]
    for root in (0, -1):
        allp, counts = comm.gather_photons(root=root)
        assert counts.tolist() == [photons.size]
        assert allp.tobytes() == got.tobytes()
    # a receive buffer that is too small: the C call moves nothing and says how much is needed ...
    small = np.zeros(10, dtype=lib.PHOTON_DTYPE)
    cnts, tot = np.zeros(1, dtype=np.int64), C.c_longlong(0)
    rc = hp.L.mcrat_b200_comm_gather_photons(comm.c, C.c_int(0), small.ctypes.data_as(C.c_void_p), C.c_longlong(small.size),
                                             cnts.ctypes.data_as(C.POINTER(C.c_longlong)), C.byref(tot))
    assert rc == -2 and tot.value == photons.size and cnts.tolist() == [photons.size] and not small["weight"].any()
    # ... and the wrapper comes back with a buffer of that size
    allp, _ = comm.gather_photons(root=0, out=small)
    assert allp.tobytes() == got.tobytes()
    t1, _ = hp.build_thermal_table(calls=3000, seed=5)
    t2, _ = comm.build_thermal_table(calls=3000, seed=5)
    assert np.array_equal(t1, t2)
    comm.bcast_thermal_table(0)
    assert comm.collectives() > 0
    comm.close()
    hp.close()


def _two_gpu_worker(rank, world, pipe, q):
    import faulthandler
    faulthandler.dump_traceback_later(150, exit=True)  # a rank stuck in a collective must not outlive the test
    import numpy as np
    from mcrat_b200 import Comm, HotPath, comm_unique_id, shard, synth
    cfg, hydro, photons, frame = synth.workload("C3", scale=1.0 / 16, n_photons=4001, seed=3)
    nranks_ref = 8  # reference ranks of the job; GPU g owns rank_slice(8, g, world) of them
    size = -(-photons.size // nranks_ref)
    sl = shard.rank_slice(nranks_ref, rank, world)
    mine = photons[sl.start * size:min(sl.stop * size, photons.size)]
    hp = HotPath(cfg, device=rank, seed=99, shard=sl.start, num_shards=sl.stop - sl.start)
    if rank == 0:
        uid = comm_unique_id()
        pipe.send(uid)
    else:
        uid = pipe.recv()
    comm = Comm(hp, world, rank, uid)
    # the table: built by both GPUs together, then rank 0's copy broadcast (must change nothing)
    table, _ = comm.build_thermal_table(calls=2000, seed=5)
    comm.bcast_thermal_table(0)
    hp.set_hydro(hydro)
    hp.set_photons(mine)
    st = hp.run_frame(frame["time_now"], 1.0 / frame["fps"], max_iters=40, switch=1)
    tot = comm.reduce_frame_stats(st)
    cnt = comm.photon_counts()
    got = hp.get_photons()
    allp, counts = comm.gather_photons(root=0)
    allp2, _ = comm.gather_photons(root=-1)
    q.put((rank, dict(st=st, tot=tot, cnt={k: v.tolist() for k, v in cnt.items()}, got=got.tobytes(),
                      allp=None if allp is None else allp.tobytes(), allp2=allp2.tobytes(), counts=counts.tolist(),
                      table=table.tobytes(), collectives=comm.collectives())))
    comm.close()
    hp.close()


@pytest.mark.gpu
@pytest.mark.timeout(400)
def test_two_gpus_table_counters_and_photon_gather():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from mcrat_b200 import HotPath
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    a, b = ctx.Pipe()
    procs = [ctx.Process(target=_two_gpu_worker, args=(r, 2, (a, b)[r], q), daemon=True) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = dict(q.get(timeout=200) for _ in procs)
        for p in procs:
            p.join(timeout=30)
    finally:
        for p in procs:  # never leave a rank behind (it would hold its GPU and keep pytest from exiting)
            if p.is_alive():
                p.kill()
                p.join(timeout=10)
    r0, r1 = res[0], res[1]
    # counters: sums and maxima of the two ranks' own statistics, the same on both
    for k in ("scatterings", "relocations", "photon_slots", "cell_evals", "box_evals", "ref_equiv_evals", "not_found"):
        assert r0["tot"][k] == r0["st"][k] + r1["st"][k] == r1["tot"][k], k
    assert r0["tot"]["iterations"] == max(r0["st"]["iterations"], r1["st"]["iterations"])
    assert r0["tot"]["time_now"] == max(r0["st"]["time_now"], r1["st"]["time_now"]) == r1["tot"]["time_now"]
    assert r0["cnt"] == r1["cnt"] and sum(r0["cnt"]["list_capacity"]) == 4001
    # gather: rank order, byte for byte what each rank holds
    assert r0["allp"] == r0["got"] + r1["got"] and r1["allp"] is None
    assert r0["allp2"] == r0["allp"] == r1["allp2"]
    assert r0["counts"] == r0["cnt"]["list_capacity"]
    # the table built by two GPUs is the table one GPU builds
    assert r0["table"] == r1["table"]
    cfg = synth.workload("C3", scale=1.0 / 16, n_photons=16, seed=3)[0]
    hp = HotPath(cfg, device=0, seed=1)
    t1, _ = hp.build_thermal_table(calls=2000, seed=5)
    hp.close()
    assert t1.tobytes() == r0["table"]
    assert r0["collectives"] > 0
