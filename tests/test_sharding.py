"""Host-side shard logic, including a world_size-2 gloo run on CPU."""
import os
import sys

import numpy as np
import pytest

from mcrat_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rank_slices_partition_the_list():
    for n in (0, 1, 7, 100000, 10 ** 7 + 3):
        for world in (1, 2, 3, 8):
            sl = [shard.rank_slice(n, r, world) for r in range(world)]
            assert sl[0].start == 0 and sl[-1].stop == n
            assert all(a.stop == b.start for a, b in zip(sl, sl[1:]))
            sizes = [s.stop - s.start for s in sl]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.rank_slice(10, 2, 2)


def test_sub_shard_ranges_match_device_layout():
    for n, s in ((100000, 16), (100000, 1021), (640, 64), (5, 8), (1, 1), (1000, 7)):
        r = shard.sub_shard_ranges(n, s)
        assert r[0][0] == 0 and sum(c for _, c in r) == n
        assert all(c > 0 for _, c in r)
        assert all(a[0] + a[1] == b[0] for a, b in zip(r, r[1:]))
        assert len(r) <= s
    ids = {shard.global_shard_id(rk, 16, k) for rk in range(8) for k in range(16)}
    assert len(ids) == 128


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1001
    sl = shard.rank_slice(n, rank, world)
    stats = dict(iterations=100 + rank, scatterings=10 * (rank + 1), relocations=rank, photon_slots=sl.stop - sl.start,
                 cell_evals=1000 * (rank + 1), not_found=0, time_now=1.0 + rank)
    out = shard.reduce_frame_stats(stats, dist=dist)
    dist.destroy_process_group()
    q.put((rank, out))


def test_reduce_frame_stats_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for r in (0, 1):
        o = res[r]
        assert o["ranks"] == 2
        assert o["scatterings"] == 30 and o["relocations"] == 1 and o["cell_evals"] == 3000
        assert o["photon_slots"] == 1001          # the two slices cover the list exactly once
        assert o["iterations"] == 101 and o["time_now_max"] == 2.0


def test_load_imbalance():
    assert shard.load_imbalance([1.0, 1.0, 1.0]) == 1.0
    assert abs(shard.load_imbalance([1.0, 3.0]) - 1.5) < 1e-15


def test_bench_splits_one_job_over_the_gpus_without_gaps_or_overlap():
    """bench.py --gpus N: the SAME list and the same 128 reference ranks whatever N is; GPU g owns rank_slice(128, g, N) and
    those ranks' photons (strong scaling, as Src/mcrat.c:139-164 splits a run over ranks)."""
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    for workload, photons, ranks in (("C5", 0, 0), ("C5", 3_000_001, 96), ("C2", 0, 0)):
        args = argparse.Namespace(workload=workload, photons=photons, ranks=ranks)
        job = bench.job_of(args)
        assert job["ranks"] * job["rank_size"] >= job["photons"] > (job["ranks"] - 1) * job["rank_size"]
        for world in (1, 2, 4, 8):
            covered, rank_owner = 0, []
            for g in range(world):
                ra, rb, lo, hi = bench.my_share(job, g, world, weak=False)
                assert lo == covered and hi > lo
                assert lo == ra * job["rank_size"] and hi == min(rb * job["rank_size"], job["photons"])
                covered = hi
                rank_owner += [g] * (rb - ra)
            assert covered == job["photons"] and len(rank_owner) == job["ranks"]
            assert max(rank_owner.count(g) for g in range(world)) - min(rank_owner.count(g) for g in range(world)) <= 1
        # weak scaling: every GPU runs the whole job
        assert bench.my_share(job, 3, 8, weak=True) == (0, job["ranks"], 0, job["photons"])
