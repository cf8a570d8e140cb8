"""Shard bookkeeping for multi-GPU runs.

The reference decomposes a run over MPI ranks that never exchange photons or time steps inside
the frame loop (Src/mcrat.c:139-164, 457-479; no MPI call between :609 and :924).  The GPU build
keeps that decomposition: one process per GPU, each owning a contiguous slice of the photon list
(plus a full replica of the cell arrays), optionally subdivided into sub-shards on the device.
The only collective is the per-frame reduction of a handful of counters (scatterings, photon
iterations, relocations ...) for reporting and load-balance checks -- something the reference
does not do (its ranks only log locally) but the north star asks for.
"""
import numpy as np


def rank_slice(n_items, rank, world):
    """Contiguous, near-equal partition of `n_items` over `world` ranks (rank r gets the r-th slice)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world: %r/%r" % (rank, world))
    base, extra = divmod(int(n_items), world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def sub_shard_ranges(n_items, num_shards):
    """Slot ranges of the device-side sub-shards (mirrors layout_shards() in csrc/mcrat_b200.cu):
    equal size ceil(n/S), last one shorter; empty shards are dropped."""
    if n_items <= 0:
        return [(0, 0)]
    s = max(1, min(int(num_shards), int(n_items)))
    size = -(-n_items // s)
    s = -(-n_items // size)
    return [(k * size, min(size, n_items - k * size)) for k in range(s)]


def global_shard_id(rank, shards_per_gpu, sub_shard):
    """Philox shard key of a sub-shard: unique across the whole job."""
    return rank * shards_per_gpu + sub_shard


COUNTER_KEYS = ("scatterings", "relocations", "photon_slots", "cell_evals", "not_found")


def reduce_frame_stats(stats, dist=None, device=None):
    """Sum the per-rank frame counters over all ranks (all_reduce SUM) and take the max of the
    iteration count; returns a dict.  `dist` is torch.distributed (NCCL on GPUs, gloo in tests)."""
    out = dict(stats)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        out["ranks"] = 1
        return out
    import torch
    t = torch.tensor([float(stats[k]) for k in COUNTER_KEYS], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    m = torch.tensor([float(stats["iterations"]), float(stats["time_now"])], dtype=torch.float64, device=device)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    for k, v in zip(COUNTER_KEYS, t.tolist()):
        out[k] = int(v)
    out["iterations"] = int(m[0].item())
    out["time_now_max"] = float(m[1].item())
    out["ranks"] = dist.get_world_size()
    return out


def load_imbalance(per_rank_seconds):
    """max/mean of the per-rank wall time of a frame (1.0 = perfectly balanced)."""
    a = np.asarray(per_rank_seconds, dtype=np.float64)
    return float(a.max() / a.mean()) if a.size and a.mean() > 0 else 1.0
