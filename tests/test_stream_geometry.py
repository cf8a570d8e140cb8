"""Launch geometry of the persistent stream (mcrat_b200/csrc/mcrat_b200.cu: frame_stream_evt_blocks, frame_stream_bps,
launch_frame_stream) swept over list lengths and sub-shard counts, without a GPU: the bounds the two kernels rely on must hold
for every combination AUTO or a caller can reach, not only for the ones that were timed."""
import ctypes as C

import pytest

from mcrat_b200 import lib

BLOCKMIN_CAP = 8192      # state.cuh
STREAM_EVT_SHARDS = 8    # frame_loop.cuh
EVT_THREADS = 128        # the event block reads the minima with one load per thread
NUM_SMS = 148


def geometry(cap, shards, sms=NUM_SMS):
    L = lib.load()
    out = [C.c_int(0) for _ in range(6)]
    rc = L.mcrat_b200_debug_stream_geometry(C.c_int(cap), C.c_int(shards), C.c_int(sms), *[C.byref(o) for o in out])
    assert rc == 0
    return dict(zip(("E", "bps", "grid", "S", "fits", "auto"), (o.value for o in out)))


CAPS = [300, 5000, 78125, 10 ** 5, 3 * 10 ** 5, 10 ** 6, 1_250_000, 2_500_000, 5 * 10 ** 6, 10 ** 7, 3 * 10 ** 7, 10 ** 8]
SHARDS = [1, 2, 3, 7, 16, 31, 32, 33, 64, 100, 128, 147, 148, 149, 296, 320, 1021, 1184, 1185, 2048, 4096]


def test_bounds_hold_for_every_list_and_shard_count():
    for cap in CAPS:
        for shards in SHARDS:
            g = geometry(cap, shards)
            S = g["S"]
            assert 1 <= S <= min(shards, cap)
            if not g["fits"]:
                assert not g["auto"]
                continue
            # event blocks: all resident beside the pass blocks, each serving at most STREAM_EVT_SHARDS sub-shards
            assert 1 <= g["E"] <= min(S, NUM_SMS)
            assert -(-S // g["E"]) <= STREAM_EVT_SHARDS
            # block minima: team = bps + 1 slots per shard in an array of BLOCKMIN_CAP, read with one load per event thread
            assert g["bps"] >= 1
            assert S * (g["bps"] + 1) <= BLOCKMIN_CAP, (cap, shards, g)
            assert g["bps"] + 1 <= EVT_THREADS
            # pass grid: 4 blocks on an SM that also holds an event block, 5 elsewhere -- never more than fits beside them
            assert g["grid"] == g["E"] * 4 + (NUM_SMS - g["E"]) * 5
            assert g["grid"] >= 1


def test_auto_picks_the_stream_where_it_was_measured_to_win():
    # the four per-GPU shares of the bench's job (10^7 photons in 128 ranks over 1, 2, 4, 8 GPUs)
    for cap, shards, E, bps in ((10 ** 7, 128, 64, 13), (5 * 10 ** 6, 64, 32, 20), (2_500_000, 32, 32, 39), (1_250_000, 16, 16, 77)):
        g = geometry(cap, shards)
        assert g["auto"] == 1 and g["fits"] == 1
        assert (g["E"], g["bps"]) == (E, bps), g
    # lists that fit in L2, few long shards, single ranks: the cooperative team kernel / the streamed loop
    for cap, shards in ((10 ** 6, 16), (10 ** 5, 16), (10 ** 7, 16), (10 ** 7, 1), (2 * 10 ** 6, 8)):
        assert geometry(cap, shards)["auto"] == 0
    # more sub-shards than event blocks can serve: the stream does not fit at all
    assert geometry(10 ** 7, 4096)["fits"] == 0


def test_other_devices():
    for sms in (1, 8, 80, 132, 160):
        for cap, shards in ((10 ** 7, 128), (10 ** 6, 1000), (5000, 5)):
            L = lib.load()
            out = [C.c_int(0) for _ in range(6)]
            assert L.mcrat_b200_debug_stream_geometry(cap, shards, sms, *[C.byref(o) for o in out]) == 0
            E, bps, grid, S, fits, _auto = (o.value for o in out)
            if fits:
                assert E <= sms and -(-S // E) <= STREAM_EVT_SHARDS and S * (bps + 1) <= BLOCKMIN_CAP and grid >= 1
    assert lib.load().mcrat_b200_debug_stream_geometry(0, 1, 148, None, None, None, None, None, None) == -2
