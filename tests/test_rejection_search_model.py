"""Model check (CPU, pure Python) of the next-candidate search after a Klein-Nishina rejection in the streamed loop
(event_body in mcrat_b200/csrc/event.cuh): the entry of the shard's (time, slot) order that follows the rejected candidate is
found from the pass blocks' minima -- a block whose minimum comes after the rejected candidate offers that minimum, a
block whose minimum is used up is read again, photons re-located in this iteration are not in any minimum and come
from the re-location list -- and must equal what reading every time of the shard gives, for any number of successive
rejections, with ties in time and with re-located photons anywhere in the order.  The device code is compared with the
persistent loop's full read by tests/test_gpu_parity.py::test_klein_nishina_rejections_streamed_equals_persistent."""
import random

import pytest

INF = (float("inf"), 2 ** 31 - 1)


def block_of(j, threads, nblk):
    return (j // threads) % nblk  # pass_body: j = b * THREADS + tid + m * nblk * THREADS


def full_scan_next(tts, prev):
    best = INF
    for i, t in enumerate(tts):
        if prev < (t, i) < best:
            best = (t, i)
    return best


def two_level_next(tts, bm, reloc, prev, threads, nblk):
    best = INF
    for b in range(nblk):
        if prev < bm[b]:
            if bm[b] < best:
                best = bm[b]
        else:  # minimum used up: read the block's photons again
            for i, t in enumerate(tts):
                if block_of(i, threads, nblk) == b and prev < (t, i) < best:
                    best = (t, i)
    for i in reloc:
        if prev < (tts[i], i) < best:
            best = (tts[i], i)
    return best


@pytest.mark.parametrize("seed", range(40))
def test_two_level_search_walks_the_same_order_as_the_full_scan(seed):
    rng = random.Random(seed)
    threads, nblk = rng.choice([(2, 3), (4, 5), (8, 2), (4, 1), (3, 7)])
    n = rng.randint(1, 120)
    levels = rng.choice([3, 10, 10 ** 6])  # few distinct times -> many ties
    tts = [float(rng.randrange(levels)) for _ in range(n)]
    reloc = sorted(rng.sample(range(n), rng.randint(0, min(n, 6))))
    # what the pass leaves behind: per-block minima over the photons it drew a time for (not the re-located ones)
    bm = [INF] * nblk
    for i, t in enumerate(tts):
        if i not in reloc:
            b = block_of(i, threads, nblk)
            if (t, i) < bm[b]:
                bm[b] = (t, i)
    # head of the order = stage 1 of the event: minima + re-location list
    head = min(bm + [(tts[i], i) for i in reloc])
    assert head == min((t, i) for i, t in enumerate(tts))
    prev, walked = head, [head]
    for _ in range(min(n - 1, 16)):  # MAX_DT successive rejections
        a = full_scan_next(tts, prev)
        b = two_level_next(tts, bm, reloc, prev, threads, nblk)
        assert a == b, (seed, prev, a, b)
        if a == INF:
            break
        prev = a
        walked.append(a)
    assert walked == sorted((t, i) for i, t in enumerate(tts))[: len(walked)]
