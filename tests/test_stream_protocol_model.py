"""Model check of the persistent stream's hand-over protocol (mcrat_b200/csrc/frame_loop.cuh: frame_stream_pass_kernel,
frame_stream_event_kernel).  No GPU: the wait conditions of the two kernels are restated as cooperating state machines and run
under thousands of random interleavings.

The protocol under test:
  * pass blocks pull items (iteration k, shard s, slice b) from ONE counter, in that order; an item of iteration k > 0 waits
    until gen[s] >= k, then reads the shard's published state; a halted shard's items are skipped; after the pass the block
    requests the next item, then draws a ticket (arrive[s] += 1);
  * an event block serves its shards s = e, e + E, ... one after the other within an iteration; for a shard it waits until
    arrive[s] >= (k + 1) * bps, runs the event and releases gen[s] = k + 1 (early: before the event's tail); a shard that halts
    (frame end, iteration cap) releases gen[s] = UINT_MAX and counts itself in `halted`;
  * pass blocks leave when halted == S.
What must hold for ANY number of resident pass blocks >= 1 and any timing: no deadlock; every (k, s, b) of a running shard is
passed exactly once, by a block that saw generation k exactly; event k of a shard sees all bps passes of iteration k and none
of iteration k + 1; every shard stops at its own iteration count.
"""
import random

import pytest

UINT_MAX = 0xFFFFFFFF


class Model:
    def __init__(self, S, bps, E, stop_at, rng):
        self.S, self.bps, self.E, self.stop_at, self.rng = S, bps, E, stop_at, rng
        self.work = 0
        self.arrive = [0] * S
        self.gen = [0] * S
        self.pub_halt = [False] * S          # the published shard state's halt flag
        self.pub_iters = [0] * S             # ... and its iteration count
        self.halted = 0
        self.passes = {}                     # (k, s, b) -> block id
        self.events = [0] * S
        self.log = []

    # ---- pass block: a generator that yields whenever it would have to wait or take time -------------------------
    def pass_block(self, ident):
        item = self.work
        self.work += 1
        while True:
            per_iter = self.S * self.bps
            k, rem = divmod(item, per_iter)
            s, b = divmod(rem, self.bps)
            while True:
                if self.halted >= self.S:
                    return
                if k == 0 or self.gen[s] >= k:
                    break
                yield "wait gen"
            halt = self.pub_halt[s] or self.stop_at[s] == 0
            if not halt:
                assert self.gen[s] == k, "a block may only pass iteration k while generation k is the current one"
                assert self.pub_iters[s] == k
                yield "pass"                                   # the pass takes time; other blocks run meanwhile
                assert (k, s, b) not in self.passes
                self.passes[(k, s, b)] = ident
            nxt = self.work                                    # the next item is requested before the ticket is drawn
            self.work += 1
            if not halt:
                yield "fence"
                self.arrive[s] += 1
            item = nxt
            yield "next"

    # ---- event block -------------------------------------------------------------------------------------------------
    def event_block(self, e):
        mine = list(range(e, self.S, self.E))
        k_of = {s: 0 for s in mine}
        running = []
        for s in mine:
            if self.stop_at[s] == 0:                           # stopped at entry
                self.gen[s] = UINT_MAX
                self.halted += 1
            else:
                running.append(s)
            yield "entry"
        while running:
            for s in list(running):
                k = k_of[s]
                while self.arrive[s] < (k + 1) * self.bps:
                    yield "wait arrive"
                for b in range(self.bps):                      # all passes of iteration k are in, none of k + 1
                    assert (k, s, b) in self.passes
                    assert (k + 1, s, b) not in self.passes
                yield "event head"
                self.events[s] += 1
                halt = (k + 1) >= self.stop_at[s]
                self.pub_iters[s] = k + 1                      # publish the state, then release (early release)
                self.pub_halt[s] = halt
                self.gen[s] = k + 1
                yield "event tail"                             # the scatter's second half, the mini-pass
                k_of[s] = k + 1
                if halt:
                    self.gen[s] = UINT_MAX
                    self.halted += 1
                    running.remove(s)


def run(S, bps, E, W, stop_at, seed):
    rng = random.Random(seed)
    m = Model(S, bps, E, stop_at, rng)
    procs = [m.event_block(e) for e in range(E)] + [m.pass_block(w) for w in range(W)]
    alive = list(range(len(procs)))
    steps = 0
    waiting_streak = 0
    while alive:
        i = rng.choice(alive)
        try:
            what = next(procs[i])
            waiting_streak = waiting_streak + 1 if what.startswith("wait") else 0
        except StopIteration:
            alive.remove(i)
            waiting_streak = 0
        steps += 1
        # every live process reporting "wait" many times in a row under a fair random scheduler = nobody can move
        assert waiting_streak < 200 * (len(procs) + 1), "deadlock: %d processes only wait" % len(alive)
        assert steps < 2_000_000
    return m


@pytest.mark.parametrize("seed", range(40))
def test_random_interleavings_neither_deadlock_nor_reorder(seed):
    rng = random.Random(1000 + seed)
    S = rng.randint(1, 7)
    bps = rng.randint(1, 4)
    E = rng.randint(1, S)
    while (S + E - 1) // E > 8:
        E += 1
    W = rng.randint(1, 9)                                      # resident pass blocks: any number >= 1 must do
    stop_at = [rng.choice([0, 1, 2, 3, 5, 8]) for _ in range(S)]  # 0: the shard is stopped before the launch
    for sched in range(25):
        m = run(S, bps, E, W, stop_at, seed * 100 + sched)
        assert m.halted == S
        for s in range(S):
            assert m.events[s] == stop_at[s], (s, m.events, stop_at)
            for k in range(stop_at[s]):
                for b in range(bps):
                    assert (k, s, b) in m.passes
            assert all(not (kk >= stop_at[s] and ss == s) for (kk, ss, _b) in m.passes), "a halted shard was passed again"
            assert m.arrive[s] == stop_at[s] * bps


def test_one_pass_block_and_one_event_block_suffice():
    # the degenerate geometry: everything funnels through two blocks
    m = run(S=5, bps=3, E=1, W=1, stop_at=[4, 0, 2, 4, 1], seed=7)
    assert m.events == [4, 0, 2, 4, 1]


# ---------------------------------------------------------------------------------------------------------------------
# the start-up handshake (stream_handshake): GO only if every event block is resident, ABORT if somebody waited too long,
# one decision for all (atomicCAS on one word), taken before anything is touched
# ---------------------------------------------------------------------------------------------------------------------
def _handshake_run(n_evt, n_pass, evt_present, patience, seed):
    rng = random.Random(seed)
    state = {"word": 0, "ready": 0}
    outcome = {}

    def cas(new):
        if state["word"] == 0:
            state["word"] = new

    def block(ident, is_pass, present):
        if not present:                      # a block that never becomes resident (the other grid fills the device)
            return
            yield
        if not is_pass:
            state["ready"] += 1
        spins = 0
        while True:
            if state["word"] != 0:
                outcome[ident] = state["word"]
                return
            if is_pass and state["ready"] >= n_evt:
                cas(1)
                continue
            spins += 1
            if spins > patience:
                cas(2)
            yield

    procs = [block(("e", k), False, k < evt_present) for k in range(n_evt)] + [block(("p", k), True, True) for k in range(n_pass)]
    alive = list(range(len(procs)))
    steps = 0
    while alive:
        i = rng.choice(alive)
        try:
            next(procs[i])
        except StopIteration:
            alive.remove(i)
        steps += 1
        assert steps < 1_000_000
    return outcome


@pytest.mark.parametrize("seed", range(30))
def test_handshake_reaches_one_decision(seed):
    rng = random.Random(seed)
    n_evt, n_pass = rng.randint(1, 6), rng.randint(1, 6)
    # all event blocks resident, generous patience: GO
    out = _handshake_run(n_evt, n_pass, n_evt, patience=10 ** 6, seed=seed)
    assert set(out.values()) == {1} and len(out) == n_evt + n_pass
    # one event block never shows up: everybody who is there agrees on ABORT
    out = _handshake_run(n_evt, n_pass, n_evt - 1, patience=50, seed=seed)
    assert set(out.values()) == {2} and len(out) == n_evt - 1 + n_pass
    # impatient blocks racing with the GO vote: whatever wins, it is ONE decision
    out = _handshake_run(n_evt, n_pass, n_evt, patience=rng.randint(0, 12), seed=seed)
    assert len(set(out.values())) == 1
