"""Model check (CPU, numpy) of the warp-wide Maxwell-Juttner sampling (warp_mj_gamma in mcrat_b200/csrc/device_math.cuh):
evaluating the rejection trials of Src/electron.c:207-226 64 at a time from the same uniform stream and taking the
first accepted one in stream order gives the sequential loop's gamma and leaves the stream at the same position --
including trials with x < 1, whose NaN acceptance function the reference treats as a rejection, and the hand-over to the
sequential loop after a bounded number of rounds.  The device code is compared with the sequential device loop by
tests/test_gpu_parity.py::test_warp_wide_maxwell_juttner_sampling_equals_the_sequential_loop."""
import numpy as np
import pytest
from scipy.special import kn


def sequential(u, start, theta, k2):
    d = start
    while True:
        x = u[d] * (1 + 100 * theta)
        with np.errstate(invalid="ignore", divide="ignore"):
            bx = np.sqrt(1 - (1 / (x * x)))
            y = u[d + 1] / 2.0
            f = x * x * (bx / k2) * np.exp(-1 * x / theta)
        d += 2
        if not (np.isnan(f) or y > f):
            return x, d


def warp_wide(u, start, theta, k2, max_rounds):
    for r in range(max_rounds):
        for half in (0, 32):
            lanes = np.arange(32)
            dd = start + 2 * (64 * r + half + lanes)
            x = u[dd] * (1 + 100 * theta)
            with np.errstate(invalid="ignore", divide="ignore"):
                bx = np.sqrt(1 - (1 / (x * x)))
                y = u[dd + 1] / 2.0
                f = x * x * (bx / k2) * np.exp(-1 * x / theta)
            acc = ~(np.isnan(f) | (y > f))
            if acc.any():
                j = int(np.argmax(acc))  # first accepted lane = __ffs(ballot) - 1
                return x[j], start + 2 * (64 * r + half + j + 1)
    return None, start + 128 * max_rounds  # the caller continues sequentially from here


@pytest.mark.parametrize("theta", [0.0017, 0.003, 0.02, 0.3, 1.0, 3.0])
@pytest.mark.parametrize("start", [0, 1, 7])  # even and odd positions of the event's stream
def test_first_accepted_trial_in_stream_order(theta, start):
    rng = np.random.default_rng(int(theta * 1e6) + start)
    k2 = kn(2, 1.0 / theta)
    u = rng.random(400_000)
    pos = start
    for _ in range(50):
        xs, ps = sequential(u, pos, theta, k2)
        for max_rounds in (0, 1, 1 << 13):
            xw, pw = warp_wide(u, pos, theta, k2, max_rounds)
            if xw is None:  # hand-over: the sequential loop continues behind the rejected trials
                xw, pw = sequential(u, pw, theta, k2)
            assert xw == xs and pw == ps and xs >= 1.0, (theta, start, max_rounds)
        pos = ps
