"""CPU: the round structure of k_pnp_sets (csrc/pnp.cu) restated in Python against the oracle's sequential generator.

The kernel lets 32 lanes take 32 consecutive RANSAC iterations from a table of raw cv::RNG outputs on the assumption
that no lane before them re-drew a duplicate; everything up to and including the first lane that did stands, the next
round starts behind it; a problem that would run past the table falls back to the sequential recurrence.  This test
replays exactly that control flow (including the fall-back decision) for small n, where re-draws are the rule, and for
the production sizes, and compares every set with oracle/pnp.py::ransac_sets (itself pinned against cv2)."""
import numpy as np
import pytest

from oracle import pnp as op


def _raw_stream(count):
    st = (1 << 64) - 1
    out = np.zeros(count, dtype=np.uint64)
    for i in range(count):
        st = ((st & 0xFFFFFFFF) * 4164903690 + (st >> 32)) & ((1 << 64) - 1)
        out[i] = st & 0xFFFFFFFF
    return out


def _rounds(n, iters, raw):
    """-> (sets or None when the kernel would fall back, number of rounds)"""
    sets = np.full((iters, 5), -1, dtype=np.int32)
    base, it0, rounds = 0, 0, 0
    while it0 < iters:
        rounds += 1
        lanes = []
        for lane in range(32):
            it, pos, idx, bad = it0 + lane, base + 5 * lane, [], False
            if it < iters:
                for i in range(5):
                    while True:
                        if pos >= len(raw):
                            bad = True
                            break
                        x = int(raw[pos]) % n
                        pos += 1
                        dup = x in idx[:i]
                        if len(idx) > i:
                            idx[i] = x
                        else:
                            idx.append(x)
                        if not dup:
                            break
                    if bad:
                        break
            lanes.append((it, pos, idx, bad, pos - (base + 5 * lane)))
        redraw = [l for l, (it, pos, idx, bad, used) in enumerate(lanes) if it < iters and used != 5]
        last = redraw[0] if redraw else 31
        if lanes[last][3]:
            return None, rounds
        for l in range(last + 1):
            it, pos, idx, bad, used = lanes[l]
            if it < iters:
                sets[it] = idx
        base = lanes[last][1]
        it0 += last + 1
    return sets, rounds


@pytest.mark.parametrize('n', [6, 7, 10, 11, 33, 334, 500, 1000])
def test_rounds_equal_the_sequential_generator(n):
    raw = _raw_stream(8192)
    sets, rounds = _rounds(n, 200, raw)
    assert sets is not None
    assert np.array_equal(sets, op.ransac_sets(n, 200)), n
    if n >= 334:
        assert rounds <= 20          # re-draws are rare: about one round per 32 iterations plus one per re-draw


def test_table_overrun_is_detected():
    raw = _raw_stream(1000)          # 200 iterations need at least 1000 draws: n = 6 re-draws all the time
    sets, _ = _rounds(6, 200, raw)
    assert sets is None              # the kernel hands such a problem to the sequential recurrence
    sets, _ = _rounds(1000, 199, raw)
    assert sets is None or np.array_equal(sets, op.ransac_sets(1000, 199))
