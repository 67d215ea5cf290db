"""GPU parity of crossCheck on the tensor cores (engine tensor4, every frame against every keyframe): the index-carrying
cells of k_tc4_top2<true> must give exactly cv2.BFMatcher(NORM_HAMMING, crossCheck=True).match(desc_t, desc_curr) -
pairs in teach-row order, lowest index on ties in BOTH directions, distances - as restated by oracle/hamming.py
(reference call site: scripts/common/visual_landmark_matcher.py:327; exp 63's whole-library ranking :314-345)."""
import numpy as np
import pytest

from oracle import hamming as oh

pytestmark = pytest.mark.gpu


def _check(kfs, q, q_n):
    from nclt_slam_project_b200 import _lib
    from nclt_slam_project_b200.library import LandmarkLibrary
    res = []
    for e in ('int', 'tensor4'):
        c = _lib.Context(0)
        c.set_engine(e)
        lib = LandmarkLibrary(kfs, ctx=c)
        pairs, dist, n = lib.cross(q, q_n, None)          # cand None: all keyframes
        res.append((pairs.copy(), dist.copy(), n.copy()))
        lib.close()
        c.close()
    (pa, da, na), (pb, db, nb) = res
    B = len(q)
    for b in range(B):
        nqb = q.shape[1] if q_n is None else int(q_n[b])
        for k, t in enumerate(kfs):
            qi, ti, d = oh.cross_check(t, q[b, :nqb]) if len(t) and nqb else ([], [], [])
            assert nb[b, k] == len(qi), (b, k, int(nb[b, k]), len(qi))
            assert np.array_equal(pb[b, k, :len(qi), 0], qi), (b, k)
            assert np.array_equal(pb[b, k, :len(qi), 1], ti), (b, k)
            assert np.array_equal(db[b, k, :len(qi)].astype(np.int32), d), (b, k)
    assert np.array_equal(na, nb)
    for b in range(B):
        for k in range(len(kfs)):
            m = int(na[b, k])
            assert np.array_equal(pa[b, k, :m], pb[b, k, :m]) and np.array_equal(da[b, k, :m], db[b, k, :m])
    return int(nb.sum())


def test_cross_tensor_small(ctx):
    rng = np.random.default_rng(3)
    kfs = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for n in (300, 47, 48, 49, 0, 1000, 1, 241)]
    q = rng.integers(0, 256, (3, 500, 32), dtype=np.uint8)
    q[0, :100] = kfs[0][:100]
    q[1, 50:150] = kfs[5][600:700]
    assert _check(kfs, q, np.array([500, 333, 1], dtype=np.int32)) > 200


def test_cross_tensor_heavy_ties(ctx):
    """2-bit descriptors: almost every distance is tied many times over - the lowest-index rule decides everything."""
    rng = np.random.default_rng(4)
    kfs = [rng.integers(0, 4, (n, 32), dtype=np.uint8) & 1 for n in (240, 480, 96, 720, 5)]
    for t in kfs:                                         # exact duplicates inside a keyframe and across tiles
        if len(t) > 100:
            t[len(t) - 40:] = t[:40]
    q = rng.integers(0, 4, (2, 300, 32), dtype=np.uint8) & 1
    q[0, :30] = kfs[1][:30]
    q[0, 30:60] = kfs[1][:30]
    _check(kfs, q, None)


@pytest.mark.timeout(600)
def test_cross_tensor_random_shapes(ctx):
    rng = np.random.default_rng(20261019)
    specials = [0, 1, 2, 47, 48, 49, 95, 96, 239, 240, 241, 287, 288, 480, 1000]
    for case in range(8):
        n_kf = int(rng.integers(1, 12))
        counts = [int(rng.choice(specials)) if rng.random() < 0.7 else int(rng.integers(0, 700)) for _ in range(n_kf)]
        hi = 256 if case % 3 else 4
        kfs = [rng.integers(0, hi, (n, 32), dtype=np.uint8) for n in counts]
        B = int(rng.integers(1, 4))
        nq = int(rng.choice([1, 31, 128, 129, 300, 640]))
        q = rng.integers(0, hi, (B, nq, 32), dtype=np.uint8)
        for k, t in enumerate(kfs):
            if len(t) >= 4 and nq >= 4:
                m = min(len(t) // 2, nq // 2, 40)
                q[k % B, :m] = t[:m]
        q_n = rng.integers(1, nq + 1, B).astype(np.int32) if case % 2 else None
        _check(kfs, q, q_n)


def test_cross_tensor_many_frames_span_groups(ctx):
    """More frame rows than one tile group (16 x 240 rows) and a library larger than one group: splits on both sides."""
    rng = np.random.default_rng(5)
    kfs = [rng.integers(0, 256, (int(n), 32), dtype=np.uint8) for n in rng.integers(200, 600, 24)]
    q = rng.integers(0, 256, (12, 400, 32), dtype=np.uint8)
    for b in range(12):
        q[b, :60] = kfs[2 * b][:60]
    assert _check(kfs, q, None) > 12 * 60


def test_cross_tensor_degenerate_libraries(ctx):
    """only empty keyframes; a single one-row keyframe; a one-row frame"""
    rng = np.random.default_rng(6)
    q = rng.integers(0, 256, (2, 40, 32), dtype=np.uint8)
    e = np.zeros((0, 32), dtype=np.uint8)
    assert _check([e, e, e], q, None) == 0
    one = rng.integers(0, 256, (1, 32), dtype=np.uint8)
    assert _check([one], q, None) == 2
    assert _check([one, e, rng.integers(0, 256, (300, 32), dtype=np.uint8)], q[:, :1], None) == 2 * 2
