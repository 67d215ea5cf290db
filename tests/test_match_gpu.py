"""GPU parity: CUDA matcher through the C ABI vs the NumPy oracle (bit-exact indices/distances).

Covers SURVEY.md section 8a rows a1 (knnMatch k=2), a2 (Lowe ratio), a3 (crossCheck) incl.
heavy-tie inputs, ragged keyframes, <2-row keyframes, empty candidate slots and short frames.
"""
import numpy as np
import pytest

from oracle import hamming as oh

pytestmark = pytest.mark.gpu


def _lib_and_frames(seed, n_kf, n_desc, B, nq, ragged=True, low=False, planted=True):
    from nclt_slam_project_b200 import synth
    from nclt_slam_project_b200.library import LandmarkLibrary
    data = synth.make_library(seed, n_kf=n_kf, n_desc=n_desc, ragged=ragged)
    if low:
        for lm in data['landmarks']:
            lm['descriptors'] &= 1
    desc, pts2d, kstar, _ = synth.make_frame_batch(
        data, range(seed * 100, seed * 100 + B), n_desc=nq, n_planted=(nq // 2 if planted else 0), low_entropy=low)
    return data, LandmarkLibrary.from_pkl_dict(data), desc


@pytest.mark.parametrize('low', [False, True])
@pytest.mark.parametrize('n_kf,n_desc,B,nq', [(6, 300, 3, 257), (4, 1000, 2, 1000), (5, 40, 2, 1100)])
def test_knn2_bit_exact(ctx, low, n_kf, n_desc, B, nq):
    data, lib, desc = _lib_and_frames(11, n_kf, n_desc, B, nq, low=low)
    idx, dist = lib.knn2(desc)
    assert idx.shape == (B, n_kf, nq, 2)
    for b in range(B):
        for k in range(n_kf):
            ri, rd = oh.knn2(desc[b], data['landmarks'][k]['descriptors'])
            assert np.array_equal(idx[b, k], ri), (b, k)
            assert np.array_equal(dist[b, k].astype(np.int32), rd), (b, k)


def test_knn2_edge_cases(ctx):
    from nclt_slam_project_b200.library import LandmarkLibrary
    rng = np.random.default_rng(5)
    kfs = [rng.integers(0, 256, (n, 32), dtype=np.uint8) for n in (0, 1, 2, 513, 1025)]
    lib = LandmarkLibrary(kfs)
    q = rng.integers(0, 256, (2, 70, 32), dtype=np.uint8)
    q_n = np.array([70, 33], dtype=np.int32)
    cand = np.array([[0, 1, 2, 3, 4], [4, -1, 3, 1, -1]], dtype=np.int32)
    idx, dist = lib.knn2(q, q_n, cand)
    for b in range(2):
        for c in range(5):
            k = cand[b, c]
            if k < 0:
                assert (idx[b, c] == -1).all() and (dist[b, c] == 65535).all()
                continue
            ri, rd = oh.knn2(q[b, :q_n[b]], kfs[k])
            assert np.array_equal(idx[b, c, :q_n[b]], ri), (b, c)
            assert np.array_equal(dist[b, c, :q_n[b]].astype(np.int32), rd & 0xFFFF), (b, c)
            assert (idx[b, c, q_n[b]:] == -1).all()


@pytest.mark.parametrize('low', [False, True])
def test_ratio_bit_exact(ctx, low):
    data, lib, desc = _lib_and_frames(23, 5, 400, 3, 500, low=low)
    pairs, n = lib.ratio(desc)
    total = 0
    for b in range(desc.shape[0]):
        for k in range(5):
            t = data['landmarks'][k]['descriptors']
            qi, ti, _ = oh.knn2_ratio(desc[b], t)
            assert n[b, k] == len(qi), (b, k)
            assert np.array_equal(pairs[b, k, :len(qi), 0], qi)
            assert np.array_equal(pairs[b, k, :len(qi), 1], ti)
            total += len(qi)
    if not low:
        assert total > 100      # the planted correspondences survive the ratio test


@pytest.mark.parametrize('low', [False, True])
@pytest.mark.parametrize('nq', [90, 500])
def test_cross_check_bit_exact(ctx, low, nq):
    data, lib, desc = _lib_and_frames(31, 6, 300, 2, nq, low=low)
    q_n = np.array([nq, max(1, nq - 17)], dtype=np.int32)
    cand = np.array([[0, 1, 2, 3], [5, -1, 4, 0]], dtype=np.int32)
    pairs, dist, n = lib.cross(desc, q_n, cand)
    for b in range(2):
        for c in range(4):
            k = cand[b, c]
            if k < 0:
                assert n[b, c] == 0
                continue
            t = data['landmarks'][k]['descriptors']
            qi, ti, d = oh.cross_check(t, desc[b, :q_n[b]])      # match(desc_t, desc_curr)
            assert n[b, c] == len(qi), (b, c)
            assert np.array_equal(pairs[b, c, :len(qi), 0], qi)
            assert np.array_equal(pairs[b, c, :len(qi), 1], ti)
            assert np.array_equal(dist[b, c, :len(qi)].astype(np.int32), d)


@pytest.mark.parametrize('low', [False, True])
@pytest.mark.parametrize('nq,rows', [(1100, [0, 1, 200, 300, 600, 800, 1030, 1100]), (640, [500, 513, 640, 255, 0, 256]),
                                     (200, [200, 31, 0])])
def test_cross_check_live_register_rows(ctx, low, nq, rows):
    """k_hamming_cross computes only the register rows of a CTA that hold frame rows: every live count of the R = 4 / 2 / 1
    kernels, frames that spill into the second CTA of an item, empty frames - against the oracle."""
    B = len(rows)
    data, lib, desc = _lib_and_frames(47, 3, 130, B, nq, low=low)
    q_n = np.array(rows, dtype=np.int32)
    cand = np.array([[b % 3, -1, (b + 1) % 3] for b in range(B)], dtype=np.int32)
    pairs, dist, n = lib.cross(desc, q_n, cand)
    for b in range(B):
        for c in range(3):
            k = cand[b, c]
            if k < 0 or q_n[b] == 0:
                assert n[b, c] == 0, (b, c)
                continue
            t = data['landmarks'][k]['descriptors']
            qi, ti, d = oh.cross_check(t, desc[b, :q_n[b]])
            assert n[b, c] == len(qi), (b, c, q_n[b])
            assert np.array_equal(pairs[b, c, :len(qi), 0], qi) and np.array_equal(pairs[b, c, :len(qi), 1], ti), (b, c, q_n[b])
            assert np.array_equal(dist[b, c, :len(qi)].astype(np.int32), d)


def test_cv2_shaped_bfmatcher(ctx):
    """The drop-in objects behave like cv2.BFMatcher at the reference's call sites."""
    cv2 = pytest.importorskip('cv2')
    from nclt_slam_project_b200 import cv2_compat as g
    rng = np.random.default_rng(9)
    q = rng.integers(0, 4, (120, 32), dtype=np.uint8)
    t = rng.integers(0, 4, (75, 32), dtype=np.uint8)
    ref = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=False).knnMatch(q, t, k=2)
    got = g.BFMatcher(g.NORM_HAMMING, crossCheck=False).knnMatch(q, t, k=2)
    assert [[(m.queryIdx, m.trainIdx, m.distance) for m in r] for r in ref] == \
           [[(m.queryIdx, m.trainIdx, m.distance) for m in r] for r in got]
    ref = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(t, q)
    got = g.BFMatcher(g.NORM_HAMMING, crossCheck=True).match(t, q)
    assert [(m.queryIdx, m.trainIdx, m.distance) for m in ref] == \
           [(m.queryIdx, m.trainIdx, m.distance) for m in got]
    # single-row train set: one neighbour per query, like cv2
    got1 = g.BFMatcher().knnMatch(q, t[:1], k=2)
    assert all(len(r) == 1 for r in got1)
    with pytest.raises(g.error):
        g.BFMatcher().knnMatch(q.astype(np.float32), t)


def test_library_append(ctx):
    from nclt_slam_project_b200.library import LandmarkLibrary
    rng = np.random.default_rng(2)
    kfs = [rng.integers(0, 256, (50, 32), dtype=np.uint8) for _ in range(3)]
    lib = LandmarkLibrary(kfs[:2])
    lib.append(kfs[2])
    q = rng.integers(0, 256, (1, 40, 32), dtype=np.uint8)
    idx, dist = lib.knn2(q)
    for k in range(3):
        ri, rd = oh.knn2(q[0], kfs[k])
        assert np.array_equal(idx[0, k], ri) and np.array_equal(dist[0, k].astype(np.int32), rd)
