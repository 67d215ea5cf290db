"""CPU: the NumPy oracle against OpenCV itself (the reference's arithmetic)."""
import numpy as np
import pytest

cv2 = pytest.importorskip('cv2')
from oracle import hamming as oh


def _rand(rng, n, low_entropy=False):
    return rng.integers(0, 2 if low_entropy else 256, size=(n, 32), dtype=np.uint8)


@pytest.mark.parametrize('low', [False, True])
@pytest.mark.parametrize('nq,nt', [(50, 70), (1, 5), (64, 2), (33, 1), (200, 300)])
def test_knn2_matches_cv2(low, nq, nt):
    rng = np.random.default_rng(nq * 1000 + nt + low)
    q, t = _rand(rng, nq, low), _rand(rng, nt, low)
    knn = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=False).knnMatch(q, t, k=2)
    idx, dist = oh.knn2(q, t)
    assert len(knn) == nq
    for i, row in enumerate(knn):
        assert len(row) == min(2, nt)
        for j, m in enumerate(row):
            assert m.queryIdx == i and m.trainIdx == idx[i, j] and m.distance == dist[i, j]
        for j in range(len(row), 2):
            assert idx[i, j] == -1


@pytest.mark.parametrize('low', [False, True])
@pytest.mark.parametrize('nq,nt', [(40, 60), (60, 40), (1, 1), (128, 500)])
def test_cross_check_matches_cv2(low, nq, nt):
    rng = np.random.default_rng(7 + nq + nt + low)
    q, t = _rand(rng, nq, low), _rand(rng, nt, low)
    ms = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(q, t)
    qi, ti, d = oh.cross_check(q, t)
    assert [m.queryIdx for m in ms] == qi.tolist()
    assert [m.trainIdx for m in ms] == ti.tolist()
    assert [m.distance for m in ms] == d.astype(float).tolist()


def test_ratio_integer_form_is_exact():
    # checkpoint_a_selftest.py:71  m.distance < 0.80 * n.distance  <=>  5*d1 < 4*d2
    d = np.arange(0, 257)
    d1, d2 = np.meshgrid(d, d, indexing='ij')
    ref = d1.astype(np.float64) < 0.80 * d2.astype(np.float64)
    assert np.array_equal(ref, oh.ratio_keep(d1, d2, 4, 5))
    ref75 = d1.astype(np.float64) < 0.75 * d2.astype(np.float64)
    assert np.array_equal(ref75, oh.ratio_keep(d1, d2, 3, 4))


def test_flat_top2_equals_knn2_on_one_segment():
    rng = np.random.default_rng(3)
    q, t = _rand(rng, 30, True), _rand(rng, 500, True)
    a = oh.knn2(q, t)
    b = oh.flat_top2(q, t)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
