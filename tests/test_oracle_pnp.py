"""CPU: pin the C oracle for PnP-RANSAC against OpenCV 4.13 itself (the reference's arithmetic
at visual_landmark_matcher.py:342-356).  Bit-exact where stated."""
import numpy as np
import pytest

cv2 = pytest.importorskip('cv2')
from oracle import pnp
from nclt_slam_project_b200 import synth

K = np.array([[320, 0, 320], [0, 320, 240], [0, 0, 1]], dtype=np.float32)
DIST = np.zeros((4, 1), dtype=np.float32)


def test_jacobi_svd_bit_exact():
    rng = np.random.default_rng(0)
    for t in range(60):
        for (m, n) in [(3, 3), (6, 4), (6, 5), (6, 3), (6, 6), (12, 12)]:
            A = rng.normal(size=(m, n))
            if (m, n) == (12, 12) and t % 2 == 0:      # rank-10: what 5-point EPnP feeds it
                M = rng.normal(size=(10, 12))
                A = M.T @ M
            w, u, vt = pnp.svd(A)
            w2, u2, vt2 = cv2.SVDecomp(A)
            assert np.array_equal(w, w2.ravel()) and np.array_equal(u, u2) and np.array_equal(vt, vt2)


def test_solve_invert_rodrigues_bit_exact():
    rng = np.random.default_rng(1)
    for t in range(100):
        for (m, n) in [(6, 4), (6, 3), (6, 5), (6, 6)]:
            A, b = rng.normal(size=(m, n)), rng.normal(size=(m, 1))
            assert np.array_equal(pnp.solve_svd(A, b), cv2.solve(A, b, flags=cv2.DECOMP_SVD)[1].ravel())
        A = rng.normal(size=(3, 3))
        assert np.array_equal(pnp.invert3_svd(A), cv2.invert(A, flags=cv2.DECOMP_SVD)[1])
        r = rng.normal(size=3) * (0.05 if t % 2 else 1.0)
        R2 = cv2.Rodrigues(r)[0]
        assert np.array_equal(pnp.rodrigues(r), R2)
        Rn = R2 + rng.normal(size=(3, 3)) * 1e-9
        assert np.array_equal(pnp.rodrigues(Rn), cv2.Rodrigues(Rn)[0].ravel())


def test_epnp_minimal_solver_bit_exact():
    """5-point EPnP is numerically chaotic; only a bit-faithful restatement
    reproduces cv2 - this one does."""
    rng = np.random.default_rng(5)
    for t in range(60):
        obj, img, _, _ = synth.make_pnp_problem(t, n=60, outlier_frac=0.3)
        for n in (5, 5, 5, 6, 9):
            idx = rng.permutation(60)[:n]
            o, im = np.ascontiguousarray(obj[idx]), np.ascontiguousarray(img[idx])
            ok, r2, t2 = cv2.solvePnP(o, im, K, DIST, flags=cv2.SOLVEPNP_EPNP)
            r, tt = pnp.solvepnp_epnp(o, im)
            assert np.array_equal(r, r2.ravel()) and np.array_equal(tt, t2.ravel()), (t, n)


def test_ransac_sets_are_the_mwc_sequence():
    s = pnp.ransac_sets(137, 200)
    assert s.shape == (200, 5) and s.min() >= 0 and s.max() < 137
    assert all(len(set(r)) == 5 for r in s.tolist())
    # first draw of cv::RNG(-1): state = 0xFFFFFFFF*4164903690 + 0xFFFFFFFF
    st = ((0xFFFFFFFF * 4164903690) + 0xFFFFFFFF) & 0xFFFFFFFFFFFFFFFF
    assert s[0, 0] == (st & 0xFFFFFFFF) % 137


def test_full_ransac_matches_cv2():
    rng = np.random.default_rng(0)
    n_ok = 0
    for t in range(80):
        n = int(rng.integers(10, 400))
        of = float(rng.uniform(0, 0.7))
        obj, img, _, _ = synth.make_pnp_problem(1000 + t, n=n, outlier_frac=of)
        ok2, r2, t2, inl2 = cv2.solvePnPRansac(obj, img, K, DIST, iterationsCount=200,
                                               reprojectionError=3.0, flags=cv2.SOLVEPNP_ITERATIVE)
        o = pnp.pnp_ransac(obj, img)
        assert o['ok'] == ok2
        if not ok2:
            continue
        n_ok += 1
        assert np.array_equal(o['inliers'], inl2.ravel()), t           # identical inlier sets
        assert np.abs(o['rvec'] - r2.ravel()).max() < 1e-7               # bar: 1e-4 rad
        assert np.abs(o['tvec'] - t2.ravel()).max() < 1e-7               # bar: 1 mm
        # the RANSAC model itself (before LM) is bit-identical to cv2's EPnP on the winning set
        idx = o['sets'][o['best_iter']]
        _, rb, tb = cv2.solvePnP(np.ascontiguousarray(obj[idx]), np.ascontiguousarray(img[idx]), K, DIST,
                                 flags=cv2.SOLVEPNP_EPNP)
        assert np.array_equal(o['models'][o['best_iter']], np.concatenate([rb.ravel(), tb.ravel()]))
        # projectPoints (a6 gate) bit-exact
        p2, _ = cv2.projectPoints(obj[inl2[:, 0]], r2, t2, K, DIST)
        _, p1 = pnp.reproj_err(obj[inl2[:, 0]], img[inl2[:, 0]], r2, t2)
        assert np.array_equal(p1, p2.reshape(-1, 2))
    assert n_ok > 60


def test_n_equals_5_and_below():
    obj, img, _, _ = synth.make_pnp_problem(3, n=5, outlier_frac=0.0)
    ok2, r2, t2, inl2 = cv2.solvePnPRansac(obj, img, K, DIST, iterationsCount=200, reprojectionError=3.0,
                                           flags=cv2.SOLVEPNP_ITERATIVE)
    o = pnp.pnp_ransac(obj, img)
    assert o['ok'] == ok2 and np.array_equal(o['inliers'], inl2.ravel())
    assert np.abs(o['rvec'] - r2.ravel()).max() < 1e-7
    with pytest.raises(ValueError):
        pnp.pnp_ransac(obj[:4], img[:4])
