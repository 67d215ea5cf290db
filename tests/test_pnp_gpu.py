"""GPU parity for PnP-RANSAC, staged as section 7 prescribes -
and, because the EPnP restatement is bit-faithful, also end to end:

 (i)   minimal sets vs the MWC sequence            100 % identical
 (ii)  scoring given the oracle's hypotheses       identical counts
 (iii) early-stop replay                           identical winning iteration / niters
 (iv)  LM refinement                               1e-4 rad / 1 mm (observed ~1e-9)
 (v)   the minimal solver itself                   hypotheses equal the oracle's to 1e-9; counts identical
 end-to-end vs the oracle (= cv2 on this container, tests/test_oracle_pnp.py) and vs the golden
 vectors produced by cv2.solvePnPRansac itself (tests/golden/pnp_golden.npz).
"""
import os

import numpy as np
import pytest

from oracle import pnp as opnp

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-4      # rad   (BASELINE.json north_star)
TRANS_TOL = 1e-3    # m


def _problems(seed0, count, nmax=400):
    from nclt_slam_project_b200 import synth
    rng = np.random.default_rng(seed0)
    obj = np.zeros((count, nmax, 3), dtype=np.float32)
    img = np.zeros((count, nmax, 2), dtype=np.float32)
    n = np.zeros(count, dtype=np.int32)
    for p in range(count):
        n[p] = int(rng.integers(10, nmax + 1))
        o, i, _, _ = synth.make_pnp_problem(seed0 * 1000 + p, n=int(n[p]), outlier_frac=float(rng.uniform(0, 0.7)))
        obj[p, :n[p]] = o
        img[p, :n[p]] = i
    return obj, img, n


def test_stage_i_minimal_sets(ctx):
    from nclt_slam_project_b200.pnp import pnp_ransac_batch
    obj, img, n = _problems(1, 6)
    out = pnp_ransac_batch(obj, img, n, debug=True)
    for p in range(len(n)):
        assert np.array_equal(out['sets'][p], opnp.ransac_sets(int(n[p]), 200))


def test_stage_ii_scoring_given_oracle_models(ctx):
    from nclt_slam_project_b200.pnp import pnp_score
    obj, img, n = _problems(2, 8)
    models = np.zeros((len(n), 200, 6))
    ref = np.zeros((len(n), 200), dtype=np.int32)
    for p in range(len(n)):
        sets = opnp.ransac_sets(int(n[p]), 200)
        for it in range(200):
            r, t = opnp.solvepnp_epnp(obj[p, sets[it]], img[p, sets[it]])
            models[p, it, :3], models[p, it, 3:] = r, t
            err, _ = opnp.reproj_err(obj[p, :n[p]], img[p, :n[p]], r, t)
            ref[p, it] = int((err <= np.float32(9.0)).sum())
    got = pnp_score(obj, img, n, models)
    assert np.array_equal(got, ref)


def test_stage_v_minimal_solver_and_end_to_end(ctx):
    from nclt_slam_project_b200.pnp import pnp_ransac_batch
    obj, img, n = _problems(3, 40)
    out = pnp_ransac_batch(obj, img, n, debug=True)
    n_ok = 0
    for p in range(len(n)):
        o = opnp.pnp_ransac(obj[p, :n[p]], img[p, :n[p]])
        # (v) hypotheses: every model the oracle evaluated (iterations before early stop)
        ran = o['counts'] >= 0
        assert np.allclose(out['models'][p][ran], o['models'][ran], rtol=0, atol=1e-9), p
        assert np.array_equal(out['counts'][p][ran], o['counts'][ran]), p
        # (iii) replay
        assert out['best_iter'][p] == o['best_iter'] and out['niters'][p] == o['niters'], p
        assert bool(out['ok'][p]) == o['ok']
        if not o['ok']:
            continue
        n_ok += 1
        assert out['n_inliers'][p] == len(o['inliers'])
        assert np.array_equal(np.nonzero(out['mask'][p, :n[p]])[0], o['inliers'])
        # (iv) refined pose
        assert np.abs(out['rvec'][p] - o['rvec']).max() < ROT_TOL
        assert np.abs(out['tvec'][p] - o['tvec']).max() < TRANS_TOL
        assert np.abs(out['rvec'][p] - o['rvec']).max() < 1e-7      # what we actually reach
        me = opnp.mean_reproj_error(obj[p, :n[p]], img[p, :n[p]], o['inliers'], o['rvec'], o['tvec'])
        assert abs(out['mean_err'][p] - me) < 1e-4
    assert n_ok > 25


def test_small_and_degenerate_problems(ctx):
    from nclt_slam_project_b200.pnp import pnp_ransac_batch
    from nclt_slam_project_b200 import synth
    obj = np.zeros((4, 16, 3), dtype=np.float32)
    img = np.zeros((4, 16, 2), dtype=np.float32)
    n = np.array([5, 3, 16, 0], dtype=np.int32)
    o5, i5, _, _ = synth.make_pnp_problem(3, n=5, outlier_frac=0.0)
    obj[0, :5], img[0, :5] = o5, i5
    o16, i16, _, _ = synth.make_pnp_problem(4, n=16, outlier_frac=0.0)
    obj[2], img[2] = o16, i16
    out = pnp_ransac_batch(obj, img, n)
    ref = opnp.pnp_ransac(o5, i5)
    assert out['ok'][0] == 1 and out['n_inliers'][0] == 5
    assert np.abs(out['rvec'][0] - ref['rvec']).max() < 1e-9      # n == 5: the kernel result, no refine
    assert out['ok'][1] == 0 and out['ok'][3] == 0 and out['n_inliers'][1] == 0
    ref16 = opnp.pnp_ransac(o16, i16)
    assert bool(out['ok'][2]) == ref16['ok'] and out['n_inliers'][2] == len(ref16['inliers'])


def test_golden_vectors_from_cv2(ctx):
    """Outputs of cv2.solvePnPRansac itself (generated in the build container by
    oracle/make_golden.py; cv2 is the reference's arithmetic here)."""
    from nclt_slam_project_b200.pnp import pnp_ransac_batch
    GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    g = np.load(os.path.join(GOLDEN, 'pnp_golden.npz'))
    out = pnp_ransac_batch(g['obj'], g['img'], g['n'])
    assert np.array_equal(out['ok'], g['ok'])
    for p in range(len(g['n'])):
        if not g['ok'][p]:
            continue
        assert out['n_inliers'][p] == g['n_inliers'][p], p
        assert np.array_equal(out['mask'][p], g['mask'][p]), p
        assert np.abs(out['rvec'][p] - g['rvec'][p]).max() < ROT_TOL
        assert np.abs(out['tvec'][p] - g['tvec'][p]).max() < TRANS_TOL
        assert abs(out['mean_err'][p] - g['mean_err'][p]) < 1e-3


def test_cv2_shaped_calls(ctx):
    cv2 = pytest.importorskip('cv2')
    from nclt_slam_project_b200 import cv2_compat as g, synth
    K = np.array([[320, 0, 320], [0, 320, 240], [0, 0, 1]], dtype=np.float32)
    DIST = np.zeros((4, 1), dtype=np.float32)
    obj, img, _, _ = synth.make_pnp_problem(77, n=150, outlier_frac=0.4)
    ok2, r2, t2, inl2 = cv2.solvePnPRansac(obj, img, K, DIST, iterationsCount=200, reprojectionError=3.0,
                                           flags=cv2.SOLVEPNP_ITERATIVE)
    ok, r, t, inl = g.solvePnPRansac(obj, img, K, DIST, iterationsCount=200, reprojectionError=3.0,
                                     flags=g.SOLVEPNP_ITERATIVE)
    assert ok == ok2 and r.shape == (3, 1) and t.shape == (3, 1) and inl.dtype == np.int32
    assert np.array_equal(inl, inl2)
    assert np.abs(r - r2).max() < ROT_TOL and np.abs(t - t2).max() < TRANS_TOL
    p2, _ = cv2.projectPoints(obj[inl2[:, 0]], r2, t2, K, DIST)
    p1, _ = g.projectPoints(obj[inl2[:, 0]], r2, t2, K, DIST)
    assert p1.shape == p2.shape and p1.dtype == np.float32
    assert np.abs(p1 - p2).max() < 1e-3          # bit-equal unless libm sin/cos differ in the last ulp
    R1, _ = g.Rodrigues(r2)
    assert np.abs(R1 - cv2.Rodrigues(r2)[0]).max() < 1e-14
    with pytest.raises(g.error):
        g.solvePnPRansac(obj[:3], img[:3], K, DIST)
