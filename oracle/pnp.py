"""TEST INFRASTRUCTURE - ctypes front end of the C oracle (oracle/cvmath.c, oracle/pnp_ransac.c).

Restates cv2.solvePnPRansac / solvePnP(EPNP) / projectPoints as used at
visual_landmark_matcher.py:342-356 and checkpoint_a_selftest.py:78-90 and exposes the
intermediate results cv2 hides (minimal sets, per-hypothesis models and inlier counts, the
early-stop trace) so that each CUDA stage can be checked on its own.
Pinned against cv2 4.13.0 in tests/test_oracle_pnp.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'liboracle.so')


def _load():
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith('.c')]
    if (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(['make', '-s', '-C', _HERE])
    return C.CDLL(_SO)


_L = _load()
_vp, _i, _d, _f = C.c_void_p, C.c_int, C.c_double, C.c_float
_L.orc_solvepnp_epnp.argtypes = [_vp, _vp, _i, _d, _d, _d, _d, _vp, _vp]
_L.orc_reproj_err.argtypes = [_vp, _vp, _i, _vp, _vp, _d, _d, _d, _d, _vp, _vp]
_L.orc_ransac_sets.argtypes = [_i, _i, _vp]
_L.orc_ransac_update_niters.argtypes = [_d, _d, _i, _i]
_L.orc_lm_refine.argtypes = [_vp, _vp, _i, _d, _d, _d, _d, _vp, _vp, _vp]
_L.orc_pnp_ransac.argtypes = [_vp, _vp, _i, _d, _d, _d, _d, _i, _f, _d, _i] + [_vp] * 11
_L.orc_svd.argtypes = [_vp, _i, _i, _vp, _vp, _vp]
_L.orc_solve_svd.argtypes = [_vp, _i, _i, _vp, _vp]
_L.orc_invert3_svd.argtypes = [_vp, _vp]
_L.orc_rodrigues_m2v.argtypes = [_vp, _vp]
_L.orc_rodrigues_v2m.argtypes = [_vp, _vp]

FX = FY = 320.0
CX, CY = 320.0, 240.0


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def svd(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    m, n = A.shape
    w, u, vt = np.zeros(n), np.zeros((m, n)), np.zeros((n, n))
    _L.orc_svd(_p(A), m, n, _p(w), _p(u), _p(vt))
    return w, u, vt


def solve_svd(A, b):
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64).ravel()
    x = np.zeros(A.shape[1])
    _L.orc_solve_svd(_p(A), A.shape[0], A.shape[1], _p(b), _p(x))
    return x


def invert3_svd(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    out = np.zeros((3, 3))
    _L.orc_invert3_svd(_p(A), _p(out))
    return out


def rodrigues(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    if x.size == 3:
        R = np.zeros((3, 3))
        _L.orc_rodrigues_v2m(_p(x.ravel()), _p(R))
        return R
    r = np.zeros(3)
    _L.orc_rodrigues_m2v(_p(x), _p(r))
    return r


def solvepnp_epnp(obj, img, fx=FX, fy=FY, cx=CX, cy=CY):
    """cv2.solvePnP(obj f32[n,3], img f32[n,2], K, zeros, flags=SOLVEPNP_EPNP) -> rvec, tvec."""
    obj = np.ascontiguousarray(obj, dtype=np.float32)
    img = np.ascontiguousarray(img, dtype=np.float32)
    r, t = np.zeros(3), np.zeros(3)
    _L.orc_solvepnp_epnp(_p(obj), _p(img), len(obj), fx, fy, cx, cy, _p(r), _p(t))
    return r, t


def reproj_err(obj, img, rvec, tvec, fx=FX, fy=FY, cx=CX, cy=CY):
    """-> (err f32[n] squared reprojection error as RANSAC sees it, proj f32[n,2] = projectPoints)."""
    obj = np.ascontiguousarray(obj, dtype=np.float32)
    img = np.ascontiguousarray(img, dtype=np.float32)
    rvec = np.ascontiguousarray(rvec, dtype=np.float64).ravel()
    tvec = np.ascontiguousarray(tvec, dtype=np.float64).ravel()
    n = len(obj)
    err = np.zeros(n, dtype=np.float32)
    proj = np.zeros((n, 2), dtype=np.float32)
    _L.orc_reproj_err(_p(obj), _p(img), n, _p(rvec), _p(tvec), fx, fy, cx, cy, _p(err), _p(proj))
    return err, proj


def ransac_sets(n, iters=200):
    sets = np.zeros((iters, 5), dtype=np.int32)
    _L.orc_ransac_sets(n, iters, _p(sets))
    return sets


def update_niters(conf, ep, model_points, niters):
    return _L.orc_ransac_update_niters(conf, ep, model_points, niters)


def lm_refine(obj, img, rvec, tvec, fx=FX, fy=FY, cx=CX, cy=CY):
    obj = np.ascontiguousarray(obj, dtype=np.float64)
    img = np.ascontiguousarray(img, dtype=np.float64)
    r = np.array(rvec, dtype=np.float64).ravel().copy()
    t = np.array(tvec, dtype=np.float64).ravel().copy()
    work = np.zeros(2 * len(obj) + 8)
    _L.orc_lm_refine(_p(obj), _p(img), len(obj), fx, fy, cx, cy, _p(r), _p(t), _p(work))
    return r, t


def pnp_ransac(obj, img, iters=200, thr=3.0, conf=0.99, refine=True, fx=FX, fy=FY, cx=CX, cy=CY):
    """solvePnPRansac(...ITERATIVE) with every intermediate exposed."""
    obj = np.ascontiguousarray(obj, dtype=np.float32).reshape(-1, 3)
    img = np.ascontiguousarray(img, dtype=np.float32).reshape(-1, 2)
    n = len(obj)
    sets = np.full((iters, 5), -1, dtype=np.int32)
    counts = np.full(iters, -1, dtype=np.int32)
    models = np.zeros((iters, 6))
    best = np.zeros(1, dtype=np.int32)
    nit = np.zeros(1, dtype=np.int32)
    mask = np.zeros(max(n, 1), dtype=np.uint8)
    r, t = np.zeros(3), np.zeros(3)
    errb = np.zeros(max(n, 1), dtype=np.float32)
    maskb = np.zeros(max(n, 1), dtype=np.uint8)
    work = np.zeros(7 * max(n, 1) + 16)
    ok = _L.orc_pnp_ransac(_p(obj), _p(img), n, fx, fy, cx, cy, iters, thr, conf, int(refine), _p(sets),
                           _p(counts), _p(models), _p(best), _p(nit), _p(mask), _p(r), _p(t), _p(errb),
                           _p(maskb), _p(work))
    if ok < 0:
        raise ValueError('n < 5 is outside the reference call sites (MIN_MATCHES = 10)')
    inl = np.nonzero(mask[:n])[0].astype(np.int32)
    return {'ok': bool(ok), 'rvec': r, 'tvec': t, 'inliers': inl, 'mask': mask[:n].copy(), 'sets': sets,
            'counts': counts, 'models': models, 'best_iter': int(best[0]), 'niters': int(nit[0])}


def mean_reproj_error(obj, img, inl, rvec, tvec):
    """The a6 gate (matcher:353-356): mean L2 pixel error over the RANSAC inliers, float32."""
    _, proj = reproj_err(obj[inl], img[inl], rvec, tvec)
    return float(np.linalg.norm(proj.reshape(-1, 2) - np.asarray(img, dtype=np.float32)[inl], axis=1).mean())
