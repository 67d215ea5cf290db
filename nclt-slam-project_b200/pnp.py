"""Batched PnP-RANSAC on the GPU (host numpy in / numpy out) + the cv2-shaped single calls.

Replaces cv2.solvePnPRansac / cv2.projectPoints at visual_landmark_matcher.py:342-355 and
checkpoint_a_selftest.py:78-86.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib as _c, ptr, PnpParams


def pnp_ransac_batch(obj, img, n=None, params=None, ctx=None, debug=False):
    """obj f32[P,Nmax,3], img f32[P,Nmax,2], n i32[P] -> dict of per-problem results.

    ok u8[P], rvec f64[P,3], tvec f64[P,3], n_inliers i32[P], mask u8[P,Nmax], mean_err f32[P]
    (mean L2 reprojection error of the refined pose over the RANSAC inliers, matcher:353-355);
    with debug=True also sets i32[P,iters,5], models f64[P,iters,6], counts i32[P,iters],
    best_iter i32[P], niters i32[P]."""
    ctx = ctx or _lib.default_context()
    prm = params or PnpParams()
    obj = np.ascontiguousarray(obj, dtype=np.float32)
    img = np.ascontiguousarray(img, dtype=np.float32)
    if obj.ndim == 2:
        obj, img = obj[None], img[None]
    P, Nmax = obj.shape[0], obj.shape[1]
    if img.shape[:2] != (P, Nmax) or obj.shape[2] != 3 or img.shape[2] != 2:
        raise ValueError(f'shape mismatch obj{obj.shape} img{img.shape}')
    n = np.full(P, Nmax, dtype=np.int32) if n is None else np.ascontiguousarray(n, dtype=np.int32).reshape(P)
    it = prm.iterations
    out = {
        'ok': np.zeros(P, dtype=np.uint8), 'rvec': np.zeros((P, 3)), 'tvec': np.zeros((P, 3)),
        'n_inliers': np.zeros(P, dtype=np.int32), 'mask': np.zeros((P, max(Nmax, 1)), dtype=np.uint8),
        'mean_err': np.zeros(P, dtype=np.float32),
    }
    dbg = {}
    if debug:
        dbg = {'sets': np.zeros((P, it, 5), dtype=np.int32), 'models': np.zeros((P, it, 6)),
               'counts': np.zeros((P, it), dtype=np.int32), 'best_iter': np.zeros(P, dtype=np.int32),
               'niters': np.zeros(P, dtype=np.int32)}
    if P == 0 or Nmax == 0:
        out.update(dbg)
        return out
    ctx.check(_c.nclt_pnp_ransac(
        ctx.h, ptr(obj), ptr(img), ptr(n), P, Nmax, C.byref(prm), ptr(out['ok']), ptr(out['rvec']),
        ptr(out['tvec']), ptr(out['n_inliers']), ptr(out['mask']), ptr(out['mean_err']),
        ptr(dbg.get('sets')), ptr(dbg.get('models')), ptr(dbg.get('counts')), ptr(dbg.get('best_iter')),
        ptr(dbg.get('niters'))))
    out.update(dbg)
    return out


def pnp_score(obj, img, n, models, params=None, ctx=None):
    """K4 alone: inlier counts i32[P,iters] of caller-supplied hypotheses f64[P,iters,6]."""
    ctx = ctx or _lib.default_context()
    obj = np.ascontiguousarray(obj, dtype=np.float32)
    img = np.ascontiguousarray(img, dtype=np.float32)
    models = np.ascontiguousarray(models, dtype=np.float64)
    P, Nmax = obj.shape[0], obj.shape[1]
    prm = params or PnpParams(iterations=models.shape[1])
    if models.shape != (P, prm.iterations, 6):
        raise ValueError('models must be f64[P,iters,6]')
    n = np.ascontiguousarray(n, dtype=np.int32).reshape(P)
    counts = np.zeros((P, prm.iterations), dtype=np.int32)
    ctx.check(_c.nclt_pnp_score(ctx.h, ptr(obj), ptr(img), ptr(n), P, Nmax, C.byref(prm), ptr(models), ptr(counts)))
    return counts


def project_points(obj, rvec, tvec, fx=320.0, fy=320.0, cx=320.0, cy=240.0, ctx=None):
    """cv2.projectPoints(obj, rvec, tvec, K, zeros) -> f32[n,2]."""
    ctx = ctx or _lib.default_context()
    obj = np.ascontiguousarray(obj, dtype=np.float32).reshape(-1, 3)
    r = np.ascontiguousarray(rvec, dtype=np.float64).reshape(3)
    t = np.ascontiguousarray(tvec, dtype=np.float64).reshape(3)
    out = np.zeros((len(obj), 2), dtype=np.float32)
    ctx.check(_c.nclt_project_points(ctx.h, ptr(obj), len(obj), ptr(r), ptr(t), fx, fy, cx, cy, ptr(out)))
    return out
