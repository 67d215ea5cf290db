"""cv2-shaped shims for the three OpenCV call sites on the hot path.

A maintainer swaps, in visual_landmark_matcher.py / checkpoint_a_selftest.py,

    self.matcher = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True)     (matcher:211)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=False)              (selftest:46)
    cv2.solvePnPRansac(...), cv2.projectPoints(...)                     (matcher:342-353)
    self.orb = cv2.ORB_create(nfeatures=500)                            (matcher:207, recorder:159)

for the objects below; argument meaning, return shapes and error behaviour follow cv2
(INTEGRATION.md shows the three-line patch).  Everything executes on the GPU through the C ABI;
a missing extension or GPU raises.
"""
import numpy as np

from . import _lib
from .library import LandmarkLibrary
from .orb import KeyPoint, ORB_create  # noqa: F401  (cv2.ORB_create(nfeatures=500) / cv2.KeyPoint, orb.py)

NORM_HAMMING = 6          # cv2.NORM_HAMMING
SOLVEPNP_ITERATIVE = 0    # cv2.SOLVEPNP_ITERATIVE


class error(_lib.NcltError):
    """Stands in for cv2.error (caught at visual_landmark_matcher.py:328)."""


class DMatch:
    __slots__ = ('queryIdx', 'trainIdx', 'imgIdx', 'distance')

    def __init__(self, queryIdx=-1, trainIdx=-1, distance=float('inf'), imgIdx=0):
        self.queryIdx = queryIdx
        self.trainIdx = trainIdx
        self.imgIdx = imgIdx
        self.distance = distance

    def __repr__(self):
        return f'DMatch(q={self.queryIdx}, t={self.trainIdx}, d={self.distance})'


def _check_desc(name, d):
    d = np.asarray(d)
    if d.dtype != np.uint8 or d.ndim != 2 or d.shape[1] != 32:
        raise error(f'{name}: expected uint8[n,32] ORB descriptors, got {d.dtype}{d.shape}')
    return np.ascontiguousarray(d)


class BFMatcher:
    """cv2.BFMatcher(NORM_HAMMING, crossCheck=...) on the GPU.

    knnMatch/match take (queryDescriptors, trainDescriptors) like cv2.  The train set is
    uploaded as a one-keyframe library per call; batch users should hold a LandmarkLibrary and
    call its knn2/ratio/cross methods instead."""

    def __init__(self, normType=NORM_HAMMING, crossCheck=False, ctx=None):
        if normType != NORM_HAMMING:
            raise error('only NORM_HAMMING is implemented (the reference uses nothing else)')
        self.crossCheck = bool(crossCheck)
        self.ctx = ctx or _lib.default_context()

    def knnMatch(self, queryDescriptors, trainDescriptors, k=2):
        if self.crossCheck and k != 1:
            raise error('knnMatch with crossCheck=True requires k == 1 (as in cv2)')
        if k not in (1, 2):
            raise error('k must be 1 or 2')
        q = _check_desc('queryDescriptors', queryDescriptors)
        t = _check_desc('trainDescriptors', trainDescriptors)
        if len(q) == 0:
            return []
        if len(t) == 0:
            return [[] for _ in range(len(q))]
        if self.crossCheck:
            out = [[] for _ in range(len(q))]
            for m in self.match(q, t):
                out[m.queryIdx].append(m)
            return out
        lib = LandmarkLibrary([t], None, ctx=self.ctx)
        try:
            idx, dist = lib.knn2(q[None], None, np.zeros((1, 1), dtype=np.int32))
        finally:
            lib.close()
        idx, dist = idx[0, 0], dist[0, 0]
        out = []
        for i in range(len(q)):
            row = []
            for j in range(k):
                if idx[i, j] >= 0:
                    row.append(DMatch(i, int(idx[i, j]), float(dist[i, j])))
            out.append(row)
        return out

    def match(self, queryDescriptors, trainDescriptors):
        q = _check_desc('queryDescriptors', queryDescriptors)
        t = _check_desc('trainDescriptors', trainDescriptors)
        if len(q) == 0 or len(t) == 0:
            return []
        if not self.crossCheck:
            return [m[0] for m in self.knnMatch(q, t, k=1) if m]
        # crossCheck: cv2's query set plays the "teach keyframe" role, the train set the frame
        lib = LandmarkLibrary([q], None, ctx=self.ctx)
        try:
            pairs, dist, n = lib.cross(t[None], None, np.zeros((1, 1), dtype=np.int32))
        finally:
            lib.close()
        m = int(n[0, 0])
        return [DMatch(int(pairs[0, 0, i, 0]), int(pairs[0, 0, i, 1]), float(dist[0, 0, i])) for i in range(m)]


def _intrinsics(cameraMatrix, distCoeffs):
    K = np.asarray(cameraMatrix, dtype=np.float64)
    if K.shape != (3, 3):
        raise error('cameraMatrix must be 3x3')
    if distCoeffs is not None and np.any(np.asarray(distCoeffs) != 0):
        raise error('only zero distortion is implemented (DIST = zeros, visual_landmark_matcher.py:52)')
    return float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])


def solvePnPRansac(objectPoints, imagePoints, cameraMatrix, distCoeffs, iterationsCount=100,
                   reprojectionError=8.0, confidence=0.99, flags=SOLVEPNP_ITERATIVE, ctx=None):
    """cv2.solvePnPRansac as called at visual_landmark_matcher.py:342-346.

    Returns (ok, rvec f64[3,1], tvec f64[3,1], inliers i32[m,1] | None).  Fewer than 4 points
    raises `error` like cv2.  Exactly 4 points: cv2 takes its P3P branch and returns ok=True with 4
    inliers; that branch is not restated here, and the call returns ok=False (inliers None) instead
    of raising - for both call sites the outcome is the same `continue`, because they reject anything
    below MIN_INLIERS = 10 (matcher:349, selftest:83) and never pass fewer than MIN_MATCHES = 10 points."""
    from .pnp import pnp_ransac_batch
    from ._lib import PnpParams
    if flags != SOLVEPNP_ITERATIVE:
        raise error('only flags=SOLVEPNP_ITERATIVE is implemented (the reference uses nothing else)')
    obj = np.ascontiguousarray(objectPoints, dtype=np.float32).reshape(-1, 3)
    img = np.ascontiguousarray(imagePoints, dtype=np.float32).reshape(-1, 2)
    if len(obj) != len(img) or len(obj) < 4:
        raise error('solvePnPRansac needs >= 4 matching object/image points')
    if len(obj) == 4:        # documented divergence: cv2's P3P branch (ok=True, 4 inliers) is reported as "no model"
        return False, np.zeros((3, 1)), np.zeros((3, 1)), None
    fx, fy, cx, cy = _intrinsics(cameraMatrix, distCoeffs)
    prm = PnpParams(fx, fy, cx, cy, int(iterationsCount), float(reprojectionError), float(confidence), 1)
    o = pnp_ransac_batch(obj[None], img[None], None, prm, ctx=ctx)
    ok = bool(o['ok'][0])
    rvec = o['rvec'][0].reshape(3, 1).copy()
    tvec = o['tvec'][0].reshape(3, 1).copy()
    if not ok:
        return False, rvec, tvec, None
    inl = np.nonzero(o['mask'][0, :len(obj)])[0].astype(np.int32).reshape(-1, 1)
    return True, rvec, tvec, inl


def projectPoints(objectPoints, rvec, tvec, cameraMatrix, distCoeffs, ctx=None):
    """cv2.projectPoints (matcher:353): returns (f32[n,1,2], None)."""
    from .pnp import project_points
    fx, fy, cx, cy = _intrinsics(cameraMatrix, distCoeffs)
    out = project_points(objectPoints, rvec, tvec, fx, fy, cx, cy, ctx=ctx)
    return out.reshape(-1, 1, 2), None


def Rodrigues(rvec):
    """cv2.Rodrigues(rvec) -> (R f64[3,3], None): 3x3 host arithmetic (matcher:361 keeps the
    pose composition in NumPy; SURVEY 8a row a8 'keep on host, not a kernel')."""
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    th = float(np.sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]))
    if th < np.finfo(np.float64).eps:
        return np.eye(3), None
    c, s = np.cos(th), np.sin(th)
    k = r / th
    rrt = np.outer(k, k)
    kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return c * np.eye(3) + (1 - c) * rrt + s * kx, None
