"""cv2-shaped shims for the three OpenCV call sites on the hot path.

A maintainer swaps, in visual_landmark_matcher.py / checkpoint_a_selftest.py,

    self.matcher = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True)     (matcher:211)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=False)              (selftest:46)
    cv2.solvePnPRansac(...), cv2.projectPoints(...)                     (matcher:342-353)

for the objects below; argument meaning, return shapes and error behaviour follow cv2
(INTEGRATION.md shows the three-line patch).  Everything executes on the GPU through the C ABI;
a missing extension or GPU raises.
"""
import numpy as np

from . import _lib
from .library import LandmarkLibrary

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
