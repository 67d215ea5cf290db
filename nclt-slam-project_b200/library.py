"""Device-resident teach library + batched matching calls (host numpy in / numpy out).

`LandmarkLibrary` holds what visual_landmark_matcher.py:179-187 loads from landmarks.pkl
(per-keyframe 'descriptors' u8[n,32] and 'keypoints_3d_cam' f32[n,3], schema written at
visual_landmark_recorder.py:290-297) in HBM, and exposes the three matching modes of
SURVEY.md section 8a (a1-a3) over batches of frames x candidate keyframes.
"""
import ctypes as C
import pickle

import numpy as np

from . import _lib
from ._lib import lib as _c, ptr, as_c

LOWE_NUM, LOWE_DEN = 4, 5      # LOWE_RATIO = 0.80 (visual_landmark_matcher.py:66) as an exact fraction


class LandmarkLibrary:
    def __init__(self, descriptors, points3d=None, ctx=None):
        """descriptors: list of u8[n_k,32] per keyframe; points3d: list of f32[n_k,3] or None."""
        self.ctx = ctx or _lib.default_context()
        counts = [0 if d is None else len(d) for d in descriptors]
        offs = np.zeros(len(counts) + 1, dtype=np.int32)
        offs[1:] = np.cumsum(counts)
        n = int(offs[-1])
        desc = np.zeros((max(n, 1), 32), dtype=np.uint8)
        pts = np.zeros((max(n, 1), 3), dtype=np.float32)
        for k, d in enumerate(descriptors):
            if counts[k]:
                d = np.asarray(d, dtype=np.uint8)
                if d.ndim != 2 or d.shape[1] != 32:
                    raise ValueError(f'keyframe {k}: descriptors must be u8[n,32], got {d.shape}')
                desc[offs[k]:offs[k + 1]] = d
                if points3d is not None and points3d[k] is not None:
                    pts[offs[k]:offs[k + 1]] = np.asarray(points3d[k], dtype=np.float32).reshape(-1, 3)
        self.offsets = offs
        self.counts = np.asarray(counts, dtype=np.int32)
        h = C.c_void_p()
        self.ctx.check(_c.nclt_lib_create(self.ctx.h, len(counts), ptr(offs), ptr(desc), ptr(pts), C.byref(h)))
        self.h = h

    # -- constructors -------------------------------------------------------------
    @classmethod
    def from_pkl_dict(cls, data, ctx=None):
        lms = data['landmarks']
        return cls([lm['descriptors'] for lm in lms], [lm['keypoints_3d_cam'] for lm in lms], ctx=ctx)

    @classmethod
    def from_pkl(cls, path, ctx=None):
        with open(path, 'rb') as f:
            return cls.from_pkl_dict(pickle.load(f), ctx=ctx)

    @classmethod
    def from_packed(cls, path, ctx=None):
        """Load the flat binary written by recorder.write_packed: offsets / descriptors / 3-D points are already in the
        layout nclt_lib_create takes, so the memory-mapped arrays go to the device without per-keyframe copies."""
        from .recorder import read_packed
        d = read_packed(path)
        self = cls.__new__(cls)
        self.ctx = ctx or _lib.default_context()
        offs = np.ascontiguousarray(d['offsets'], dtype=np.int32)
        n_kf = len(offs) - 1
        desc = np.ascontiguousarray(d['descriptors']) if len(d['descriptors']) else np.zeros((1, 32), dtype=np.uint8)
        pts = np.ascontiguousarray(d['points3d'], dtype=np.float32) if len(d['points3d']) else np.zeros((1, 3), dtype=np.float32)
        self.offsets = offs
        self.counts = np.diff(offs).astype(np.int32)
        self.poses = np.array(d['poses'])
        h = C.c_void_p()
        self.ctx.check(_c.nclt_lib_create(self.ctx.h, n_kf, ptr(offs), ptr(desc), ptr(pts), C.byref(h)))
        self.h = h
        return self

    def append(self, descriptors, points3d=None):
        """Add one keyframe at runtime (visual_landmark_matcher.py:492-496)."""
        d = as_c(descriptors, np.uint8).reshape(-1, 32)
        p = None if points3d is None else as_c(points3d, np.float32).reshape(-1, 3)
        self.ctx.check(_c.nclt_lib_append(self.ctx.h, self.h, ptr(d), ptr(p), len(d)))
        self.offsets = np.append(self.offsets, self.offsets[-1] + len(d)).astype(np.int32)
        self.counts = np.append(self.counts, len(d)).astype(np.int32)

    @property
    def n_keyframes(self):
        return len(self.counts)

    @property
    def max_rows(self):
        return int(self.counts.max()) if len(self.counts) else 0

    def close(self):
        if getattr(self, 'h', None) is not None and self.h.value and self.ctx.h.value:
            _c.nclt_lib_destroy(self.ctx.h, self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ------------------------------------------------------------------
    @staticmethod
    def _prep(q, q_n, cand, n_kf):
        q = as_c(q, np.uint8)
        if q.ndim == 2:
            q = q[None]
        B, Nq = q.shape[0], q.shape[1]
        if q.shape[2] != 32:
            raise ValueError('descriptors must be 32 bytes wide')
        q_n = None if q_n is None else as_c(q_n, np.int32).reshape(B)
        if cand is None:
            C_ = n_kf
        else:
            cand = as_c(cand, np.int32).reshape(B, -1)
            C_ = cand.shape[1]
        return q, q_n, cand, B, Nq, C_

    # -- a1: knnMatch(k=2) per (frame, candidate) ----------------------------------
    def knn2(self, q, q_n=None, cand=None):
        """-> idx i32[B,C,Nq,2], dist u16[B,C,Nq,2] (checkpoint_a_selftest.py:68)."""
        q, q_n, cand, B, Nq, C_ = self._prep(q, q_n, cand, self.n_keyframes)
        idx = np.empty((B, C_, Nq, 2), dtype=np.int32)
        dist = np.empty((B, C_, Nq, 2), dtype=np.uint16)
        if Nq == 0 or C_ == 0 or B == 0:
            return idx, dist
        self.ctx.check(_c.nclt_match_knn2(self.ctx.h, self.h, ptr(q), ptr(q_n), B, Nq, ptr(cand), C_,
                                          ptr(idx), ptr(dist)))
        return idx, dist

    # -- a1+a2: knn2 + Lowe ratio ---------------------------------------------------
    def ratio(self, q, q_n=None, cand=None, num=LOWE_NUM, den=LOWE_DEN):
        """-> pairs i32[B,C,Nq,2] (queryIdx, trainIdx), n i32[B,C] (selftest:68-71)."""
        q, q_n, cand, B, Nq, C_ = self._prep(q, q_n, cand, self.n_keyframes)
        pairs = np.full((B, C_, Nq, 2), -1, dtype=np.int32)
        n = np.zeros((B, C_), dtype=np.int32)
        if Nq == 0 or C_ == 0 or B == 0:
            return pairs, n
        self.ctx.check(_c.nclt_match_ratio(self.ctx.h, self.h, ptr(q), ptr(q_n), B, Nq, ptr(cand), C_,
                                           int(num), int(den), ptr(pairs), ptr(n)))
        return pairs, n

    # -- a3: crossCheck match(desc_t, desc_curr) ------------------------------------
    def cross(self, q, q_n=None, cand=None):
        """-> pairs i32[B,C,Nmax,2] (teach row, frame row), dist u16[B,C,Nmax], n i32[B,C]
        (visual_landmark_matcher.py:327)."""
        q, q_n, cand, B, Nq, C_ = self._prep(q, q_n, cand, self.n_keyframes)
        Nmax = max(self.max_rows, 1)
        pairs = np.full((B, C_, Nmax, 2), -1, dtype=np.int32)
        dist = np.zeros((B, C_, Nmax), dtype=np.uint16)
        n = np.zeros((B, C_), dtype=np.int32)
        if Nq == 0 or C_ == 0 or B == 0:
            return pairs, dist, n
        self.ctx.check(_c.nclt_match_cross(self.ctx.h, self.h, ptr(q), ptr(q_n), B, Nq, ptr(cand), C_, Nmax,
                                           ptr(pairs), ptr(dist), ptr(n)))
        return pairs, dist, n
