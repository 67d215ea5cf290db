"""Teach-time map builder on the GPU with the reference's call surface.

`TeachDepthMapper` mirrors scripts/common/teach_run_depth_mapper.py (class TeachDepthMapper:
__init__(out_prefix, origin_x, origin_y, width_m, height_m, res), cb(PointCloud2), save(), attrs
grid / frames_integrated / total_points_integrated / frames_skipped_empty) without rclpy: the
ROS node keeps its subscriptions, TF lookup and signal handling and forwards cb()/save() here
(INTEGRATION.md).  `depth_to_points` mirrors the relay's depth_cb
(scripts/common/tf_wall_clock_relay.py:868-903).  `integrate_depth` fuses both for batches of
frames (BASELINE.json config 3).
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import lib as _c, ptr, as_c

# log-odds constants of teach_run_depth_mapper.py:28-37 (kept for callers that import them)
L_FREE = -0.4
L_OCC = +1.4
L_MIN = -5.0
L_MAX = +5.0
THRESH_OCC = 0.65
THRESH_FREE = 0.25
FREE_L_TH = np.log(THRESH_FREE / (1 - THRESH_FREE))
OCC_L_TH = np.log(THRESH_OCC / (1 - THRESH_OCC))


def tf_to_matrix(tx, ty, tz, qx, qy, qz, qw):
    """map<-camera_link 4x4 from a TF (translation, quaternion): the expression order of
    teach_run_depth_mapper.py:64-80, float64."""
    x, y, z, w = qx, qy, qz, qw
    M = np.eye(4, dtype=np.float64)
    M[0, 0] = 1 - 2 * (y * y + z * z)
    M[0, 1] = 2 * (x * y - z * w)
    M[0, 2] = 2 * (x * z + y * w)
    M[1, 0] = 2 * (x * y + z * w)
    M[1, 1] = 1 - 2 * (x * x + z * z)
    M[1, 2] = 2 * (y * z - x * w)
    M[2, 0] = 2 * (x * z - y * w)
    M[2, 1] = 2 * (y * z + x * w)
    M[2, 2] = 1 - 2 * (x * x + y * y)
    M[0, 3] = tx
    M[1, 3] = ty
    M[2, 3] = tz
    return M


def depth_to_points(depth, fx=320.0, fy=320.0, cx=320.0, cy=240.0, ctx=None):
    """Relay depth_cb: depth f32[H,W] m (32FC1) or u16[H,W] mm (16UC1), or a batch [F,H,W] ->
    list of f32[N,3] clouds (z, -px, -py) in row-major pixel order; a single frame returns one array."""
    ctx = ctx or _lib.default_context()
    d = np.asarray(depth)
    single = d.ndim == 2
    if single:
        d = d[None]
    is_u16 = d.dtype == np.uint16
    d = np.ascontiguousarray(d if is_u16 else d.astype(np.float32, copy=False))
    F, H, W = d.shape
    cap = ((H + 3) // 4) * ((W + 3) // 4)
    out = np.zeros((F, cap, 3), dtype=np.float32)
    n = np.zeros(F, dtype=np.int32)
    ctx.check(_c.nclt_depth_to_points(ctx.h, ptr(d), int(is_u16), F, H, W, fx, fy, cx, cy, ptr(out), ptr(n), cap))
    clouds = [out[f, :n[f]].copy() for f in range(F)]
    return clouds[0] if single else clouds


class TeachDepthMapper:
    def __init__(self, out_prefix, origin_x=-110.0, origin_y=-50.0, width_m=200.0, height_m=60.0, res=0.1,
                 ctx=None):
        self.ctx = ctx or _lib.default_context()
        self.out_prefix = out_prefix
        self.res = res
        self.origin_x = origin_x
        self.origin_y = origin_y
        self.W = int(width_m / res)
        self.H = int(height_m / res)
        h = C.c_void_p()
        self.ctx.check(_c.nclt_occ_create(self.ctx.h, float(origin_x), float(origin_y), float(res), self.W, self.H,
                                          C.byref(h)))
        self.h = h
        self.frames_skipped_tf = 0

    # -- integration ----------------------------------------------------------------------
    def cb(self, points, tf):
        """One PointCloud2 worth of camera_link points f32[N,3] + the map<-camera_link TF as
        (tx,ty,tz,qx,qy,qz,qw) (what tf_buf.lookup_transform returned at mapper:128) or a 4x4."""
        pts = as_c(points, np.float32).reshape(-1, 3)
        T = self._matrix(tf)
        n = np.array([len(pts)], dtype=np.int32)
        if len(pts) == 0:
            pts = np.zeros((1, 3), dtype=np.float32)
        self.ctx.check(_c.nclt_occ_integrate_points(self.ctx.h, self.h, ptr(pts), ptr(n), 1, max(int(n[0]), 1), ptr(T)))

    def integrate_depth(self, depth, tfs, fx=320.0, fy=320.0, cx=320.0, cy=240.0):
        """depth [F,H,W] (f32 m or u16 mm), tfs: F transforms -> relay depth_cb + mapper cb, in order."""
        d = np.asarray(depth)
        if d.ndim == 2:
            d, tfs = d[None], [tfs]
        is_u16 = d.dtype == np.uint16
        d = np.ascontiguousarray(d if is_u16 else d.astype(np.float32, copy=False))
        T = np.ascontiguousarray(np.stack([self._matrix(t) for t in tfs]))
        F, H, W = d.shape
        self.ctx.check(_c.nclt_occ_integrate_depth(self.ctx.h, self.h, ptr(d), int(is_u16), F, H, W, ptr(T),
                                                   fx, fy, cx, cy))

    @staticmethod
    def _matrix(tf):
        a = np.asarray(tf, dtype=np.float64)
        if a.shape == (4, 4):
            return np.ascontiguousarray(a)
        return np.ascontiguousarray(tf_to_matrix(*[float(v) for v in a.reshape(7)]))

    # -- read-back ------------------------------------------------------------------------
    def _read(self, logodds=False, pgm=False, units=False):
        lo = np.zeros((self.H, self.W), dtype=np.float32) if logodds else None
        img = np.zeros((self.H, self.W), dtype=np.uint8) if pgm else None
        un = np.zeros((self.H, self.W), dtype=np.int32) if units else None
        cnt = np.zeros(3, dtype=np.int64)
        self.ctx.check(_c.nclt_occ_read(self.ctx.h, self.h, ptr(lo), ptr(img), ptr(un), ptr(cnt)))
        return lo, img, un, cnt

    @property
    def grid(self):
        """(H,W) float32 log-odds, like TeachDepthMapper.grid (= 0.2 * integer units)."""
        return self._read(logodds=True)[0]

    @property
    def units(self):
        return self._read(units=True)[2]

    @property
    def frames_integrated(self):
        return int(self._read()[3][0])

    @property
    def total_points_integrated(self):
        return int(self._read()[3][1])

    @property
    def frames_skipped_empty(self):
        return int(self._read()[3][2])

    def render(self):
        """uint8[H,W] PGM payload (thresholded, flipped)."""
        return self._read(pgm=True)[1]

    def save(self):
        """Writes out_prefix.pgm / .yaml exactly like teach_run_depth_mapper.py:208-233."""
        import yaml
        img = self.render()
        pgm_path = self.out_prefix + '.pgm'
        d = os.path.dirname(pgm_path)
        if d:
            os.makedirs(d, exist_ok=True)
        with open(pgm_path, 'wb') as f:
            f.write(b'P5\n')
            f.write(b'# exp 52 teach-run depth map\n')
            f.write(f'{self.W} {self.H}\n'.encode())
            f.write(b'255\n')
            f.write(img.tobytes())
        with open(self.out_prefix + '.yaml', 'w') as f:
            yaml.safe_dump({'image': pgm_path, 'resolution': self.res,
                            'origin': [self.origin_x, self.origin_y, 0.0], 'occupied_thresh': 0.65,
                            'free_thresh': 0.25, 'negate': 0}, f, default_flow_style=False)
        return pgm_path

    def reset(self):
        self.ctx.check(_c.nclt_occ_reset(self.ctx.h, self.h))

    def close(self):
        if getattr(self, 'h', None) is not None and self.h.value and self.ctx.h.value:
            _c.nclt_occ_destroy(self.ctx.h, self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def integrate_depth_device(mapper, depth_dev, T_dev, fx=320.0, fy=320.0, cx=320.0, cy=240.0):
    """Device-resident variant for replay workloads: depth_dev f32|u16 [F,H,W] and T_dev f64[F,4,4] are
    CUDA tensors already in HBM; enqueues on the mapper's context stream and returns."""
    import torch
    is_u16 = depth_dev.dtype in (torch.uint16, torch.int16)
    F, H, W = depth_dev.shape
    mapper.ctx.check(_c.nclt_occ_integrate_depth_dev(mapper.ctx.h, mapper.h, depth_dev.data_ptr(), int(is_u16), F, H, W,
                                                     T_dev.data_ptr(), fx, fy, cx, cy, None, None, 0))
