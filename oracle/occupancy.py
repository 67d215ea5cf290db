"""TEST INFRASTRUCTURE - ctypes front end of oracle/occupancy.c (the teach-map builder oracle).

Restates tf_wall_clock_relay.py:868-887 (depth_cb) and teach_run_depth_mapper.py:125-216
(cb, _bresenham_mark, save).  `OracleMapper` keeps the reference's float32 log-odds grid;
`OracleMapperInt` is the exact-integer model the CUDA path uses - both must render the same PGM.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import ctypes as C
import math

import numpy as np

from .pnp import _L   # liboracle.so (built on demand)

_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_L.orc_depth_to_points.argtypes = [_vp, _i, _i, _i, _d, _d, _d, _d, _vp]
_L.orc_depth16_to_points.argtypes = [_vp, _i, _i, _i, _d, _d, _d, _d, _vp]
_L.orc_mapper_integrate.argtypes = [_vp, _i, _i, _d, _d, _d, _vp, _vp, _i, _vp, _vp, _vp]
_L.orc_mapper_integrate_int.argtypes = [_vp, _i, _i, _d, _d, _d, _vp, _vp, _i, _vp, _vp, _vp]
_L.orc_mapper_rays.argtypes = [_vp, _vp, _i, _i, _i, _d, _d, _d, _vp, _vp, _vp]
_L.orc_mapper_render.argtypes = [_vp, _i, _i, _vp]
_L.orc_mapper_render_int.argtypes = [_vp, _i, _i, _vp]


def _p(a):
    return a.ctypes.data_as(_vp)


def depth_to_points(depth, fx=320.0, fy=320.0, cx=320.0, cy=240.0, step=4):
    """relay depth_cb: depth f32[H,W] metres (or u16 millimetres) -> points f32[N,3] (FLU)."""
    depth = np.ascontiguousarray(depth)
    H, W = depth.shape
    cap = ((H + step - 1) // step) * ((W + step - 1) // step)
    pts = np.zeros((cap, 3), dtype=np.float32)
    if depth.dtype == np.uint16:
        n = _L.orc_depth16_to_points(_p(depth), H, W, step, fx, fy, cx, cy, _p(pts))
    else:
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        n = _L.orc_depth_to_points(_p(depth), H, W, step, fx, fy, cx, cy, _p(pts))
    return pts[:n].copy()


def tf_to_matrix(tx, ty, tz, x, y, z, w):
    """teach_run_depth_mapper.py:64-80, same expression order (float64)."""
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


class OracleMapper:
    """TeachDepthMapper without ROS: grid (H,W) float32 log-odds, cb(points, tf), render()."""
    integer = False

    def __init__(self, origin_x=-110.0, origin_y=-45.0, width_m=195.0, height_m=90.0, res=0.1):
        self.res = res
        self.origin_x = origin_x
        self.origin_y = origin_y
        self.W = int(width_m / res)
        self.H = int(height_m / res)
        self.grid = np.zeros((self.H, self.W), dtype=np.int32 if self.integer else np.float32)
        self.counters = np.zeros(3, dtype=np.int64)   # integrated, total points, skipped_empty

    frames_integrated = property(lambda s: int(s.counters[0]))
    total_points_integrated = property(lambda s: int(s.counters[1]))
    frames_skipped_empty = property(lambda s: int(s.counters[2]))

    def cb(self, points, tf):
        """points f32[N,3] in camera_link, tf = (tx,ty,tz,qx,qy,qz,qw) map->camera_link."""
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        T = np.ascontiguousarray(tf_to_matrix(*tf))
        rays = np.zeros((len(pts) // 4 + 2, 2), dtype=np.int32)
        inb = np.zeros(len(pts) // 4 + 2, dtype=np.uint8)
        fn = _L.orc_mapper_integrate_int if self.integer else _L.orc_mapper_integrate
        return fn(_p(self.grid), self.H, self.W, self.origin_x, self.origin_y, self.res, _p(T), _p(pts), len(pts),
                  _p(self.counters), _p(rays), _p(inb))

    def render(self):
        """The PGM payload of save(): uint8[H,W], already flipped."""
        img = np.zeros((self.H, self.W), dtype=np.uint8)
        (_L.orc_mapper_render_int if self.integer else _L.orc_mapper_render)(_p(self.grid), self.H, self.W, _p(img))
        return img

    def logodds(self):
        return self.grid.astype(np.float32) * np.float32(0.2) if self.integer else self.grid


class OracleMapperInt(OracleMapper):
    integer = True


def pgm_bytes(img):
    """teach_run_depth_mapper.py:218-223 file layout."""
    h, w = img.shape
    return b'P5\n# exp 52 teach-run depth map\n' + f'{w} {h}\n'.encode() + b'255\n' + img.tobytes()


def rays_for(points, tf, H, W, ox, oy, res):
    """Debug: the ordered (r1,c1) endpoints, in-bounds flags and sensor cell of one cloud."""
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
    T = np.ascontiguousarray(tf_to_matrix(*tf))
    rays = np.zeros((len(pts) // 4 + 2, 2), dtype=np.int32)
    inb = np.zeros(len(pts) // 4 + 2, dtype=np.uint8)
    r0c0 = np.zeros(2, dtype=np.int32)
    m = _L.orc_mapper_rays(_p(T), _p(pts), len(pts), H, W, ox, oy, res, _p(r0c0), _p(rays), _p(inb))
    return m, r0c0, rays[:max(m, 0)], inb[:max(m, 0)]


def yaw_tf(x, y, yaw, z=0.48, cam_dx=0.5):
    """Planar base pose composed with the static base_link->camera_link offset (0.5, 0, 0.48)
    (tf_wall_clock_relay.py:63-69)."""
    return (x + math.cos(yaw) * cam_dx, y + math.sin(yaw) * cam_dx, z, 0.0, 0.0, math.sin(yaw / 2), math.cos(yaw / 2))
