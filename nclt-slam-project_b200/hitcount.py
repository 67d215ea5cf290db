"""Hit-count occupancy grid on the GPU with the reference's call surface:
`build_occupancy(points, labels, grid_res)` mirrors datasets/rover/scripts/occupancy_astar.py:142-187 (same
arguments, same returns) with the np.add.at scatter and the thresholding as CUDA kernels."""
import numpy as np

from . import _lib
from ._lib import lib as _c, ptr, as_c

MIN_HITS_TOTAL = 3        # occupancy_astar.py:56
MIN_HITS_OBSTACLE = 5     # occupancy_astar.py:57
GRID_RES = 0.05           # occupancy_astar.py:36


def build_occupancy(points, labels, grid_res=GRID_RES, min_total=MIN_HITS_TOTAL, min_obstacle=MIN_HITS_OBSTACLE,
                    return_counts=False, ctx=None):
    """points f64[N,3], labels i8[N] -> (occupancy i8[nz,nx], x_min, z_min, nx, nz) [+ floor, obstacle counts]."""
    ctx = ctx or _lib.default_context()
    pts = as_c(points, np.float64).reshape(-1, 3)
    lab = as_c(labels, np.int8).reshape(-1)
    if len(pts) != len(lab) or len(pts) == 0:
        raise ValueError('points / labels mismatch or empty')
    cls = lab >= 0
    if not cls.any():
        raise ValueError('no classified points')
    # capacity from a host-side bound of the extent (cheap: two min/max passes NumPy does at memory speed)
    ext_x = float(pts[cls, 0].max() - pts[cls, 0].min()) + 1.0
    ext_z = float(pts[cls, 2].max() - pts[cls, 2].min()) + 1.0
    cap = (int(ext_x / grid_res) + 2) * (int(ext_z / grid_res) + 2)
    origin = np.zeros(2, dtype=np.float64)
    dims = np.zeros(2, dtype=np.int32)
    occ = np.zeros(cap, dtype=np.int8)
    fl = np.zeros(cap, dtype=np.int32) if return_counts else None
    ob = np.zeros(cap, dtype=np.int32) if return_counts else None
    ctx.check(_c.nclt_hitcount_occupancy(ctx.h, ptr(pts), ptr(lab), len(pts), float(grid_res), int(min_total),
                                         int(min_obstacle), cap, ptr(origin), ptr(dims), ptr(occ), ptr(fl), ptr(ob)))
    nx, nz = int(dims[0]), int(dims[1])
    out = (occ[:nx * nz].reshape(nz, nx).copy(), float(origin[0]), float(origin[1]), nx, nz)
    if return_counts:
        out += (fl[:nx * nz].reshape(nz, nx).copy(), ob[:nx * nz].reshape(nz, nx).copy())
    return out
