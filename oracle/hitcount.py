"""TEST INFRASTRUCTURE (never imported by the product) - CPU restatement of the hit-count occupancy grid of
datasets/rover/scripts/occupancy_astar.py:142-187 (`build_occupancy`): classified world points ->
floor / obstacle hit counts on the X-Z plane -> {-1 unknown, 0 free, 1 occupied}.  Pinned against the reference
function itself (tests/golden/hitcount_golden.npz, oracle/make_golden_ref.py::golden_hitcount imports the script with
its plotting imports stubbed)."""
import numpy as np

MIN_HITS_TOTAL = 3        # occupancy_astar.py:56
MIN_HITS_OBSTACLE = 5     # occupancy_astar.py:57
GRID_RES = 0.05           # occupancy_astar.py:36


def build_occupancy(points, labels, grid_res=GRID_RES, min_total=MIN_HITS_TOTAL, min_obstacle=MIN_HITS_OBSTACLE):
    """points f64[N,3] world, labels i8[N] (0 floor, 1 obstacle, -1 ignore) ->
    (occupancy i8[nz,nx], x_min, z_min, nx, nz, floor i32[nz,nx], obstacle i32[nz,nx])."""
    points = np.asarray(points, dtype=np.float64)
    labels = np.asarray(labels)
    cls = labels >= 0
    if not cls.any():
        raise ValueError('no classified points')                  # the reference would fail on min() of an empty array
    x, z = points[cls, 0], points[cls, 2]
    x_min, x_max = x.min() - 0.5, x.max() + 0.5                    # :155-158
    z_min, z_max = z.min() - 0.5, z.max() + 0.5
    nx = int((x_max - x_min) / grid_res) + 1                       # :160-161
    nz = int((z_max - z_min) / grid_res) + 1
    grids = []
    for lab in (0, 1):                                             # :167-175; np.add.at == a histogram of the cell ids
        m = labels == lab
        xi = np.clip(((points[m, 0] - x_min) / grid_res).astype(int), 0, nx - 1)
        zi = np.clip(((points[m, 2] - z_min) / grid_res).astype(int), 0, nz - 1)
        grids.append(np.bincount(zi * nx + xi, minlength=nz * nx).astype(np.int32).reshape(nz, nx))
    floor, obs = grids
    occ = np.full((nz, nx), -1, dtype=np.int8)                     # :177-181
    has = (floor + obs) >= min_total
    occ[has & (obs < min_obstacle)] = 0
    occ[has & (obs >= min_obstacle)] = 1
    return occ, x_min, z_min, nx, nz, floor, obs
