"""TEST INFRASTRUCTURE - the teach-map path restated in the reference's OWN language and cost structure.

The reference's map builder is NumPy + a pure-Python Bresenham loop (tf_wall_clock_relay.py:868-887 depth_cb,
teach_run_depth_mapper.py:125-195 cb / _bresenham_mark).  /root/reference cannot travel to the GPU box, so
`bench.py --impl reference --workload map` and the map workload's cpu_baseline time THIS restatement: the same
NumPy calls and the same per-cell Python loop with the float32 grid (what the reference costs on a host core),
checked against the reference-produced golden map (tests/test_oracle_occupancy.py).  oracle/occupancy.c is the
fast C checker used by the parity tests; this file exists for honest CPU timing only.

Only tests/ and bench.py's cpu_baseline / reference legs may import this.
"""
import numpy as np

L_FREE, L_OCC, L_MIN, L_MAX = -0.4, 1.4, -5.0, 5.0      # teach_run_depth_mapper.py:28-33


def depth_to_cloud(depth, fx=320.0, fy=320.0, cx=320.0, cy=240.0, step=4):
    """relay depth_cb (tf_wall_clock_relay.py:868-887): f32 metres or u16 millimetres -> f32[N,3] camera_link."""
    if depth.dtype == np.uint16:
        depth = depth.astype(np.float32) / 1000.0
    h, w = depth.shape
    vv, uu = np.meshgrid(np.arange(0, h, step), np.arange(0, w, step), indexing='ij')
    z = depth[vv, uu]
    ok = (z > 0.3) & (z < 10.0) & np.isfinite(z)
    z = z[ok]
    uf = uu[ok].astype(np.float32)
    vf = vv[ok].astype(np.float32)
    px = (uf - cx) / fx * z
    py = (vf - cy) / fy * z
    return np.stack([z, -px, -py], axis=-1).astype(np.float32)


class PyMapper:
    """TeachDepthMapper.cb / _bresenham_mark without ROS (float32 grid, Python loop per cell)."""

    def __init__(self, origin_x=-110.0, origin_y=-45.0, width_m=195.0, height_m=90.0, res=0.1):
        self.res, self.origin_x, self.origin_y = res, origin_x, origin_y
        self.W, self.H = int(width_m / res), int(height_m / res)
        self.grid = np.zeros((self.H, self.W), dtype=np.float32)
        self.frames_integrated = 0
        self.total_points_integrated = 0
        self.frames_skipped_empty = 0

    def _pix(self, x, y):                                     # mapper:120-123, truncation toward zero
        return int((y - self.origin_y) / self.res), int((x - self.origin_x) / self.res)

    def cb(self, pts_cam, T):
        """pts_cam f32[N,3], T f64[4,4] map<-camera_link (mapper:125-170)."""
        if len(pts_cam) == 0:
            self.frames_skipped_empty += 1
            return
        homog = np.column_stack([pts_cam, np.ones(len(pts_cam))])
        in_map = (T @ homog.T).T[:, :3]
        zz = in_map[:, 2]
        in_map = in_map[(zz > 0.2) & (zz < 2.0)]
        if len(in_map) == 0:
            return
        in_map = in_map[::4]
        r0, c0 = self._pix(T[0, 3], T[1, 3])
        if not (0 <= r0 < self.H and 0 <= c0 < self.W):
            return
        for (x, y, _) in in_map:
            r1, c1 = self._pix(x, y)
            if 0 <= r1 < self.H and 0 <= c1 < self.W:
                self._ray(r0, c0, r1, c1)
        self.frames_integrated += 1
        self.total_points_integrated += len(in_map)

    def _ray(self, r0, c0, r1, c1):                          # mapper:172-195
        g = self.grid
        dr, dc = abs(r1 - r0), abs(c1 - c0)
        sr = 1 if r0 < r1 else -1
        sc = 1 if c0 < c1 else -1
        err = dr - dc
        r, c = r0, c0
        while True:
            if r == r1 and c == c1:
                g[r, c] = min(L_MAX, g[r, c] + L_OCC)
                return
            g[r, c] = max(L_MIN, g[r, c] + L_FREE)
            e2 = 2 * err
            if e2 > -dc:
                err -= dc
                r += sr
            if e2 < dr:
                err += dr
                c += sc
