"""Debug: mixed fast / fallback frames in one call vs the sequential oracle; prints where the grids differ."""
import sys
import numpy as np
sys.path.insert(0, '/root/repo')
import nclt_slam_project_b200  # noqa
from nclt_slam_project_b200.mapper import TeachDepthMapper
from nclt_slam_project_b200._lib import lib as _c, ptr
from oracle import occupancy as oo

cfg = (-110.0, -45.0, 195.0, 90.0, 0.1)


def integ(m, clouds, tfs):
    nmax = max(max(len(c) for c in clouds), 1)
    pts = np.zeros((len(clouds), nmax, 3), dtype=np.float32)
    n = np.array([len(c) for c in clouds], dtype=np.int32)
    for i, c in enumerate(clouds):
        pts[i, :len(c)] = c
    T = np.ascontiguousarray(np.stack([oo.tf_to_matrix(*t) for t in tfs]))
    m.ctx.check(_c.nclt_occ_integrate_points(m.ctx.h, m.h, ptr(pts), ptr(n), len(clouds), nmax, ptr(T)))


def run(tag, clouds, tfs):
    m = TeachDepthMapper('/tmp/unused', *cfg)
    ref = oo.OracleMapperInt(*cfg)
    for c, t in zip(clouds, tfs):
        ref.cb(c, t)
    integ(m, clouds, tfs)
    g = m.units
    bad = np.argwhere(g != ref.grid)
    print(tag, 'frames', len(clouds), 'mismatching cells', len(bad), 'touched', int((ref.grid != 0).sum()), flush=True)
    for (r, c) in bad[:12]:
        print('   cell', r, c, 'gpu', g[r, c], 'ref', ref.grid[r, c])
    if len(bad):
        print('   rows', bad[:, 0].min(), bad[:, 0].max(), 'cols', bad[:, 1].min(), bad[:, 1].max())


rng = np.random.default_rng(14)


def cloud(n, reach, side):
    return np.stack([rng.uniform(0.5, reach, n), rng.uniform(-side, side, n), rng.uniform(-0.2, 1.0, n)], axis=-1).astype(np.float32)


small = [cloud(900, 6, 2) for _ in range(10)]
big = cloud(6000, 60, 30)
tf = lambda i: oo.yaw_tf(-20.0 + 0.05 * i, 0.0, 0.2)
run('fast only x1', small[:1], [tf(0)])
run('fast only x5', small[:5], [tf(i) for i in range(5)])
run('fallback only', [big], [tf(5)])
run('fast, fallback', small[:1] + [big], [tf(0), tf(1)])
run('fallback, fast', [big] + small[:1], [tf(0), tf(1)])
run('5 fast, fallback, 5 fast', small[:5] + [big] + small[5:], [tf(i) for i in range(11)])
