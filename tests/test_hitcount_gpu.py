"""GPU parity of the hit-count occupancy grid against the reference function's output and the oracle."""
import os

import numpy as np
import pytest

from oracle import hitcount as oh

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'hitcount_golden.npz')


def test_equals_reference_function(ctx):
    from nclt_slam_project_b200.hitcount import build_occupancy
    g = np.load(G)
    occ, x_min, z_min, nx, nz = build_occupancy(g['points'], g['labels'], float(g['grid_res']), ctx=ctx)
    assert (x_min, z_min, nx, nz) == (float(g['x_min']), float(g['z_min']), int(g['nx']), int(g['nz']))
    assert np.array_equal(occ, g['occupancy'])


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_random_clouds_equal_oracle(ctx, seed):
    from nclt_slam_project_b200.hitcount import build_occupancy
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 400000))
    pts = rng.normal(0, [6.0, 0.5, 9.0], (n, 3))
    pts[: n // 10] = np.round(pts[: n // 10] / 0.05) * 0.05          # points exactly on cell edges
    lab = rng.choice(np.array([-1, 0, 1], dtype=np.int8), n, p=[0.2, 0.6, 0.2])
    if seed == 2:
        lab[:] = -1
        lab[0] = 1                                                    # a single classified point
    res = float(rng.choice([0.05, 0.1, 0.37]))
    got = build_occupancy(pts, lab, res, return_counts=True, ctx=ctx)
    want = oh.build_occupancy(pts, lab, res)
    assert got[1:5] == want[1:5]
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[5], want[5]) and np.array_equal(got[6], want[6])


def test_errors(ctx):
    from nclt_slam_project_b200.hitcount import build_occupancy
    with pytest.raises(ValueError):
        build_occupancy(np.zeros((4, 3)), np.full(4, -1, dtype=np.int8), ctx=ctx)
