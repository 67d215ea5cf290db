"""CPU: the hit-count occupancy oracle against the reference's own build_occupancy (tests/golden/hitcount_golden.npz)."""
import os

import numpy as np

from oracle import hitcount as oh

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'hitcount_golden.npz')


def test_oracle_equals_reference_function():
    g = np.load(G)
    occ, x_min, z_min, nx, nz, fl, ob = oh.build_occupancy(g['points'], g['labels'], float(g['grid_res']),
                                                           int(g['min_total']), int(g['min_obstacle']))
    assert (x_min, z_min, nx, nz) == (float(g['x_min']), float(g['z_min']), int(g['nx']), int(g['nz']))
    assert np.array_equal(occ, g['occupancy'])
    assert int(fl.sum()) == int((g['labels'] == 0).sum()) and int(ob.sum()) == int((g['labels'] == 1).sum())
    assert (occ == 1).sum() > 100 and (occ == 0).sum() > 1000
