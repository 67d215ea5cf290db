"""CPU: the C oracle of the teach-map builder against golden vectors produced by the reference's
own modules (tf_wall_clock_relay.depth_cb -> teach_run_depth_mapper.cb/save, imported unmodified
under ROS stubs by oracle/make_golden_ref.py)."""
import os

import numpy as np
import pytest

from oracle import occupancy as oo

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'map_golden.npz')


@pytest.fixture(scope='module')
def gold():
    return np.load(G)


def _depth(g, f):
    return g['depth_u16'] if f in g['u16_frames'] else g['depth'][f]


def test_depth_cb_point_clouds_bit_exact(gold):
    for f in range(len(gold['cloud_n'])):
        pts = oo.depth_to_points(_depth(gold, f))
        n = int(gold['cloud_n'][f])
        assert len(pts) == n, f
        assert np.array_equal(pts.view(np.uint32), gold['cloud'][f, :n].view(np.uint32)), f


@pytest.mark.parametrize('cls', [oo.OracleMapper, oo.OracleMapperInt])
def test_mapper_grid_and_pgm(gold, cls):
    ox, oy, wm, hm, res = gold['cfg'].tolist()
    m = cls(ox, oy, wm, hm, res)
    snaps = {int(f): i for i, f in enumerate(gold['grid_snap_frames'])}
    for f in range(len(gold['cloud_n'])):
        m.cb(oo.depth_to_points(_depth(gold, f)), tuple(gold['tf'][f]))
        if f in snaps:
            ref = gold['grid_snaps'][snaps[f]]
            if cls is oo.OracleMapper:
                assert np.array_equal(m.grid, ref), f          # float32 log-odds, bit-exact
            else:
                assert np.abs(m.logodds() - ref).max() < 1e-5, f
    assert m.frames_integrated == int(gold['frames_integrated'])
    assert m.total_points_integrated == int(gold['total_points'])
    assert m.frames_skipped_empty == int(gold['skipped_empty'])
    assert oo.pgm_bytes(m.render()) == gold['pgm'].tobytes()       # every occupancy cell, byte for byte


def test_integer_model_never_changes_a_class():
    """float32 reference semantics vs the exact-integer model on random update sequences."""
    rng = np.random.default_rng(0)
    for _ in range(3000):
        n = int(rng.integers(1, 400))
        occ = rng.random(n) < rng.uniform(0.05, 0.95)
        g = np.float32(0.0)
        u = 0
        for o in occ:
            if o:
                v = g + np.float32(1.4)
                g = v if v < 5.0 else np.float32(5.0)
                u = min(u + 7, 25)
            else:
                v = g + np.float32(-0.4)
                g = v if v > -5.0 else np.float32(-5.0)
                u = max(u - 2, -25)
        cls_f = 0 if g > np.log(0.65 / 0.35) else (254 if g < np.log(0.25 / 0.75) else 205)
        cls_i = 0 if u >= 4 else (254 if u <= -6 else 205)
        assert cls_f == cls_i
        assert abs(float(g) - 0.2 * u) < 2e-5


def test_reference_language_restatement(gold):
    """oracle/occupancy_ref.py (NumPy + the per-cell Python loop, what bench.py's reference arm times for the map
    workload) against the grids the reference's own modules produced: float32 log-odds bit for bit."""
    from oracle import occupancy_ref as orf
    ox, oy, wm, hm, res = gold['cfg'].tolist()
    m = orf.PyMapper(ox, oy, wm, hm, res)
    snaps = {int(f): i for i, f in enumerate(gold['grid_snap_frames'])}
    last = sorted(snaps)[min(1, len(snaps) - 1)]
    for f in range(last + 1):
        cloud = orf.depth_to_cloud(_depth(gold, f))
        n = int(gold['cloud_n'][f])
        assert np.array_equal(cloud.view(np.uint32), gold['cloud'][f, :n].view(np.uint32)), f
        m.cb(cloud, oo.tf_to_matrix(*gold['tf'][f]))
        if f in snaps:
            assert np.array_equal(m.grid, gold['grid_snaps'][snaps[f]]), f
