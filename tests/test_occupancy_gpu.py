"""GPU parity of the teach-map builder: point clouds, every
occupancy cell of the PGM and the counters must equal what the reference modules produced
(tests/golden/map_golden.npz), and the integer grid must equal the sequential oracle cell by cell."""
import os

import numpy as np
import pytest

from oracle import occupancy as oo

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'map_golden.npz')


@pytest.fixture(scope='module')
def gold():
    return np.load(G)


def _depth(g, f):
    return g['depth_u16'] if f in g['u16_frames'] else g['depth'][f]


def test_depth_cb_clouds_bit_exact(ctx, gold):
    from nclt_slam_project_b200.mapper import depth_to_points
    f32_frames = [f for f in range(len(gold['cloud_n'])) if f not in gold['u16_frames']]
    clouds = depth_to_points(gold['depth'][f32_frames])
    for f, c in zip(f32_frames, clouds):
        n = int(gold['cloud_n'][f])
        assert len(c) == n, f
        assert np.array_equal(c.view(np.uint32), gold['cloud'][f, :n].view(np.uint32)), f
    f = int(gold['u16_frames'][0])
    c = depth_to_points(gold['depth_u16'])
    assert np.array_equal(c.view(np.uint32), gold['cloud'][f, :gold['cloud_n'][f]].view(np.uint32))


@pytest.mark.parametrize('mode', ['points', 'depth', 'depth_batched'])
def test_reference_golden_map(ctx, gold, mode, tmp_path):
    from nclt_slam_project_b200.mapper import TeachDepthMapper
    ox, oy, wm, hm, res = gold['cfg'].tolist()
    m = TeachDepthMapper(str(tmp_path / 'teach_map'), ox, oy, wm, hm, res)
    ref = oo.OracleMapperInt(ox, oy, wm, hm, res)
    F = len(gold['cloud_n'])
    snaps = {int(f): i for i, f in enumerate(gold['grid_snap_frames'])}
    if mode == 'depth_batched':
        u16 = int(gold['u16_frames'][0])
        m.integrate_depth(gold['depth'][:u16], gold['tf'][:u16])
        m.integrate_depth(gold['depth_u16'][None], gold['tf'][u16:u16 + 1])
        m.integrate_depth(gold['depth'][u16 + 1:], gold['tf'][u16 + 1:])
    for f in range(F):
        cloud = gold['cloud'][f, :gold['cloud_n'][f]]
        ref.cb(cloud, tuple(gold['tf'][f]))
        if mode == 'points':
            m.cb(cloud, tuple(gold['tf'][f]))
        elif mode == 'depth':
            m.integrate_depth(_depth(gold, f), tuple(gold['tf'][f]))
        if mode != 'depth_batched' and f in snaps:
            assert np.array_equal(m.units, ref.grid), f                      # exact, cell by cell
            assert np.abs(m.grid - gold['grid_snaps'][snaps[f]]).max() < 1e-5   # vs the reference's float32 grid
    assert np.array_equal(m.units, ref.grid)
    assert m.frames_integrated == int(gold['frames_integrated'])
    assert m.total_points_integrated == int(gold['total_points'])
    assert m.frames_skipped_empty == int(gold['skipped_empty'])
    path = m.save()
    assert open(path, 'rb').read() == gold['pgm'].tobytes()                  # teach_map.pgm byte for byte
    yml = open(str(tmp_path / 'teach_map.yaml'), 'rb').read().replace(str(tmp_path).encode(), b'<TMP>')
    assert yml == gold['yaml'].tobytes()


def test_full_size_grid_against_sequential_oracle(ctx):
    """run_teach.sh:29 geometry (1950 x 900 @ 0.1 m), 40 frames along a path, depth mode, one batch."""
    from nclt_slam_project_b200 import synth
    from nclt_slam_project_b200.mapper import TeachDepthMapper
    cfg = (-110.0, -45.0, 195.0, 90.0, 0.1)
    poses = synth.boustrophedon_path(4000, step_m=0.05)[::100]
    depth = np.stack([synth.make_depth_frame(500 + i, p, cyl_density=0.08) for i, p in enumerate(poses)])
    tfs = [synth.camera_link_transform(*p) for p in poses]
    m = TeachDepthMapper('/tmp/unused', *cfg)
    m.integrate_depth(depth, tfs)
    ref = oo.OracleMapperInt(*cfg)
    fref = oo.OracleMapper(*cfg)
    for d, t in zip(depth, tfs):
        pts = oo.depth_to_points(d)
        ref.cb(pts, t)
        fref.cb(pts, t)
    assert np.array_equal(m.units, ref.grid)
    assert np.array_equal(m.render(), fref.render())          # = the float32 reference semantics
    assert (m.units != 0).sum() > 5000
    assert m.frames_integrated == ref.frames_integrated and m.total_points_integrated == ref.total_points_integrated


def test_order_dependence_is_preserved(ctx):
    """Saturated cells + mixed pass/hit sequences: the result depends on ray order (SURVEY hard
    parts); integrate the same frame many times so that clamps are active everywhere."""
    from nclt_slam_project_b200 import synth
    from nclt_slam_project_b200.mapper import TeachDepthMapper
    cfg = (-12.0, -9.0, 40.0, 30.0, 0.1)
    pose = (0.0, 0.0, 0.4)
    d = synth.make_depth_frame(9, pose, cyl_density=0.15)
    t = synth.camera_link_transform(*pose)
    m = TeachDepthMapper('/tmp/unused', *cfg)
    ref = oo.OracleMapperInt(*cfg)
    pts = oo.depth_to_points(d)
    for _ in range(12):
        ref.cb(pts, t)
    m.integrate_depth(np.stack([d] * 12), [t] * 12)
    assert np.array_equal(m.units, ref.grid)
    assert (ref.grid == 25).any() and (ref.grid == -25).any()


def test_fallback_path_large_cloud(ctx):
    """A cloud whose rays span more than the shared-memory window -> exact sequential fallback."""
    from nclt_slam_project_b200.mapper import TeachDepthMapper
    cfg = (-110.0, -45.0, 195.0, 90.0, 0.1)
    rng = np.random.default_rng(4)
    pts = np.stack([rng.uniform(0.5, 60, 6000), rng.uniform(-30, 30, 6000), rng.uniform(-0.2, 1.0, 6000)],
                   axis=-1).astype(np.float32)
    t = oo.yaw_tf(-20.0, 0.0, 0.2)
    m = TeachDepthMapper('/tmp/unused', *cfg)
    ref = oo.OracleMapperInt(*cfg)
    for _ in range(2):
        m.cb(pts, t)
        ref.cb(pts, t)
    assert np.array_equal(m.units, ref.grid)
    assert m.total_points_integrated == ref.total_points_integrated


@pytest.mark.parametrize('H,W', [(480, 640), (478, 640), (37, 50)])
def test_host_path_equals_device_path_and_oracle(ctx, H, W):
    """The host-pointer entry copies only the sampled rows (strided 2-D copy) when H % 4 == 0 and whole frames
    otherwise; both must give the grid of the device-resident path and of the sequential oracle."""
    import torch
    from nclt_slam_project_b200.mapper import TeachDepthMapper, integrate_depth_device, tf_to_matrix
    rng = np.random.default_rng(H * 1000 + W)
    F = 5
    depth = rng.uniform(0.2, 9.0, (F, H, W)).astype(np.float32)
    depth[rng.random((F, H, W)) < 0.05] = np.nan
    tfs = [oo.yaw_tf(1.0 + 0.3 * f, -0.5 * f, 0.4 * f) for f in range(F)]
    T = np.stack([tf_to_matrix(*t) for t in tfs])
    cfg = (-20.0, -20.0, 40.0, 40.0, 0.1)
    kw = dict(fx=W / 2.0, fy=W / 2.0, cx=W / 2.0, cy=H / 2.0)
    a = TeachDepthMapper('/tmp/nclt_t_host', *cfg, ctx=ctx)
    a.integrate_depth(depth, T, **kw)
    b = TeachDepthMapper('/tmp/nclt_t_dev', *cfg, ctx=ctx)
    integrate_depth_device(b, torch.from_numpy(depth).cuda(), torch.from_numpy(T).cuda(), **kw)
    ctx.sync()
    ref = oo.OracleMapperInt(*cfg)
    for f in range(F):
        ref.cb(oo.depth_to_points(depth[f], **kw), tfs[f])
    ua, ub = a.units, b.units
    assert np.array_equal(ua, ub)
    assert np.array_equal(ua, ref.grid)
    assert a.total_points_integrated == ref.total_points_integrated
    a.close()
    b.close()


def _integrate_clouds(m, clouds, tfs):
    """Several clouds in ONE call (nclt_occ_integrate_points: pts f32[F,Nmax,3], n i32[F], T f64[F,16])."""
    from nclt_slam_project_b200._lib import lib as _c, ptr
    nmax = max(max(len(c) for c in clouds), 1)
    pts = np.zeros((len(clouds), nmax, 3), dtype=np.float32)
    n = np.array([len(c) for c in clouds], dtype=np.int32)
    for i, c in enumerate(clouds):
        pts[i, :len(c)] = c
    T = np.ascontiguousarray(np.stack([oo.tf_to_matrix(*t) for t in tfs]))
    m.ctx.check(_c.nclt_occ_integrate_points(m.ctx.h, m.h, ptr(pts), ptr(n), len(clouds), nmax, ptr(T)))


def test_fallback_frame_between_fast_frames_keeps_the_order(ctx):
    """One call: small clouds (fast path), a cloud spanning more than the shared-memory window (exact ray-by-ray
    path inside the tile gather), small clouds again - all over the same saturating cells."""
    from nclt_slam_project_b200.mapper import TeachDepthMapper
    cfg = (-110.0, -45.0, 195.0, 90.0, 0.1)
    rng = np.random.default_rng(14)

    def cloud(n, reach, side):
        return np.stack([rng.uniform(0.5, reach, n), rng.uniform(-side, side, n), rng.uniform(-0.2, 1.0, n)], axis=-1).astype(np.float32)

    near = cloud(900, 6, 2)
    clouds = [near] * 5 + [cloud(6000, 60, 30)] + [near] * 3 + [cloud(900, 6, 2) for _ in range(2)]      # repeated hits saturate
    tfs = [oo.yaw_tf(-20.0 + (0.05 * i if i > 8 else 0.0), 0.0, 0.2) for i in range(len(clouds))]
    m = TeachDepthMapper('/tmp/unused', *cfg)
    ref = oo.OracleMapperInt(*cfg)
    for c, t in zip(clouds, tfs):
        ref.cb(c, t)
    _integrate_clouds(m, clouds, tfs)
    assert np.array_equal(m.units, ref.grid)
    assert (ref.grid == 25).any() and (ref.grid == -25).any()
    assert m.frames_integrated == ref.frames_integrated and m.total_points_integrated == ref.total_points_integrated


def test_many_mixed_cells_take_several_bitmap_passes(ctx):
    """Hits strung along a few bearings: every endpoint cell is also crossed by all the longer rays of its bearing, so
    hundreds of cells see both passes and hits (more than one 256-cell bitmap pass); order matters in each."""
    from nclt_slam_project_b200.mapper import TeachDepthMapper
    cfg = (-30.0, -30.0, 60.0, 60.0, 0.1)
    rng = np.random.default_rng(15)
    pts = []
    for d in np.arange(0.6, 19.0, 0.1):
        for yaw in (-0.5, -0.2, 0.0, 0.3, 0.6):
            pts.append((d * np.cos(yaw), d * np.sin(yaw), 0.5))
    pts = np.array(pts, dtype=np.float32)
    pts = np.repeat(pts, 4, axis=0)                     # the mapper keeps every 4th point
    rng.shuffle(pts.reshape(-1, 4, 3))                  # ray order decides the interleaving of passes and hits
    t = oo.yaw_tf(0.0, 0.0, 0.1)
    m = TeachDepthMapper('/tmp/unused', *cfg)
    ref = oo.OracleMapperInt(*cfg)
    for _ in range(3):
        ref.cb(pts, t)
    _integrate_clouds(m, [pts] * 3, [t] * 3)
    assert np.array_equal(m.units, ref.grid)


def test_more_frames_than_one_chunk(ctx):
    """2100 small frames in one call (stage A / stage B pairs of 2048 frames): grid and counters as the oracle."""
    from nclt_slam_project_b200.mapper import TeachDepthMapper
    cfg = (-12.0, -9.0, 40.0, 30.0, 0.1)
    rng = np.random.default_rng(16)
    clouds, tfs = [], []
    for i in range(2100):
        n = int(rng.integers(0, 60))
        clouds.append(np.stack([rng.uniform(0.5, 5, n), rng.uniform(-2, 2, n), rng.uniform(-0.3, 1.0, n)], axis=-1).astype(np.float32))
        tfs.append(oo.yaw_tf(-8.0 + 0.006 * i, 0.002 * i, 0.001 * i))
    m = TeachDepthMapper('/tmp/unused', *cfg)
    ref = oo.OracleMapperInt(*cfg)
    for c, t in zip(clouds, tfs):
        ref.cb(c, t)
    _integrate_clouds(m, clouds, tfs)
    assert np.array_equal(m.units, ref.grid)
    assert (m.frames_integrated, m.total_points_integrated, m.frames_skipped_empty) == \
        (ref.frames_integrated, ref.total_points_integrated, ref.frames_skipped_empty)
