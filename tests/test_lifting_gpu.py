"""GPU parity of the teach-time keypoint lifting: the records LandmarkRecorder builds through the
CUDA kernel must equal the records the reference recorder node appended (tests/golden/lift_golden.npz), bit for bit,
and the kernel must equal the oracle on adversarial inputs."""
import os
import pickle

import numpy as np
import pytest

from oracle import lifting as ol

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'lift_golden.npz')


def test_recorder_equals_the_reference_node(ctx, tmp_path):
    from nclt_slam_project_b200.recorder import LandmarkRecorder, write_packed, read_packed
    from nclt_slam_project_b200.library import LandmarkLibrary
    g = np.load(G)
    rec = LandmarkRecorder(str(tmp_path / 'out' / 'landmarks.pkl'), min_disp_m=2.0, ctx=ctx)
    n_rec = 0
    for i in range(len(g['kinds'])):
        n = int(g['n_kpts'][i])
        r = rec.tick(g['kpts'][i, :n], g['desc'][i, :n], g['depth'][i], tuple(g['base_pose'][i]), float(100.0 + i))
        assert (r is not None) == bool(g['recorded'][i]), (i, str(g['kinds'][i]))
        if r is None:
            continue
        n_rec += 1
        m = int(g['n_feat'][i])
        assert r['n_features'] == m and r['ts'] == g['ts'][i]
        assert np.array_equal(np.array(r['pose']), g['cam_pose'][i])
        assert np.array_equal(r['descriptors'], g['rec_desc'][i, :m])
        assert np.array_equal(r['keypoints_2d'].view(np.uint32), g['rec_kp2d'][i, :m].view(np.uint32))
        assert np.array_equal(r['keypoints_3d_cam'].view(np.uint32), g['rec_kp3d'][i, :m].view(np.uint32))
    assert n_rec >= 4
    # the pickle has the reference's schema and loads into the matcher's library; so does the packed file
    path = rec.save()
    d = pickle.load(open(path, 'rb'))
    assert sorted(d.keys()) == list(g['pkl_keys'])
    assert [d['intrinsics'][k] for k in ('fx', 'fy', 'cx', 'cy', 'width', 'height')] == g['pkl_intrinsics'].tolist()
    assert np.array_equal(np.array(d['base_to_cam_rot']), g['pkl_b2c_R'])
    lib_a = LandmarkLibrary.from_pkl(path, ctx=ctx)
    packed = rec.save_packed(str(tmp_path / 'landmarks.nclt'))
    lib_b = LandmarkLibrary.from_packed(packed, ctx=ctx)
    assert np.array_equal(lib_a.offsets, lib_b.offsets)
    q = g['desc'][0, :200][None]
    ia, da = lib_a.knn2(q)
    ib, db = lib_b.knn2(q)
    assert np.array_equal(ia, ib) and np.array_equal(da, db)
    p = read_packed(packed)
    assert np.array_equal(p['poses'], np.array([lm['pose'] for lm in rec.landmarks]))
    assert np.array_equal(np.asarray(p['keypoints_2d']), np.concatenate([lm['keypoints_2d'] for lm in rec.landmarks]))
    lib_a.close()
    lib_b.close()


def test_kernel_equals_oracle_on_adversarial_frames(ctx):
    """Batched call: depth discontinuities exactly at the std threshold, holes leaving 2 / 3 valid neighbours, depth at
    the range limits, keypoints on the borders and on exact .5 pixel coordinates, ragged keypoint counts."""
    from nclt_slam_project_b200.recorder import lift_keypoints
    rng = np.random.default_rng(77)
    F, Nmax, H, W = 5, 700, 480, 640
    depth = rng.integers(300, 16000, (F, H, W)).astype(np.uint16)
    depth[0] = (2000 + rng.integers(-450, 450, (H, W))).astype(np.uint16)        # std around 0.26-0.30 m
    depth[1][rng.random((H, W)) < 0.7] = 0                                         # mostly holes
    depth[2] = rng.choice(np.array([499, 500, 501, 14999, 15000, 15001, 10, 11], dtype=np.uint16), (H, W))
    depth[3] = 3000
    depth[3][::2] += 300                                                            # rows alternate 3.0 / 3.3 m
    kp = np.stack([rng.uniform(-2, W + 2, (F, Nmax)), rng.uniform(170, H + 2, (F, Nmax))], axis=2).astype(np.float32)
    kp[:, :60, 0] = np.round(kp[:, :60, 0]) + 0.5
    kp[:, 60:120, 1] = np.round(kp[:, 60:120, 1]) + 0.5
    kp[:, 120, :] = (0.49, 200.0)
    kp[:, 121, :] = (638.5, 478.5)
    kp[:, 122, :] = (320.0, 180.4)
    kp[:, 123, :] = (320.0, 180.6)
    n = np.array([700, 0, 333, 1, 700], dtype=np.int32)
    res = lift_keypoints(kp, depth, n_kpts=n, ctx=ctx)
    total = 0
    for f in range(F):
        keep, pts = ol.lift_keypoints(kp[f, :n[f]], depth[f])
        assert np.array_equal(res[f][0], keep), f
        assert np.array_equal(res[f][1].view(np.uint32), pts.view(np.uint32)), f
        total += len(keep)
    assert total > 300
