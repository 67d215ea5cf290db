"""CPU: the keypoint-lifting oracle (oracle/lifting.py) against the records the REFERENCE recorder
node produced (tests/golden/lift_golden.npz: VisualLandmarkRecorder._tick run unmodified under ROS stubs) and its
float32 std restatement against NumPy."""
import os

import numpy as np
import pytest

from oracle import lifting as ol

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'lift_golden.npz')


@pytest.fixture(scope='module')
def gold():
    return np.load(G)


def test_std_restatement_is_numpy_bit_for_bit():
    rng = np.random.default_rng(1)
    for t in range(4000):
        n = int(rng.integers(3, 10))
        a = rng.integers(11, 16000, n).astype(np.float32) / np.float32(1000.0)
        if t % 2:
            a = (a[0] + rng.integers(-40, 40, n).astype(np.float32) / np.float32(1000.0)).astype(np.float32)
        assert a.std().view(np.uint32) == np.float32(ol.np_std_f32(list(a))).view(np.uint32)


def test_records_equal_the_reference_node(gold):
    kinds = [str(k) for k in gold['kinds']]
    assert kinds.count('ok') >= 4 and 'far' in kinds and 'few' in kinds
    for i, kind in enumerate(kinds):
        n = int(gold['n_kpts'][i])
        if kind == 'near':                   # displacement gate (recorder:236): the node never looked at the frame
            assert not gold['recorded'][i]
            continue
        rec = ol.make_record(gold['kpts'][i, :n], gold['desc'][i, :n], gold['depth'][i], gold['cam_pose'][i], gold['ts'][i])
        assert (rec is not None) == bool(gold['recorded'][i]), (i, kind)
        if rec is None:
            continue
        m = int(gold['n_feat'][i])
        assert rec['n_features'] == m
        assert np.array_equal(rec['descriptors'], gold['rec_desc'][i, :m])
        assert np.array_equal(rec['keypoints_2d'].view(np.uint32), gold['rec_kp2d'][i, :m].view(np.uint32))
        assert np.array_equal(rec['keypoints_3d_cam'].view(np.uint32), gold['rec_kp3d'][i, :m].view(np.uint32))


def test_pickle_schema_constants(gold):
    assert list(gold['pkl_keys']) == ['base_to_cam_rot', 'base_to_cam_translation', 'intrinsics', 'landmarks']
    assert gold['pkl_intrinsics'].tolist() == [ol.FX, ol.FY, ol.CX, ol.CY, ol.W, ol.H]
