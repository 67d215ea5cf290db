"""CPU, build container only: the UNMODIFIED reference node (visual_landmark_matcher.py, imported from /root/reference
under ROS stubs) driven through the INTEGRATION.md patch - `cv2.BFMatcher`, `cv2.solvePnPRansac`, `cv2.projectPoints`
swapped for `nclt_slam_project_b200.cv2_compat` - must log and publish exactly what it does on OpenCV
(tests/golden/tick_golden.npz).

/root/reference cannot travel to the GPU box and this container has no GPU, so the node and the kernels never meet
in one process.  What this test pins is the shim SURFACE the node sees (DMatch objects, argument / return shapes and
dtypes, the cv2.error contract), with the device calls behind the shims answered by the CPU oracle (test doubles
patched in below - the product modules contain no such switch).  That the device calls return what the oracle
returns is the job of the `-m gpu` parity tests (tests/test_match_gpu.py, test_pnp_gpu.py, test_matcher_gpu.py)."""
import os
import pickle

import numpy as np
import pytest

REF = '/root/reference/simulation/isaac/scripts/common/visual_landmark_matcher.py'
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason='/root/reference is only present in the build container')
GD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class _OracleLibrary:
    """Stands in for LandmarkLibrary inside cv2_compat.BFMatcher: one keyframe, answers with the NumPy oracle."""

    def __init__(self, descriptors, points3d=None, ctx=None):
        self.kf = [np.asarray(d) for d in descriptors]

    def knn2(self, q, q_n=None, cand=None):
        from oracle import hamming as oh
        idx, dist = oh.knn2(q[0], self.kf[0])
        return idx[None, None].astype(np.int32), dist[None, None].astype(np.uint16)

    def cross(self, q, q_n=None, cand=None):
        from oracle import hamming as oh
        qi, ti, d = oh.cross_check(self.kf[0], q[0])        # library rows are cv2's query set
        n = len(qi)
        pairs = np.full((1, 1, max(len(self.kf[0]), 1), 2), -1, dtype=np.int32)
        dist = np.zeros((1, 1, max(len(self.kf[0]), 1)), dtype=np.uint16)
        pairs[0, 0, :n, 0], pairs[0, 0, :n, 1], dist[0, 0, :n] = qi, ti, d
        return pairs, dist, np.array([[n]], dtype=np.int32)

    def close(self):
        pass


def _oracle_pnp_batch(obj, img, n=None, params=None, ctx=None, debug=False):
    from oracle import pnp as op
    o = op.pnp_ransac(np.ascontiguousarray(obj[0]), np.ascontiguousarray(img[0]), params.iterations, params.reproj_error)
    N = obj.shape[1]
    mask = np.zeros((1, N), dtype=np.uint8)
    if o['ok']:
        mask[0, o['inliers']] = 1
    return {'ok': np.array([1 if o['ok'] else 0], dtype=np.uint8), 'rvec': np.asarray(o['rvec'], dtype=np.float64).reshape(1, 3),
            'tvec': np.asarray(o['tvec'], dtype=np.float64).reshape(1, 3), 'mask': mask,
            'n_inliers': np.array([int(mask.sum())], dtype=np.int32)}


def _oracle_project(obj, rvec, tvec, fx=320.0, fy=320.0, cx=320.0, cy=240.0, ctx=None):
    from oracle import pnp as op
    obj = np.ascontiguousarray(obj, dtype=np.float32).reshape(-1, 3)
    return op.reproj_err(obj, np.zeros((len(obj), 2), dtype=np.float32), rvec, tvec, fx, fy, cx, cy)[1]


def test_unmodified_node_runs_on_the_shims(monkeypatch, tmp_path):
    import cv2
    from oracle import ros_stubs
    import nclt_slam_project_b200  # noqa: F401
    from nclt_slam_project_b200 import synth, cv2_compat as g2, _lib, pnp as gp
    # device calls behind the shims -> CPU oracle (test doubles; nothing of this exists in the product modules)
    monkeypatch.setattr(_lib, 'default_context', lambda device=0: object())
    monkeypatch.setattr(g2, 'LandmarkLibrary', _OracleLibrary)
    monkeypatch.setattr(gp, 'pnp_ransac_batch', _oracle_pnp_batch)
    monkeypatch.setattr(gp, 'project_points', _oracle_project)

    vm = ros_stubs.import_reference()['visual_landmark_matcher']
    monkeypatch.setattr(vm, 'ACCUM_ENABLE', False)          # as in the golden run

    # ---- the INTEGRATION.md patch: three call sites (+ the exception class the node catches at :328) ----------------
    class _Cv2Patched:
        BFMatcher = staticmethod(g2.BFMatcher)
        solvePnPRansac = staticmethod(g2.solvePnPRansac)
        projectPoints = staticmethod(g2.projectPoints)
        NORM_HAMMING = g2.NORM_HAMMING
        SOLVEPNP_ITERATIVE = g2.SOLVEPNP_ITERATIVE
        error = (cv2.error, g2.error)

        def __getattr__(self, name):                        # everything else (cvtColor, Rodrigues, ORB_create) stays OpenCV
            return getattr(cv2, name)

    monkeypatch.setattr(vm, 'cv2', _Cv2Patched())

    g = np.load(os.path.join(GD, 'tick_golden.npz'))
    data = synth.make_library(int(g['lib_seed']), n_kf=40, n_desc=300, ragged=True, route_len_m=80.0)
    pkl = str(tmp_path / 'south_landmarks.pkl')
    with open(pkl, 'wb') as f:
        pickle.dump(data, f)
    csv = str(tmp_path / 'log' / 'anchor_matches.csv')
    node = vm.VisualLandmarkMatcher(pkl, csv)
    assert isinstance(node.matcher, g2.BFMatcher) and node.matcher.crossCheck      # the node built OUR matcher
    node.last_rgb = np.zeros((480, 640, 3), dtype=np.uint8)
    node.last_depth = np.full((480, 640), 2000, dtype=np.uint16)

    class _Kp:
        def __init__(self, pt):
            self.pt = (float(pt[0]), float(pt[1]))

    class _Orb:
        next = None

        def detectAndCompute(self, gray, mask):
            d, p = self.next
            return [_Kp(q) for q in p], d

    node.orb = _Orb()
    for i in range(len(g['kinds'])):
        n = int(g['n_desc'][i])
        node.orb.next = (g['desc'][i, :n], g['pts2d'][i, :n])
        node._read_pose = (lambda bp=tuple(g['base_pose'][i]): bp)
        n_before = len(node.anchor_pub.sent)
        node._tick()
        got = open(csv).read().strip().split('\n')[-1].split(',')[1:]
        ref = str(g['csv'][i]).split(',')
        assert got[-1] == ref[-1], (i, got, ref)                # outcome string incl. std / shift
        assert got[:5] == ref[:5], (i, got, ref)                # vio, candidates tried, inliers, reprojection error
        published = len(node.anchor_pub.sent) > n_before
        assert published == bool(g['published'][i])
        if published:
            m = node.anchor_pub.sent[-1]
            p, o = m.pose.pose.position, m.pose.pose.orientation
            a = np.array([p.x, p.y, p.z, o.x, o.y, o.z, o.w])
            assert np.abs(a[:3] - g['anchor'][i, :3]).max() < 1e-3 and np.abs(a[3:] - g['anchor'][i, 3:]).max() < 1e-4
            assert np.allclose(list(m.pose.covariance), g['cov'][i], rtol=0, atol=1e-12)
    assert node.n_published == int(g['published'].sum()) > 0
