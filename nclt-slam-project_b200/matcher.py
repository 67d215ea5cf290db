"""Host-side mirror of the repeat-time matcher node (scripts/common/visual_landmark_matcher.py)
without rclpy: candidate selection (a9), the GPU candidate loop (a1-a7), pose composition (a8),
consistency gate, covariance and the anchor_matches.csv log.

The ROS node keeps its subscriptions, ORB extraction, pose-file reader and publisher; its
`_tick` body from the candidate selection to the CSV line is what `LandmarkMatcher.tick`
reproduces.  Module constants keep the reference's names because
checkpoint_a_selftest.py:31-36 imports them by name.
"""
import math
import os
import pickle

import numpy as np

from ._lib import LocalizeParams, PnpParams
from .library import LandmarkLibrary
from .pipeline import localize_batch, MODE_CROSSCHECK, MODE_RATIO

# --- constants of visual_landmark_matcher.py:46-89 ------------------------------------------
FX, FY = 320.0, 320.0
CX, CY = 320.0, 240.0
K = np.array([[FX, 0, CX], [0, FY, CY], [0, 0, 1]], dtype=np.float32)
DIST = np.zeros((4, 1), dtype=np.float32)
CANDIDATE_RADIUS_M = 8.0
MAX_CANDIDATES = 5
HEADING_TOL_DEG = 90.0
MIN_MATCHES = 10
LOWE_RATIO = 0.80
REPROJ_MAX_PX = 2.0
RANSAC_REPROJ_PX = 3.0
RANSAC_ITERATIONS = 200
MIN_INLIERS = 10
CONSISTENCY_M = 5.0
TICK_HZ = 2.0
# continuous landmark accumulation (visual_landmark_matcher.py:85-89; shipped enabled)
ACCUM_ENABLE = True
ACCUM_SILENCE_S = 5.0
ACCUM_MIN_DIST_M = 5.0
ACCUM_MIN_KPTS = 30
ACCUM_SAVE_PKL = True
# exp 63 global relocalisation (experiments/63_global_reloc/scripts/visual_landmark_matcher.py:82-86)
RELOC_AGE_S = 20.0
RELOC_DRIFT_M = 3.0
RELOC_MAX_CANDIDATES = 25
RELOC_MIN_INLIERS = 18
RELOC_REPROJ_MAX_PX = 1.5

CSV_HEADER = ('ts,vio_x,vio_y,candidates_tried,best_n_inliers,'
              'best_reproj_err,anchor_x,anchor_y,outcome\n')


def quat_to_rot(qx, qy, qz, qw):
    """Unit quaternion -> rotation matrix (same element formulas as matcher:115-127)."""
    xx, yy, zz = qx * qx, qy * qy, qz * qz
    xy, xz, yz = qx * qy, qx * qz, qy * qz
    wx, wy, wz = qw * qx, qw * qy, qw * qz
    return np.array([[1 - 2 * (yy + zz), 2 * (xy - wz), 2 * (xz + wy)],
                     [2 * (xy + wz), 1 - 2 * (xx + zz), 2 * (yz - wx)],
                     [2 * (xz - wy), 2 * (yz + wx), 1 - 2 * (xx + yy)]], dtype=np.float64)


def rot_to_quat(R):
    """Rotation matrix -> (qx,qy,qz,qw), branch on the trace / largest diagonal like matcher:129-157."""
    m00, m11, m22 = R[0, 0], R[1, 1], R[2, 2]
    tr = m00 + m11 + m22
    if tr > 0:
        s = 0.5 / math.sqrt(tr + 1.0)
        return ((R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s, 0.25 / s)
    if m00 > m11 and m00 > m22:
        s = 2.0 * math.sqrt(1.0 + m00 - m11 - m22)
        return (0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s, (R[2, 1] - R[1, 2]) / s)
    if m11 > m22:
        s = 2.0 * math.sqrt(1.0 + m11 - m00 - m22)
        return ((R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s, (R[0, 2] - R[2, 0]) / s)
    s = 2.0 * math.sqrt(1.0 + m22 - m00 - m11)
    return ((R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s, (R[1, 0] - R[0, 1]) / s)


def cam_world_to_base_world(cam_pose, base_to_cam_t, base_to_cam_R):
    """Camera world pose -> base_link world pose (matcher:160-172)."""
    R_wc = quat_to_rot(*cam_pose[3:7])
    R_wb = R_wc @ np.asarray(base_to_cam_R).T
    t_wb = np.array(cam_pose[:3], dtype=np.float64) - R_wb @ np.asarray(base_to_cam_t)
    q = rot_to_quat(R_wb)
    return (float(t_wb[0]), float(t_wb[1]), float(t_wb[2]), q[0], q[1], q[2], q[3])


def rodrigues(rvec):
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    th = math.sqrt(float(r @ r))
    if th < np.finfo(np.float64).eps:
        return np.eye(3)
    k = r / th
    kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return math.cos(th) * np.eye(3) + (1 - math.cos(th)) * np.outer(k, k) + math.sin(th) * kx


def compose_anchor(teach_pose, rvec, tvec, base_to_cam_t, base_to_cam_R):
    """PnP result (teach camera in the current camera frame) -> current base_link pose in the
    teach map (matcher:361-378)."""
    R_teach_cur = rodrigues(rvec).T
    t_teach_cur = -R_teach_cur @ np.asarray(tvec, dtype=np.float64).reshape(3)
    R_world_teach = quat_to_rot(*teach_pose[3:7])
    R_world_cur = R_world_teach @ R_teach_cur
    t_world_cur = np.array(teach_pose[:3], dtype=np.float64) + R_world_teach @ t_teach_cur
    q = rot_to_quat(R_world_cur)
    cam_pose_world = (float(t_world_cur[0]), float(t_world_cur[1]), float(t_world_cur[2]), q[0], q[1], q[2], q[3])
    return cam_world_to_base_world(cam_pose_world, base_to_cam_t, base_to_cam_R)


def anchor_std(n_inliers):
    """Inlier count -> anchor std (matcher:400-405)."""
    if n_inliers >= 25:
        return 0.05
    if n_inliers >= 15:
        return 0.05 + 0.15 * (25 - n_inliers) / 10.0
    return 0.2


def anchor_covariance(std):
    cov = [0.0] * 36
    cov[0] = std * std
    cov[7] = std * std
    cov[14] = 0.25
    cov[21] = 0.05
    cov[28] = 0.05
    cov[35] = 0.05
    return cov


class LandmarkMatcher:
    """`mode='crosscheck'` is the production node (matcher:318-380); `mode='ratio'` is the offline
    validator's k=2 + Lowe variant (checkpoint_a_selftest.py:62-103, no heading gate, 5 nearest)."""

    def __init__(self, pkl_path_or_dict, log_csv=None, mode='crosscheck', ctx=None):
        self.pkl_path = None
        if isinstance(pkl_path_or_dict, (str, os.PathLike)):
            self.pkl_path = str(pkl_path_or_dict)
            with open(pkl_path_or_dict, 'rb') as f:
                data = pickle.load(f)
        else:
            data = pkl_path_or_dict
        self.pkl_data = data
        self.landmarks = data['landmarks']
        self.base_to_cam_t = np.array(data['base_to_cam_translation'])
        self.base_to_cam_R = np.array(data['base_to_cam_rot'])
        self.xy = np.array([[lm['pose'][0], lm['pose'][1]] for lm in self.landmarks])
        self.heading = np.array([self._lm_heading_rad(lm) for lm in self.landmarks])
        self.mode = mode
        self.library = LandmarkLibrary.from_pkl_dict(data, ctx=ctx)
        self.params = LocalizeParams(
            mode=MODE_CROSSCHECK if mode == 'crosscheck' else MODE_RATIO, ratio_num=4, ratio_den=5,
            min_matches=MIN_MATCHES, min_inliers=MIN_INLIERS, reproj_max_px=REPROJ_MAX_PX,
            pnp=PnpParams(FX, FY, CX, CY, RANSAC_ITERATIONS, RANSAC_REPROJ_PX, 0.99, 1))
        self.n_attempts = 0
        self.n_published = 0
        self.last_anchor_ts = 0.0
        self.n_initial_landmarks = len(self.landmarks)
        self.n_accumulated = 0
        self.log_csv = log_csv
        if log_csv:
            os.makedirs(os.path.dirname(log_csv), exist_ok=True)
            with open(log_csv, 'w') as f:
                f.write(CSV_HEADER)

    def _lm_heading_rad(self, lm):
        R_wb = quat_to_rot(*lm['pose'][3:7]) @ self.base_to_cam_R.T
        fwd = R_wb @ np.array([1.0, 0.0, 0.0])
        return math.atan2(fwd[1], fwd[0])

    @staticmethod
    def _heading_of(base_pose):
        fwd = quat_to_rot(*base_pose[3:7]) @ np.array([1.0, 0.0, 0.0])
        return math.atan2(fwd[1], fwd[0])

    def select_candidates(self, base_pose):
        """a9: <= MAX_CANDIDATES landmarks by VIO distance (matcher:293-302 / selftest:54-57)."""
        d = np.linalg.norm(self.xy - np.array(base_pose[:2]), axis=1)
        order = np.argsort(d)
        if self.mode == 'ratio':
            return [int(i) for i in order[:MAX_CANDIDATES] if d[i] < CANDIDATE_RADIUS_M]
        cur = self._heading_of(base_pose)
        dh = self.heading - cur
        hdg_err = np.abs(np.arctan2(np.sin(dh), np.cos(dh)))
        tol = math.radians(HEADING_TOL_DEG)
        cand = [int(i) for i in order[:MAX_CANDIDATES * 3] if d[i] < CANDIDATE_RADIUS_M and hdg_err[i] < tol]
        return cand[:MAX_CANDIDATES]

    def _log(self, ts, vio_xy, n_tried, n_in, err, anchor_xy, outcome):
        if not self.log_csv:
            return
        ax = anchor_xy[0] if anchor_xy else ''
        ay = anchor_xy[1] if anchor_xy else ''
        with open(self.log_csv, 'a') as f:
            f.write(f'{ts:.3f},{vio_xy[0]:.3f},{vio_xy[1]:.3f},{n_tried},{n_in},{err},{ax},{ay},{outcome}\n')

    def accumulate(self, base_pose, desc_curr, pts_curr_2d, depth_mm, ts):
        """`_maybe_accumulate` (visual_landmark_matcher.py:434-500): after ACCUM_SILENCE_S without an anchor and with no
        landmark within ACCUM_MIN_DIST_M, the current frame becomes a NEW landmark at the current VIO pose - appended to
        `landmarks`, to the xy / heading indices AND to the device library (nclt_lib_append), so the very next tick can
        match against it.  Same arithmetic as the node: np.round to the pixel, float32 millimetres / 1000, float64
        back-projection rounded to float32, camera pose through the static base -> camera offset.  Returns True when a
        landmark was added."""
        if not ACCUM_ENABLE:
            return False
        if ts - self.last_anchor_ts < ACCUM_SILENCE_S:
            return False
        vio_xy = (base_pose[0], base_pose[1])
        d = np.linalg.norm(self.xy - np.array(vio_xy), axis=1)
        if d.min() < ACCUM_MIN_DIST_M:
            return False
        if depth_mm is None or pts_curr_2d is None or len(pts_curr_2d) == 0:
            return False
        pts = np.asarray(pts_curr_2d)
        desc_curr = np.asarray(desc_curr)
        uu = np.round(pts[:, 0]).astype(np.int32)
        vv = np.round(pts[:, 1]).astype(np.int32)
        H, W = depth_mm.shape
        valid = (uu >= 1) & (uu < W - 1) & (vv >= 1) & (vv < H - 1)
        uu, vv = uu[valid], vv[valid]
        pts2, desc2 = pts[valid], desc_curr[valid]
        if len(uu) == 0:
            return False
        d_c = depth_mm[vv, uu].astype(np.float32) / 1000.0
        ok = (d_c > 0.5) & (d_c < 15.0)
        if ok.sum() < ACCUM_MIN_KPTS:
            return False
        uu, vv = uu[ok], vv[ok]
        kpts2d_kept, desc_kept, d_c = pts2[ok], desc2[ok], d_c[ok]
        x_cam = (uu - CX) * d_c / FX
        y_cam = (vv - CY) * d_c / FY
        kpts_3d_cam = np.stack([x_cam, y_cam, d_c], axis=-1).astype(np.float32)
        R_wb = quat_to_rot(*base_pose[3:7])
        cam_xyz = np.array([base_pose[0], base_pose[1], base_pose[2]]) + R_wb @ self.base_to_cam_t
        R_wc = R_wb @ self.base_to_cam_R
        from scipy.spatial.transform import Rotation as SR      # the node's own conversion (matcher:478-479)
        qx, qy, qz, qw = SR.from_matrix(R_wc).as_quat()
        new_lm = {
            'pose': (float(cam_xyz[0]), float(cam_xyz[1]), float(cam_xyz[2]), float(qx), float(qy), float(qz), float(qw)),
            'descriptors': desc_kept, 'keypoints_2d': kpts2d_kept, 'keypoints_3d_cam': kpts_3d_cam,
            'ts': ts, 'n_features': int(len(kpts_3d_cam)), 'accumulated': True,
        }
        self.landmarks.append(new_lm)
        self.xy = np.vstack([self.xy, [vio_xy[0], vio_xy[1]]])
        self.heading = np.append(self.heading, self._lm_heading_rad(new_lm))
        self.library.append(desc_kept, kpts_3d_cam)            # the device library grows with the host one
        self.n_accumulated += 1
        return True

    def save_augmented(self, path=None):
        """The node's SIGTERM handler (matcher:192-202): with accumulated landmarks, pickle the whole dict (same schema)
        next to the teach pickle as `<name>_augmented.pkl`.  Returns the path written, or None when there is nothing to
        save."""
        if not (ACCUM_SAVE_PKL and self.n_accumulated > 0):
            return None
        if path is None:
            if self.pkl_path is None:
                raise ValueError('save_augmented: the matcher was built from a dict - pass a path')
            path = self.pkl_path.replace('.pkl', '_augmented.pkl')
        self.pkl_data['landmarks'] = self.landmarks
        with open(path, 'wb') as f:
            pickle.dump(self.pkl_data, f)
        return path

    def reloc_candidates(self, base_pose, desc_curr):
        """exp 63 global relocalisation pool (63_global_reloc/.../visual_landmark_matcher.py:328-345): every
        heading-compatible landmark, scored by its crossCheck match count against the current frame (one batched GPU
        call instead of a cv2 call per landmark), >= MIN_MATCHES, sorted by (count, index) descending, top 25."""
        cur = self._heading_of(base_pose)
        dh = self.heading - cur
        hdg_err = np.abs(np.arctan2(np.sin(dh), np.cos(dh)))
        pool = [int(i) for i in np.where(hdg_err < math.radians(HEADING_TOL_DEG))[0]
                if self.library.counts[i] >= MIN_MATCHES]
        if not pool:
            return []
        if getattr(self.library.ctx, 'engine', 0) == 2 and 4 * len(pool) >= self.library.n_keyframes:
            # a large pool on the fp4 tensor engine: crossCheck against the WHOLE library in two tcgen05 passes
            # (nclt_match_cross with cand == NULL) and read the pool's counts - identical counts, keyframes are independent
            _, _, n_all = self.library.cross(np.asarray(desc_curr)[None], None, None)
            counts = n_all[0, pool]
        else:
            _, _, n = self.library.cross(np.asarray(desc_curr)[None], None, np.asarray(pool, dtype=np.int32)[None])
            counts = n[0]
        scored = [(int(c), li) for c, li in zip(counts, pool) if c >= MIN_MATCHES]
        scored.sort(reverse=True)
        return [li for _, li in scored[:RELOC_MAX_CANDIDATES]]

    def tick_image(self, bgr, base_pose, ts=0.0, drift_est=0.0):
        """The tick from the camera image onwards (matcher:305-310): BGR -> gray -> ORB(500) on the GPU
        (orb.py, bit-identical to cv2's detectAndCompute), `pts_curr_2d = [k.pt ...]`, then `tick`."""
        if getattr(self, 'orb', None) is None:
            from .orb import ORB
            self.orb = ORB(nfeatures=500, ctx=self.library.ctx)                      # matcher:207
        kp, desc, n = self.orb.detect_and_compute_batch(np.asarray(bgr)[None])
        m = int(n[0])
        if m == 0:
            return self.tick(None, None, base_pose, ts, drift_est)
        return self.tick(desc[0, :m], kp[0, :m, :2], base_pose, ts, drift_est)

    def tick(self, desc_curr, pts_curr_2d, base_pose, ts=0.0, drift_est=0.0, depth_mm=None):
        """One matcher tick from the ORB output onwards. Returns a dict with 'outcome' (the CSV
        outcome string), and on publish 'anchor_pose', 'std', 'covariance', 'n_inliers',
        'reproj_err', 'lm_idx'.  drift_est: tf_relay's SLAM-vs-encoder disagreement (/tmp/drift_est.txt in exp 63);
        with the default 0 the tick is the production node's, otherwise exp 63's kidnapped-robot fallback can fire.
        depth_mm: the aligned depth image (u16 millimetres, `self.last_depth` of the node); when given, the three
        outcomes after which the node calls `_maybe_accumulate` (no_candidates, no_pnp_accept, consistency_fail) do the
        same here and report it as res['accumulated']."""
        self.n_attempts += 1
        vio_xy = (base_pose[0], base_pose[1])
        cand_idx = self.select_candidates(base_pose)
        if desc_curr is None or len(desc_curr) < MIN_MATCHES:
            self._log(ts, vio_xy, len(cand_idx), 0, '', None, 'curr_no_features')
            return {'outcome': 'curr_no_features', 'candidates': cand_idx}
        relocating = False
        if (not cand_idx and self.mode != 'ratio' and (ts - self.last_anchor_ts) > RELOC_AGE_S
                and drift_est > RELOC_DRIFT_M):
            cand_idx = self.reloc_candidates(base_pose, desc_curr)
            relocating = True
        if not cand_idx:
            self._log(ts, vio_xy, 0, 0, '', None, 'no_candidates')
            acc = depth_mm is not None and self.accumulate(base_pose, desc_curr, pts_curr_2d, depth_mm, ts)
            return {'outcome': 'no_candidates', 'candidates': cand_idx, 'accumulated': acc}
        cand = np.full((1, max(MAX_CANDIDATES, len(cand_idx))), -1, dtype=np.int32)
        cand[0, :len(cand_idx)] = cand_idx
        params = self.params
        if relocating:
            params = LocalizeParams(mode=self.params.mode, min_matches=MIN_MATCHES, min_inliers=RELOC_MIN_INLIERS,
                                    reproj_max_px=RELOC_REPROJ_MAX_PX, pnp=self.params.pnp)
        out = localize_batch(self.library, np.asarray(desc_curr)[None], np.asarray(pts_curr_2d, dtype=np.float32)[None],
                             None, cand, params)
        slot = int(out['best_cand'][0])
        if slot < 0:
            self._log(ts, vio_xy, len(cand_idx), 0, '', None, 'no_pnp_accept')
            acc = depth_mm is not None and self.accumulate(base_pose, desc_curr, pts_curr_2d, depth_mm, ts)
            return {'outcome': 'no_pnp_accept', 'candidates': cand_idx, 'accumulated': acc}
        lm_idx = cand_idx[slot]
        n_inliers = int(out['n_inliers'][0])
        reproj_err = float(out['reproj'][0])
        anchor_pose = compose_anchor(self.landmarks[lm_idx]['pose'], out['rvec'][0], out['tvec'][0],
                                     self.base_to_cam_t, self.base_to_cam_R)
        consistency_d = math.hypot(anchor_pose[0] - vio_xy[0], anchor_pose[1] - vio_xy[1])
        res = {'candidates': cand_idx, 'anchor_pose': anchor_pose, 'n_inliers': n_inliers, 'reproj_err': reproj_err,
               'lm_idx': lm_idx, 'shift': consistency_d, 'relocating': relocating}
        if not relocating and consistency_d > CONSISTENCY_M:     # the jump is the point of a relocalisation (exp 63)
            res['outcome'] = f'consistency_fail_{consistency_d:.1f}m'
            self._log(ts, vio_xy, len(cand_idx), n_inliers, f'{reproj_err:.2f}', anchor_pose[:2], res['outcome'])
            res['accumulated'] = depth_mm is not None and self.accumulate(base_pose, desc_curr, pts_curr_2d, depth_mm, ts)
            return res
        std = anchor_std(n_inliers)
        res.update(std=std, covariance=anchor_covariance(std),
                   outcome=f'published_std{std:.2f}_shift{consistency_d:.1f}')
        self.n_published += 1
        self.last_anchor_ts = ts
        self._log(ts, vio_xy, len(cand_idx), n_inliers, f'{reproj_err:.2f}', anchor_pose[:2], res['outcome'])
        return res
