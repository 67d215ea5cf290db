"""Teach-time landmark recorder on the GPU with the reference's call surface.

`LandmarkRecorder` mirrors scripts/common/visual_landmark_recorder.py (class VisualLandmarkRecorder:
__init__(out_pkl, min_disp_m), the body of _tick from the displacement gate to the appended record,
_save()) without rclpy / cv2: the ROS node keeps its subscriptions, ORB and pose file and forwards
`tick(kpts_xy, desc, depth_mm, base_pose, ts)` here.  The per-keypoint Python loop
(recorder:247-291) runs as one CUDA kernel (`nclt_lift_keypoints`); module constants keep the
reference's names.  `save()` writes the reference's pickle; `save_packed()` writes the same library as
one flat binary that `LandmarkLibrary.from_packed` maps straight into the device layout.
"""
import math
import os
import pickle

import numpy as np

from . import _lib
from ._lib import LiftParams, lib as _c, ptr, as_c
from .matcher import quat_to_rot, rot_to_quat

# visual_landmark_recorder.py:53-72, 82-90
FX, FY = 320.0, 320.0
CX, CY = 320.0, 240.0
W, H = 640, 480
DEPTH_MIN_M = 0.5
DEPTH_MAX_M = 15.0
DEPTH_VAR_MAX_M = 0.30
GROUND_Y_THRESHOLD = 180
MIN_POINTS = 30
BASE_TO_CAM_TRANSLATION = np.array([0.35, 0.0, 0.18])
BASE_TO_CAM_ROT = np.array([[0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]])

PACKED_MAGIC = b'NCLTLIB1'


def base_to_cam_world(base_x, base_y, base_z, base_qx, base_qy, base_qz, base_qw):
    """base_link world pose -> camera optical frame world pose (recorder:137-151)."""
    R_world_base = quat_to_rot(base_qx, base_qy, base_qz, base_qw)
    cam_pos_world = np.array([base_x, base_y, base_z]) + R_world_base @ BASE_TO_CAM_TRANSLATION
    R_world_cam = R_world_base @ BASE_TO_CAM_ROT
    qx, qy, qz, qw = rot_to_quat(R_world_cam)
    return (float(cam_pos_world[0]), float(cam_pos_world[1]), float(cam_pos_world[2]), qx, qy, qz, qw)


def lift_keypoints(kpts_xy, depth_mm, n_kpts=None, params=None, ctx=None):
    """kpts_xy f32[N,2] + depth_mm u16[H,W]  ->  (keep i32[M], pts3d f32[M,3]);
    batched: kpts_xy f32[F,Nmax,2], depth_mm u16[F,H,W], n_kpts i32[F] -> lists of those."""
    ctx = ctx or _lib.default_context()
    prm = params or LiftParams()
    k = as_c(kpts_xy, np.float32)
    d = as_c(depth_mm, np.uint16)
    single = d.ndim == 2
    if single:
        k = k.reshape(1, -1, 2)
        d = d[None]
    F, Nmax = k.shape[0], max(k.shape[1], 1)
    if k.shape[1] == 0:
        k = np.zeros((F, 1, 2), dtype=np.float32)
    n = np.full(F, k.shape[1] if kpts_xy is not None and np.asarray(kpts_xy).size else 0, dtype=np.int32) \
        if n_kpts is None else as_c(n_kpts, np.int32).reshape(F)
    keep = np.zeros((F, Nmax), dtype=np.int32)
    pts = np.zeros((F, Nmax, 3), dtype=np.float32)
    out_n = np.zeros(F, dtype=np.int32)
    ctx.check(_c.nclt_lift_keypoints(ctx.h, ptr(d), F, d.shape[1], d.shape[2], ptr(k), ptr(n), Nmax, prm, ptr(keep),
                                     ptr(pts), ptr(out_n)))
    res = [(keep[f, :out_n[f]].copy(), pts[f, :out_n[f]].copy()) for f in range(F)]
    return res[0] if single else res


class LandmarkRecorder:
    def __init__(self, out_pkl, min_disp_m=2.0, ctx=None):
        self.ctx = ctx or _lib.default_context()
        self.out_pkl = out_pkl
        self.min_disp_m = min_disp_m
        self.last_landmark_pose_world = None
        self.landmarks = []

    def tick(self, kpts_xy, desc, depth_mm, base_pose, ts):
        """One evaluation of recorder._tick (lines 219-298) given this tick's ORB output.  Returns the appended
        record, or None (displacement below min_disp_m, no features, fewer than 30 liftable keypoints)."""
        cam_pose = base_to_cam_world(*base_pose)
        cx, cy = cam_pose[0], cam_pose[1]
        if self.last_landmark_pose_world is None:
            disp = float('inf')
        else:
            lx, ly, _ = self.last_landmark_pose_world[:3]
            disp = math.hypot(cx - lx, cy - ly)
        if disp < self.min_disp_m:
            return None
        if desc is None or len(kpts_xy) == 0:
            return None
        kpts_xy = np.asarray(kpts_xy, dtype=np.float32).reshape(-1, 2)
        keep, pts3 = lift_keypoints(kpts_xy, depth_mm, ctx=self.ctx)
        if len(keep) < MIN_POINTS:
            return None
        record = {'pose': cam_pose, 'descriptors': np.asarray(desc)[keep], 'keypoints_2d': kpts_xy[keep],
                  'keypoints_3d_cam': pts3, 'ts': ts, 'n_features': int(len(keep))}
        self.landmarks.append(record)
        self.last_landmark_pose_world = cam_pose
        return record

    def tick_image(self, bgr, depth_mm, base_pose, ts):
        """recorder._tick from the camera image onwards (lines 240-246): BGR -> gray -> ORB(500) on the GPU (orb.py,
        bit-identical to cv2's detectAndCompute), then `tick`."""
        if getattr(self, 'orb', None) is None:
            from .orb import ORB
            self.orb = ORB(nfeatures=500, ctx=self.ctx)                              # recorder:159
        kp, desc, n = self.orb.detect_and_compute_batch(np.asarray(bgr)[None])
        m = int(n[0])
        return self.tick(kp[0, :m, :2], desc[0, :m] if m else None, depth_mm, base_pose, ts)

    def as_pkl_dict(self):
        return {'intrinsics': {'fx': FX, 'fy': FY, 'cx': CX, 'cy': CY, 'width': W, 'height': H},
                'base_to_cam_translation': BASE_TO_CAM_TRANSLATION.tolist(),
                'base_to_cam_rot': BASE_TO_CAM_ROT.tolist(), 'landmarks': self.landmarks}

    def save(self):
        """landmarks.pkl exactly as recorder._save (lines 313-325) writes it."""
        if not self.landmarks:
            return None
        os.makedirs(os.path.dirname(self.out_pkl), exist_ok=True)
        with open(self.out_pkl, 'wb') as f:
            pickle.dump(self.as_pkl_dict(), f)
        return self.out_pkl

    def save_packed(self, path):
        return write_packed(self.as_pkl_dict(), path)


def write_packed(pkl_dict, path):
    """The landmark set as one flat little-endian file: 8-byte magic, i64 n_kf, i64 N, f64[6] intrinsics, f64[3]
    base->cam translation, f64[9] rotation, then i32[n_kf+1] row offsets, f64[n_kf,7] poses, f64[n_kf] ts,
    u8[N,32] descriptors, f32[N,3] 3-D points, f32[N,2] 2-D keypoints - the arrays nclt_lib_create takes, in order."""
    lms = pkl_dict['landmarks']
    counts = np.array([len(lm['descriptors']) for lm in lms], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    N = int(offs[-1])
    intr = pkl_dict['intrinsics']
    with open(path, 'wb') as f:
        f.write(PACKED_MAGIC)
        np.array([len(lms), N], dtype='<i8').tofile(f)
        np.array([intr[k] for k in ('fx', 'fy', 'cx', 'cy', 'width', 'height')], dtype='<f8').tofile(f)
        np.asarray(pkl_dict['base_to_cam_translation'], dtype='<f8').reshape(3).tofile(f)
        np.asarray(pkl_dict['base_to_cam_rot'], dtype='<f8').reshape(9).tofile(f)
        offs.astype('<i4').tofile(f)
        np.array([lm['pose'] for lm in lms], dtype='<f8').reshape(len(lms), 7).tofile(f)
        np.array([lm.get('ts', 0.0) or 0.0 for lm in lms], dtype='<f8').tofile(f)
        for key, dt, w in (('descriptors', np.uint8, 32), ('keypoints_3d_cam', '<f4', 3), ('keypoints_2d', '<f4', 2)):
            for lm in lms:
                np.ascontiguousarray(lm[key], dtype=dt).reshape(-1, w).tofile(f)
    return path


def read_packed(path):
    """-> dict(offsets i32[n_kf+1], poses f64[n_kf,7], ts, descriptors u8[N,32], points3d f32[N,3], keypoints_2d,
    intrinsics, base_to_cam_translation, base_to_cam_rot); the big arrays are memory maps (no per-keyframe copies)."""
    with open(path, 'rb') as f:
        if f.read(8) != PACKED_MAGIC:
            raise ValueError(f'{path}: not a packed landmark library')
        n_kf, N = (int(v) for v in np.fromfile(f, dtype='<i8', count=2))
        intr = np.fromfile(f, dtype='<f8', count=6)
        t = np.fromfile(f, dtype='<f8', count=3)
        R = np.fromfile(f, dtype='<f8', count=9).reshape(3, 3)
        pos = f.tell()
    out = {'intrinsics': dict(zip(('fx', 'fy', 'cx', 'cy', 'width', 'height'), intr.tolist())),
           'base_to_cam_translation': t, 'base_to_cam_rot': R}
    for name, dt, shape in (('offsets', '<i4', (n_kf + 1,)), ('poses', '<f8', (n_kf, 7)), ('ts', '<f8', (n_kf,)),
                            ('descriptors', np.uint8, (N, 32)), ('points3d', '<f4', (N, 3)), ('keypoints_2d', '<f4', (N, 2))):
        cnt = int(np.prod(shape))
        out[name] = np.memmap(path, dtype=dt, mode='r', offset=pos, shape=shape) if cnt else np.zeros(shape, dtype=dt)
        pos += cnt * np.dtype(dt).itemsize
    return out
