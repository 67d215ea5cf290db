"""TEST INFRASTRUCTURE (never imported by the product) - CPU restatement of the teach-time keypoint
lifting of scripts/common/visual_landmark_recorder.py:247-291: ORB keypoints +
the aligned depth image -> the 3-D points of one `landmarks.pkl` record.

Scalar float32 arithmetic, spelled out operation by operation (NumPy's float32 pairwise summation
inside `ndarray.std()` included), so that the CUDA kernel can be held to bit-exactness.  Pinned
against the reference node itself (`tests/golden/lift_golden.npz`, made by
oracle/make_golden_ref.py::golden_lift, which runs VisualLandmarkRecorder._tick unmodified under
ROS stubs) and against `np.std` (tests/test_oracle_lifting.py).
"""
import numpy as np

F32 = np.float32

# visual_landmark_recorder.py:53-60
FX, FY, CX, CY = 320.0, 320.0, 320.0, 240.0
W, H = 640, 480
DEPTH_MIN_M, DEPTH_MAX_M, DEPTH_VAR_MAX_M = 0.5, 15.0, 0.30
GROUND_Y_THRESHOLD = 180
MIN_POINTS = 30              # recorder:269 `if ok.sum() < 30: return`


def np_sum_f32(a):
    """NumPy's float32 add-reduction of a short contiguous vector (pairwise_sum in
    numpy/_core/src/umath/loops_utils.h.src): n < 8 sequential from 0; 8 <= n <= 128 eight
    accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the remainder in order."""
    n = len(a)
    if n < 8:
        res = F32(0.0)
        for v in a:
            res = F32(res + F32(v))
        return res
    r = [F32(a[i]) for i in range(8)]
    i = 8
    while i < n - (n % 8):
        for j in range(8):
            r[j] = F32(r[j] + F32(a[i + j]))
        i += 8
    res = F32(F32(F32(r[0] + r[1]) + F32(r[2] + r[3])) + F32(F32(r[4] + r[5]) + F32(r[6] + r[7])))
    while i < n:
        res = F32(res + F32(a[i]))
        i += 1
    return res


def np_std_f32(a):
    """ndarray.std() of a float32 vector (numpy/_core/_methods.py::_var, ddof = 0), all in float32."""
    n = len(a)
    mean = F32(np_sum_f32(a) / F32(n))
    sq = [F32(F32(v - mean) * F32(v - mean)) for v in a]
    var = F32(np_sum_f32(sq) / F32(n))
    return F32(np.sqrt(var))


def lift_keypoints(kpts_xy, depth_mm, fx=FX, fy=FY, cx=CX, cy=CY, ground_y=GROUND_Y_THRESHOLD, dmin=DEPTH_MIN_M,
                   dmax=DEPTH_MAX_M, var_max=DEPTH_VAR_MAX_M):
    """kpts_xy f32[N,2] (ORB `kp.pt`), depth_mm u16[H,W] -> (keep i32[M] indices into the N keypoints,
    ascending; pts3d f32[M,3] optical-frame points; d_std f32[N'] diagnostics).  recorder:247-291
    without the `< 30 points` frame gate (the caller applies it)."""
    kpts_xy = np.asarray(kpts_xy, dtype=np.float32).reshape(-1, 2)
    Hh, Ww = depth_mm.shape
    keep, pts = [], []
    for i in range(len(kpts_xy)):
        u = int(np.rint(kpts_xy[i, 0]))                  # np.round: half to even   (recorder:250-251)
        v = int(np.rint(kpts_xy[i, 1]))
        if not (1 <= u < Ww - 1 and 1 <= v < Hh - 1 and v > ground_y):     # recorder:252-253
            continue
        d = F32(F32(depth_mm[v, u]) / F32(1000.0))       # recorder:260
        vals = []
        for dv in (-1, 0, 1):                            # recorder:264-267: 3x3 patch, row-major, non-zero only
            for du in (-1, 0, 1):
                p = F32(F32(depth_mm[v + dv, u + du]) / F32(1000.0))
                if p > F32(0.01):
                    vals.append(p)
        std = np_std_f32(vals) if len(vals) >= 3 else F32(999.0)
        if not (d > F32(dmin) and d < F32(dmax) and std < F32(var_max)):   # recorder:268-270
            continue
        # recorder:283-286: int - python float -> float64; x float32 depth -> float64; / float -> float64; cast at the end
        x = (float(u) - cx) * float(d) / fx
        y = (float(v) - cy) * float(d) / fy
        keep.append(i)
        pts.append((F32(x), F32(y), d))
    return np.array(keep, dtype=np.int32), np.array(pts, dtype=np.float32).reshape(-1, 3)


def make_record(kpts_xy, desc, depth_mm, cam_pose, ts, **kw):
    """One `landmarks` entry as recorder:289-296 builds it, or None when fewer than 30 points survive."""
    keep, pts3 = lift_keypoints(kpts_xy, depth_mm, **kw)
    if len(keep) < MIN_POINTS:
        return None
    kpts_xy = np.asarray(kpts_xy, dtype=np.float32).reshape(-1, 2)
    return {'pose': tuple(cam_pose), 'descriptors': np.asarray(desc)[keep], 'keypoints_2d': kpts_xy[keep],
            'keypoints_3d_cam': pts3, 'ts': ts, 'n_features': int(len(keep))}
