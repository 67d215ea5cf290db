"""Seeded synthetic inputs of the BASELINE.json shapes.

The reference ships no landmarks.pkl, no bag and no depth frames, so
tests, bench.py and the golden-vector generator all draw from these generators.
Everything is `numpy.random.default_rng(seed)`; the same arrays go to the CPU
reference and to the GPU path.

Record layout follows visual_landmark_recorder.py:290-297,319-325 (the pickle
schema consumed at visual_landmark_matcher.py:179-187).
"""
import math

import numpy as np

FX = FY = 320.0
CX, CY = 320.0, 240.0
IMG_W, IMG_H = 640, 480


def rodrigues(rvec):
    """rvec[3] -> R[3,3] float64 (plain Rodrigues formula)."""
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    th = float(np.linalg.norm(r))
    if th < 1e-12:
        return np.eye(3)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(th) * Kx + (1 - math.cos(th)) * (Kx @ Kx)


def project(p3d, rvec, tvec):
    """Pinhole projection with the matcher's K (matcher:49-52), float64."""
    R = rodrigues(rvec)
    pc = p3d.astype(np.float64) @ R.T + np.asarray(tvec, dtype=np.float64).reshape(1, 3)
    u = FX * pc[:, 0] / pc[:, 2] + CX
    v = FY * pc[:, 1] / pc[:, 2] + CY
    return np.stack([u, v], axis=-1)


def make_library(seed, n_kf=400, n_desc=1000, ragged=False, route_len_m=400.0):
    """A landmarks.pkl-shaped dict: n_kf keyframes, n_desc descriptors each.

    ragged=True draws per-keyframe sizes in [30, n_desc] (real landmarks have
    30-500 descriptors, SURVEY section 7 hard parts)."""
    rng = np.random.default_rng(seed)
    lms = []
    for k in range(n_kf):
        n = int(rng.integers(30, n_desc + 1)) if ragged else n_desc
        desc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        p3 = np.stack([rng.uniform(-6, 6, n), rng.uniform(-1, 1.5, n),
                       rng.uniform(0.5, 15, n)], axis=-1).astype(np.float32)
        # camera-optical pose in world: travelling along +x, yaw 0
        # base FLU -> cam optical RDF rotation (matcher:103-107) as quaternion
        x = route_len_m * k / max(n_kf, 1)
        pose = (x + 0.35, 0.0, 0.18, -0.5, 0.5, -0.5, 0.5)
        k2d = np.stack([rng.uniform(0, IMG_W, n), rng.uniform(0, IMG_H, n)], axis=-1).astype(np.float32)
        lms.append({'pose': pose, 'descriptors': desc, 'keypoints_2d': k2d,
                    'keypoints_3d_cam': p3, 'ts': float(k), 'n_features': n})
    return {
        'intrinsics': {'fx': FX, 'fy': FY, 'cx': CX, 'cy': CY, 'width': IMG_W, 'height': IMG_H},
        'base_to_cam_translation': [0.35, 0.0, 0.18],
        'base_to_cam_rot': [[0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]],
        'landmarks': lms,
    }


def make_frame(lib, seed, k_star=None, n_desc=1000, n_planted=500, flip_p=0.06,
               px_sigma=0.5, outlier_frac=0.25, outlier_px=80.0, low_entropy=False):
    """One query frame with `n_planted` true correspondences into keyframe k_star.

    Returns dict(desc u8[n,32], pts2d f32[n,2], k_star, rvec, tvec, q_slots, t_idx).
    low_entropy=True confines every descriptor byte to {0,1} -> massive distance ties
    (exercises the lowest-index tie rule)."""
    rng = np.random.default_rng(seed)
    lms = lib['landmarks']
    if k_star is None:
        k_star = seed % len(lms)
    hi = 2 if low_entropy else 256
    desc = rng.integers(0, hi, size=(n_desc, 32), dtype=np.uint8)
    pts2d = np.stack([rng.uniform(0, IMG_W, n_desc), rng.uniform(0, IMG_H, n_desc)],
                     axis=-1).astype(np.float32)
    lm = lms[k_star]
    nt = len(lm['descriptors'])
    m = min(n_planted, nt, n_desc)
    q_slots = rng.permutation(n_desc)[:m]
    t_idx = rng.permutation(nt)[:m]
    rvec = rng.normal(0, 0.05, 3)
    tvec = rng.normal(0, 0.4, 3)
    if m:
        src = lm['descriptors'][t_idx]
        flips = np.packbits(rng.random((m, 256)) < flip_p, axis=1)
        desc[q_slots] = src ^ flips
        uv = project(lm['keypoints_3d_cam'][t_idx], rvec, tvec)
        uv += rng.normal(0, px_sigma, uv.shape)
        out = rng.random(m) < outlier_frac
        uv[out] += rng.uniform(-outlier_px, outlier_px, (int(out.sum()), 2))
        pts2d[q_slots] = uv.astype(np.float32)
    return {'desc': desc, 'pts2d': pts2d, 'k_star': int(k_star), 'rvec': rvec, 'tvec': tvec,
            'q_slots': q_slots, 't_idx': t_idx}


def make_frame_batch(lib, seeds, **kw):
    """Stack make_frame() outputs: desc u8[B,n,32], pts2d f32[B,n,2], k_star i32[B]."""
    fr = [make_frame(lib, int(s), **kw) for s in seeds]
    return (np.stack([f['desc'] for f in fr]), np.stack([f['pts2d'] for f in fr]),
            np.array([f['k_star'] for f in fr], dtype=np.int32), fr)


def make_pnp_problem(seed, n=200, outlier_frac=0.3, px_sigma=0.6):
    """A standalone PnP-RANSAC problem (obj f32[n,3], img f32[n,2], rvec, tvec)."""
    rng = np.random.default_rng(seed)
    obj = np.stack([rng.uniform(-6, 6, n), rng.uniform(-1, 1.5, n),
                    rng.uniform(0.5, 15, n)], axis=-1).astype(np.float32)
    rvec = rng.normal(0, 0.05, 3)
    tvec = rng.normal(0, 0.4, 3)
    uv = project(obj, rvec, tvec) + rng.normal(0, px_sigma, (n, 2))
    out = rng.random(n) < outlier_frac
    uv[out] = np.stack([rng.uniform(0, IMG_W, int(out.sum())),
                        rng.uniform(0, IMG_H, int(out.sum()))], axis=-1)
    return obj, uv.astype(np.float32), rvec, tvec


# ---------------------------------------------------------------------------
# teach-map inputs (config 3)
# ---------------------------------------------------------------------------

def boustrophedon_path(n_frames, step_m=0.05, x0=-100.0, x1=75.0, y0=-35.0, lane_m=6.0):
    """Planar poses (x, y, yaw) along a serpentine inside the run_teach.sh:29 extent."""
    poses = []
    x, y, direction = x0, y0, 1.0
    turning = 0.0
    for _ in range(n_frames):
        if turning > 0:
            y += step_m
            turning -= step_m
            yaw = math.pi / 2
            if turning <= 0:
                direction = -direction
        else:
            x += direction * step_m
            yaw = 0.0 if direction > 0 else math.pi
            if (direction > 0 and x >= x1) or (direction < 0 and x <= x0):
                turning = lane_m
        poses.append((x, y, yaw))
    return poses


def camera_link_transform(x, y, yaw):
    """map -> camera_link as (tx,ty,tz,qx,qy,qz,qw): planar base pose composed with the
    static base_link->camera_link offset (0.5, 0, 0.48) (tf_wall_clock_relay.py:63-69)."""
    c, s = math.cos(yaw), math.sin(yaw)
    return (x + c * 0.5, y + s * 0.5, 0.48, 0.0, 0.0, math.sin(yaw / 2), math.cos(yaw / 2))


def make_depth_frame(seed, pose, cyl_density=0.02, h=IMG_H, w=IMG_W, bad_frac=0.02):
    """Analytic ray-cast of a ground plane at z=0 (camera 0.48 m up, looking along +x of
    camera_link) plus random vertical cylinders; 2% of pixels set to 0/inf/NaN."""
    rng = np.random.default_rng(seed)
    x, y, yaw = pose
    u = (np.arange(w, dtype=np.float64) - CX) / FX
    v = (np.arange(h, dtype=np.float64) - CY) / FY
    uu, vv = np.meshgrid(u, v)
    # optical ray (x right, y down, z fwd); ground: y_down * z = 0.48
    with np.errstate(divide='ignore', invalid='ignore'):
        z_ground = np.where(vv > 1e-6, 0.48 / vv, np.inf)
    depth = z_ground
    # cylinders in camera-forward planar frame: forward f in [1,12], lateral l in [-6,6]
    n_cyl = rng.poisson(cyl_density * 12 * 12)
    cyl_rng = np.random.default_rng(((int(x * 10) * 73856093) ^ (int(y * 10) * 19349663)) & 0x7FFFFFFF)
    for _ in range(n_cyl):
        f = cyl_rng.uniform(1.0, 12.0)
        l = cyl_rng.uniform(-6.0, 6.0)
        r = cyl_rng.uniform(0.3, 0.7)
        # ray in plane: (f_dir=1, l_dir=-uu) * z ; intersect circle
        a = 1.0 + uu * uu
        b = -2.0 * (f + (-uu) * l)
        c = f * f + l * l - r * r
        disc = b * b - 4 * a * c
        with np.errstate(invalid='ignore'):
            zc = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), np.inf)
        zc = np.where(zc > 0.05, zc, np.inf)
        # cylinder is 2.5 m tall from the ground: height of hit = 0.48 - vv*z
        hgt = 0.48 - vv * zc
        zc = np.where((hgt >= 0) & (hgt <= 2.5), zc, np.inf)
        depth = np.minimum(depth, zc)
    depth = depth.astype(np.float32)
    bad = rng.random((h, w)) < bad_frac
    kinds = rng.integers(0, 3, size=(h, w))
    depth[bad & (kinds == 0)] = 0.0
    depth[bad & (kinds == 1)] = np.inf
    depth[bad & (kinds == 2)] = np.nan
    return depth


def make_camera_frame(seed, h=IMG_H, w=IMG_W, bgr=False, n_rect=300, noise=4.0):
    """A synthetic textured camera image for ORB: overlapping constant rectangles (corners and
    edges at every scale) plus sensor noise, uint8 [h,w] or BGR [h,w,3]."""
    rng = np.random.default_rng(seed)
    ch = 3 if bgr else 1
    img = np.zeros((h, w, ch), np.float32)
    for _ in range(n_rect):
        x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
        rw, rh = (int(v) for v in rng.integers(5, 80, 2))
        img[y:y + rh, x:x + rw] += rng.uniform(-60, 60, ch).astype(np.float32)
    img += 128 + rng.normal(0, noise, (h, w, ch)).astype(np.float32)
    out = np.clip(img, 0, 255).astype(np.uint8)
    return out if bgr else out[..., 0]
